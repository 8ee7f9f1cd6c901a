"""P2P microbenchmark: the all-pairs tiled kernel (k_direct) on n_tgt x n_src, FP32 TFLOP/s at 20 flop per pair."""
import sys
import numpy as np
sys.path.insert(0, ".")
import nbody_b200
from nbody_b200 import workloads
n_src = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
n_tgt = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 17
P = workloads.plummer(n_src)
posq = np.ascontiguousarray(np.concatenate([P[:, 0:3], P[:, 9:10]], axis=1))
f, ms = nbody_b200.direct_field(posq, posq[:n_tgt], 0.01, repeats=3)
print(f"k_direct: {n_tgt} x {n_src} pairs in {ms:.3f} ms -> {20.0 * n_tgt * n_src / (ms * 1e-3) / 1e12:.2f} TFLOP/s (20 flop/pair)")
