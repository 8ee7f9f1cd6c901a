"""Single-GPU evidence at the particle counts of BASELINE configs 4 and 5: a few timed steps, then the accuracy of one force
evaluation on the EVOLVED state against direct summation (GPU all-pairs kernel on a 65,536-target subsample, FP64 oracle on
256 of them). Prints one JSON line per run.   python tests/tools/scale_evidence.py KIND N CAPACITY [POOL_SCALE] [STEPS]"""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
import nbody_b200, oracle
from nbody_b200 import workloads

kind, n, cap = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
pool = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
t0 = time.time(); P = workloads.GENERATORS[kind](n); t_gen = time.time() - t0
G = workloads.force_constant(kind, n)
t0 = time.time()
sim = nbody_b200.CudaSimulation([1, 1, 1], P, 1e-3, leaf_capacity=cap, pool_scale=pool, force_constant=G)
t_create = time.time() - t0
del P
rows = []
for s in range(steps):
    t0 = time.time(); sim.step(); wall = time.time() - t0
    st = sim.stats()
    rows.append({"wall_ms": round(1e3 * wall, 1), **{k: (round(v, 2) if isinstance(v, float) else v) for k, v in st.items()}})
    print(f"[{kind} N={n} cap={cap}] step {s}: wall {1e3 * wall:.1f} ms, device {st['ms_total']:.1f} ms, retries {st['retries']}, "
          f"{st['device_bytes'] / 2**30:.1f} GiB on the device", flush=True)
out = sim.particles()
sim.close()
inside = float(np.mean((out[:, 0:3] >= 0).all(1) & (out[:, 0:3] < 1).all(1)))
sim = nbody_b200.CudaSimulation([1, 1, 1], out, 1e-3, leaf_capacity=cap, pool_scale=pool, force_constant=G, flags=nbody_b200.FLAG_NO_INTEGRATE)
sim.step()
acc = sim.accelerations(); srt = sim.particles()
sim.close()
posq = np.ascontiguousarray(np.concatenate([srt[:, 0:3], srt[:, 9:10]], axis=1))
tg = np.linspace(0, n - 1, 65536).astype(np.int64)
f, ms = nbody_b200.direct_field(posq, posq[tg], 0.01)
scale = (G * srt[:, 9] / srt[:, 8])[:, None]
ref = f.astype(np.float64) * scale[tg]
err = float(np.sqrt(((acc[tg] - ref) ** 2).sum() / (ref ** 2).sum()))
spot = tg[::(256 if n <= 1 << 26 else 1024)].astype(np.uint32)
t0 = time.time(); gd = oracle.direct_field(posq, spot, 0.01) * scale[spot]; t_or = time.time() - t0
e64 = float(np.sqrt(((acc[spot] - gd) ** 2).sum() / (gd ** 2).sum()))
last = rows[-1]
print(json.dumps({"workload": f"{kind} N={n}", "leaf_capacity": cap, "pool_scale": pool, "generate_s": round(t_gen, 1), "create_s": round(t_create, 2),
                  "steps": rows, "particle_steps_per_s": n / (last["ms_total"] * 1e-3), "fraction_inside_bounds_after": inside,
                  "rms_rel_error_vs_gpu_direct_65536_targets": err, "rms_rel_error_vs_fp64_oracle_256_targets": e64,
                  "direct_kernel_ms": ms, "direct_kernel_tflops": 20.0 * 65536 * n / (ms * 1e-3) / 1e12, "oracle_s": round(t_or, 1)}), flush=True)
