"""Full-size accuracy check: RMS relative acceleration error of one FMM force evaluation against direct
summation (GPU all-pairs kernel, FP32) on a fixed 65,536-target subsample, FP64 CPU oracle on 256 of them."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import nbody_b200, oracle
from nbody_b200 import workloads
kind = sys.argv[1]; n = int(sys.argv[2]); caps = [int(x) for x in sys.argv[3].split(",")]
P = workloads.GENERATORS[kind](n)
for cap in caps:
    sim = nbody_b200.CudaSimulation([1, 1, 1], P, 1e-3, leaf_capacity=cap, flags=nbody_b200.FLAG_NO_INTEGRATE)
    sim.step()
    acc = sim.accelerations(); out = sim.particles(); st = sim.stats()
    sim.close()
    posq = np.ascontiguousarray(np.concatenate([out[:, 0:3], out[:, 9:10]], axis=1))
    tg = np.linspace(0, n - 1, 65536).astype(np.int64)
    f, ms = nbody_b200.direct_field(posq, posq[tg], 0.01)
    scale = (out[:, 9] / out[:, 8])[:, None]
    ref = f.astype(np.float64) * scale[tg]
    err = np.sqrt(((acc[tg] - ref) ** 2).sum() / (ref ** 2).sum())
    spot = tg[::256].astype(np.uint32)
    t0 = time.time(); gd = oracle.direct_field(posq, spot, 0.01); t1 = time.time()
    e64 = np.sqrt(((acc[spot] - gd * scale[spot]) ** 2).sum() / ((gd * scale[spot]) ** 2).sum())
    print(f"{kind} N={n} capacity={cap}: RMS rel error vs GPU direct (65536 targets) {err:.3e}; vs FP64 oracle ({len(spot)} targets, {t1-t0:.0f}s) {e64:.3e}; "
          f"step {st['ms_total']:.1f} ms, low-order fraction {st['m2l_interactions_low']/max(1,st['m2l_interactions']):.2f}", flush=True)
