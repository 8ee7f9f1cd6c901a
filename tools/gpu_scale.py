"""Scale probe: step timing + stats at growing N, flushing progress."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import nbody_b200
from nbody_b200 import workloads
kind = sys.argv[1]; sizes = [int(x) for x in sys.argv[2].split(",")]
order = int(sys.argv[3]) if len(sys.argv) > 3 else 4
cap = int(sys.argv[4]) if len(sys.argv) > 4 else 8
flags = int(sys.argv[5]) if len(sys.argv) > 5 else 0
for n in sizes:
    t0 = time.time(); P = workloads.GENERATORS[kind](n); print(f"[{kind} n={n}] generated in {time.time()-t0:.1f}s", flush=True)
    t0 = time.time(); sim = nbody_b200.CudaSimulation([1, 1, 1], P, 1e-3, order=order, leaf_capacity=cap, flags=flags); print(f"  created in {time.time()-t0:.2f}s", flush=True)
    for s in range(3):
        t0 = time.time(); sim.step(); dt = time.time() - t0
        st = sim.stats()
        print(f"  step {s}: wall {dt*1e3:.1f} ms  " + " ".join(f"{k}={v:.2f}" if isinstance(v, float) else f"{k}={v}" for k, v in st.items()), flush=True)
    sim.close()
