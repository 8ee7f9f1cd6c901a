"""Tiny driver for ncu: `warm` untimed steps then one step of a Plummer sphere."""
import sys
sys.path.insert(0, ".")
import nbody_b200
from nbody_b200 import workloads
n = int(sys.argv[1]); warm = int(sys.argv[2]) if len(sys.argv) > 2 else 1
order = int(sys.argv[3]) if len(sys.argv) > 3 else 4
cap = int(sys.argv[4]) if len(sys.argv) > 4 else 8
P = workloads.plummer(n)
sim = nbody_b200.CudaSimulation([1, 1, 1], P, 1e-3, order=order, leaf_capacity=cap)
for _ in range(warm + 1):
    sim.step()
print(sim.stats())
sim.close()
