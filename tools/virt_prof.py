"""Developer tool: the partitioned scheme with W virtual ranks on ONE GPU (every rank's kernels run alone, one after the other, exactly
as they would on W GPUs): per-rank stage times of the last step.  usage: python tools/virt_prof.py N W steps [capacity]"""
import sys
sys.path.insert(0, ".")
import numpy as np
import nbody_b200
from nbody_b200 import workloads
n, W, steps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
cap = int(sys.argv[4]) if len(sys.argv) > 4 else 48
g = nbody_b200.VirtualGroup([1, 1, 1], workloads.plummer(n), 1e-3, W, leaf_capacity=cap)
for _ in range(steps):
    g.step()
st = g.stats()
cols = ("ms_sort", "ms_tree", "ms_upsweep", "ms_traverse", "ms_m2l", "ms_l2l", "ms_leaf", "ms_import", "ms_halo", "ms_balance")
for r, s in enumerate(st):
    print(r, s["n_particles"], s["p2p_interactions"], {c[3:]: round(s[c], 3) for c in cols})
print("sum over ranks:", {c[3:]: round(sum(s[c] for s in st), 2) for c in cols}, "p2p", sum(s["p2p_interactions"] for s in st), flush=True)
g.close()
