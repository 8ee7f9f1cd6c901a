"""First-light diagnostic on the GPU box: each stage against the oracle, verbose."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import oracle, nbody_b200
from nbody_b200 import workloads
sys.path.insert(0, "tests")
from conftest import sorted_system, rms_rel

def run(kind, n, order=4, cap=8):
    print(f"=== {kind} n={n} order={order} cap={cap}", flush=True)
    P = workloads.GENERATORS[kind](n)
    sim = nbody_b200.CudaSimulation([1, 1, 1], P, 1e-3, order=order, leaf_capacity=cap, flags=nbody_b200.FLAG_NO_INTEGRATE)
    t0 = time.time(); sim.step(); print("step ok", time.time() - t0, sim.stats(), flush=True)
    o = sorted_system(P, capacity=cap)
    k = sim.keys(); print("keys equal:", np.array_equal(k, o["keys"]))
    print("perm equal:", np.array_equal(sim.permutation(), o["perm"]))
    t = sim.tree(); tr = o["tree"]
    print("nodes", len(t["depth"]), tr.num_nodes)
    if len(t["depth"]) == tr.num_nodes:
        for name, ref in (("depth", tr.depth), ("prefix", tr.prefix), ("leaf_index", tr.leaf_index), ("leaf_count", tr.leaf_count),
                          ("has_children", tr.has_children), ("child_off", tr.child_off), ("parent_off", tr.parent_off),
                          ("sibling", tr.sibling), ("geom", tr.geom)):
            print("  tree", name, np.array_equal(t[name], ref))
    m2l_o, p2p_o = tr.traverse(0.5)
    m2l, p2p = sim.lists()
    def directed(u):
        d = np.concatenate([u, u[:, ::-1]], axis=0).astype(np.uint64)
        return np.unique(d[:, 0] << np.uint64(32) | d[:, 1])
    def packed(d):
        d = d.astype(np.uint64); return np.sort(d[:, 0] << np.uint64(32) | d[:, 1])
    a, b = packed(m2l), directed(m2l_o)
    print("m2l directed", len(a), len(b), np.array_equal(a, b), "dups", len(a) - len(np.unique(a)))
    a, b = packed(p2p), directed(p2p_o)
    print("p2p directed", len(a), len(b), np.array_equal(a, b), "dups", len(a) - len(np.unique(a)))
    g_fmm, Mo, Lo = tr.fmm_field(o["posq"], order, 0.01, want_expansions=True)
    M, L = sim.expansions()
    ne = tr.leaf_count > 0
    print("M rel", rms_rel(M[ne], Mo[ne]))
    # oracle L are Taylor coefficients; device stores pure derivatives: multiply by n!
    idx = []
    for oo in range(order + 1):
        for i in range(oo, -1, -1):
            for j in range(oo - i, -1, -1):
                idx.append((i, j, oo - i - j))
    from math import factorial as f
    fac = np.array([f(i) * f(j) * f(k) for i, j, k in idx], np.float64)
    Ld = Lo * fac[None, :]
    print("L rel (orders>=1)", rms_rel(L[ne][:, 1:], Ld[ne][:, 1:]))
    acc = sim.accelerations()
    Ps = o["P"]
    scale = (Ps[:, 9] / Ps[:, 8])[:, None]
    print("acc vs oracle fmm (same lists, fp64):", rms_rel(acc, g_fmm * scale))
    nt = min(n, 4096); tg = np.linspace(0, n - 1, nt).astype(np.uint32)
    gd = oracle.direct_field(o["posq"], tg, 0.01)
    print("acc vs direct:", rms_rel(acc[tg], gd * scale[tg]), " oracle fmm vs direct:", rms_rel(g_fmm[tg], gd), flush=True)
    sim.close()

if __name__ == "__main__":
    import torch
    print(torch.cuda.get_device_name(0))
    run("uniform", 5, 4); run("uniform", 4096, 4); run("uniform", 30000, 3); run("plummer", 30000, 4); run("uniform", 20000, 2, cap=32)
