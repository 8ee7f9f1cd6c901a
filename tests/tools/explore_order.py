"""Exploration: FMM accuracy vs expansion order at the reference MAC, list sizes.
Run here (CPU only); results recorded in DESIGN.md."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import oracle
from nbody_b200 import workloads

def run(kind, n, cap, orders, eps=0.01):
    P = workloads.GENERATORS[kind](n)
    keys = oracle.morton_keys(P[:, 0:3], [1, 1, 1])
    sk, perm = oracle.sort_keys(keys)
    Ps = P[perm]
    posq = np.ascontiguousarray(np.concatenate([Ps[:, 0:3], Ps[:, 9:10]], axis=1))
    t0 = time.time()
    tr = oracle.Tree(sk, [1, 1, 1], cap)
    m2l, p2p = tr.traverse(0.5)
    t1 = time.time()
    cnt = tr.leaf_count
    leaves = (tr.has_children == 0) & (cnt > 0)
    p2p_evals = 2 * (cnt[p2p[:, 0]].astype(np.int64) * cnt[p2p[:, 1]]).sum() - (cnt[p2p[p2p[:,0]==p2p[:,1],0]].astype(np.int64)**2).sum() - cnt[leaves].sum()
    print(f"{kind} n={n} cap={cap}: nodes={tr.num_nodes} nonempty={int((cnt>0).sum())} leaves={int(leaves.sum())} maxdepth={tr.depth.max()} "
          f"m2l_pairs={len(m2l)} ({2*len(m2l)/max(1,(cnt>0).sum()):.0f}/node) p2p_pairs={len(p2p)} p2p_evals/particle={p2p_evals/n:.0f} rounds={tr.rounds} [{t1-t0:.1f}s]", flush=True)
    nt = min(n, 4096)
    tg = np.linspace(0, n - 1, nt).astype(np.uint32)
    gd = oracle.direct_field(posq, tg, eps)
    for p in orders:
        t0 = time.time()
        g = tr.fmm_field(posq, p, eps)
        err = np.sqrt(((g[tg] - gd) ** 2).sum() / (gd ** 2).sum())
        rel = np.sqrt(((g[tg] - gd) ** 2).sum(1) / (gd ** 2).sum(1))
        print(f"   p={p}: rms_rel={err:.3e} median={np.median(rel):.2e} p99={np.percentile(rel,99):.2e} max={rel.max():.2e} [{time.time()-t0:.1f}s]", flush=True)

if __name__ == "__main__":
    kind = sys.argv[1]; n = int(sys.argv[2]); cap = int(sys.argv[3])
    orders = [int(x) for x in sys.argv[4].split(",")]
    run(kind, n, cap, orders)
