"""Accuracy / M2L cost of the adaptive-order threshold `low_order_tau`, from the oracle (CPU only): RMS relative
acceleration error against FP64 direct summation on 4096 targets, the fraction of M2L pairs evaluated at order P-1, and the M2L
flop count relative to tau = 0.    python tests/tools/explore_tau.py KIND N CAPACITY [ORDER]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import oracle
from nbody_b200 import workloads

def run(kind, n, cap, order=4, taus=(0.0, 0.13, 0.16, 0.19, 0.22), eps=0.01):
    P = workloads.GENERATORS[kind](n)
    sk, perm = oracle.sort_keys(oracle.morton_keys(P[:, 0:3], [1, 1, 1]))
    Ps = P[perm]
    posq = np.ascontiguousarray(np.concatenate([Ps[:, 0:3], Ps[:, 9:10]], axis=1))
    tr = oracle.Tree(sk, [1, 1, 1], cap)
    tr.traverse(0.5)
    tg = np.linspace(0, n - 1, min(n, 4096)).astype(np.uint32)
    gd = oracle.direct_field(posq, tg, eps)
    flop = {4: (546.0, 229.0), 3: (229.0, 98.0), 2: (98.0, 98.0)}[order]
    for tau in taus:
        g = tr.fmm_field(posq, order, eps, low_order_tau=tau)
        err = np.sqrt(((g[tg] - gd) ** 2).sum() / (gd ** 2).sum())
        lf = tr.low_fraction
        print(f"{kind} n={n} cap={cap} order={order} tau={tau:.2f}: rms_rel={err:.3e} low-order pairs {100 * lf:.0f} % "
              f"M2L flops x{((1 - lf) * flop[0] + lf * flop[1]) / flop[0]:.2f}", flush=True)

if __name__ == "__main__":
    run(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]) if len(sys.argv) > 4 else 4)
