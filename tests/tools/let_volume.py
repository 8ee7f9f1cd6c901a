"""What a locally-essential tree would have to move: for a Morton-range partition of one particle
set over W ranks, the remote multipoles and the remote leaf particles each rank's share of the reference-rule traversal reads,
computed exactly from the oracle's octree and interaction lists (CPU only, no GPU).
    python tests/tools/let_volume.py KIND N CAPACITY WORLD [ORDER]
Definitions: the partition cuts the tree-ordered particle array into W equal parts and snaps each cut down to a leaf boundary
(comm.cu:k_partition). A node is LOCAL to rank r if all its particles are r's, REMOTE if none are, SHARED otherwise (the few
top-tree nodes a cut passes through; their multipoles are sums of per-rank partial sums). For every traversal pair (A, B) both
directions are evaluated (src/field.cl:25-30): rank r needs B's multipole (M2L pair) or B's particles (P2P pair) whenever A holds
particles of r."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import oracle
from nbody_b200 import workloads


def let_volume(kind, n, cap, world, order=4):
    P = workloads.GENERATORS[kind](n)
    t0 = time.time()
    sk, perm = oracle.sort_keys(oracle.morton_keys(P[:, 0:3], (1.0, 1.0, 1.0)))
    tree = oracle.Tree(sk, (1.0, 1.0, 1.0), cap, 21)
    m2l, p2p = tree.traverse(0.5)
    t_trav = time.time() - t0
    begin = np.asarray(tree.leaf_index, np.int64)
    count = np.asarray(tree.leaf_count, np.int64)
    end = begin + count
    childless = np.asarray(tree.has_children) == 0
    # partition: equal counts, each cut moved down to the first particle of the leaf that contains it
    leaf_begin = np.sort(begin[childless & (count > 0)])
    part = np.zeros(world + 1, np.int64)
    part[world] = n
    for r in range(1, world):
        want = n * r // world
        part[r] = leaf_begin[np.searchsorted(leaf_begin, want, side="right") - 1]
    lo_rank = np.searchsorted(part, begin, side="right") - 1            # rank of a node's first particle
    hi_rank = np.searchsorted(part, np.maximum(end - 1, begin), side="right") - 1   # ... and of its last one
    nc = (order + 1) * (order + 2) * (order + 3) // 6
    rows = []
    for r in range(world):
        holds = (count > 0) & (lo_rank <= r) & (hi_rank >= r)             # node contains particles of rank r
        local = (lo_rank == r) & (hi_rank == r)
        def sources_needed(pairs):
            a, b = pairs[:, 0], pairs[:, 1]
            need = np.zeros(len(count), bool)
            need[b[holds[a]]] = True                                      # direction A <- B
            need[a[holds[b]]] = True                                      # direction B <- A
            return need
        need_m = sources_needed(m2l) & ~local
        need_p = sources_needed(p2p) & ~local                             # P2P partners are childless nodes: local or remote, never shared
        shared_m = need_m & holds
        remote_m = need_m & ~holds
        rows.append({"rank": r, "particles": int(part[r + 1] - part[r]),
                     "remote_multipoles": int(remote_m.sum()), "shared_multipoles": int(shared_m.sum()),
                     "halo_leaves": int(need_p.sum()), "halo_particles": int(count[need_p].sum()),
                     "let_bytes": int(remote_m.sum() * (4 * nc + 16) + shared_m.sum() * 4 * nc + count[need_p].sum() * 16 + need_p.sum() * 8)})
    repl = 32 * n * (world - 1) // world                                   # today: every rank receives everybody else's positions and velocities
    return {"workload": f"{kind} N={n}", "leaf_capacity": cap, "world": world, "order": order, "nodes": int(len(count)),
            "m2l_pairs": int(len(m2l)), "p2p_pairs": int(len(p2p)), "oracle_seconds": round(t_trav, 1), "per_rank": rows,
            "replicated_exchange_bytes_per_rank": int(repl),
            "let_bytes_per_rank_max": max(x["let_bytes"] for x in rows),
            "halo_fraction_max": max(x["halo_particles"] / max(x["particles"], 1) for x in rows)}


if __name__ == "__main__":
    kind, n, cap, world = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    order = int(sys.argv[5]) if len(sys.argv) > 5 else 4
    print(json.dumps(let_volume(kind, n, cap, world, order)))
