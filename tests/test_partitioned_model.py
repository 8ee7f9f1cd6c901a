"""CPU test (gloo, world_size 2 and 3) of the partitioned tree build's rule (tests/models/partitioned_tree_model.py = the rule of
nbody_b200/csrc/let.cu + tree.cu:node_splits): the ranks all-gather their local counts of the cells that straddle a splitter, split
those by the global count, and every rank's tree must then be the GLOBAL octree (the oracle's, built from all keys) restricted to the
cells that hold its own particles. Also: workload slices of every generator tile the global set (what bench.py hands each rank)."""
import json
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "models"))
import oracle
from nbody_b200 import workloads
import partitioned_tree_model as M


def _worker(rank, world, port, kind, n, cap, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P = workloads.GENERATORS[kind](n)
    keys = np.sort(oracle.morton_keys(P[:, 0:3], [1, 1, 1]))
    # splitters: particle-count quantiles, deliberately NOT aligned with any cell boundary
    split = [0] + [int(keys[n * k // world]) for k in range(1, world)] + [M.KEY_END]
    mine = keys[(keys >= np.uint64(split[rank])) & (keys < np.uint64(split[rank + 1]) if split[rank + 1] < M.KEY_END else True)]
    cells = M.straddling_cells(split)
    allc = [None] * world
    dist.all_gather_object(allc, M.local_straddle_counts(mine, split))           # X2 of let.cu
    glob = np.sum(np.array(allc, dtype=np.int64), axis=0)
    forced = frozenset((d, p) for (b, d, p), c in zip(cells, glob) if c > cap)
    local = M.build_tree(mine, cap, M.MAX_DEPTH, forced)
    ref = M.build_tree(keys, cap, M.MAX_DEPTH)                                   # the global octree
    bad = []
    for cell, (c, split_here) in local.items():
        if c == 0:
            continue                                                             # empty siblings exist only to keep groups of 8 whole
        if cell not in ref or ref[cell][1] != split_here:
            bad.append((cell, c, split_here, ref.get(cell)))
    missing = [cell for cell, (c, _) in ref.items() if M.count_in(mine, *M.cell_range(*cell)) > 0 and cell not in local]
    with open(os.path.join(tmp, f"r{rank}.json"), "w") as f:
        json.dump({"bad": bad[:5], "missing": missing[:5], "n_local": int(len(mine)), "n_forced": len(forced), "n_cells": len(local),
                   "ref_cells": len(ref), "straddling_global": [int(x) for x in glob]}, f)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("kind,n,cap,world,port", [("plummer", 3000, 8, 2, 29541), ("uniform", 2000, 3, 3, 29542), ("two_galaxies", 3000, 16, 2, 29543)])
def test_every_ranks_tree_is_the_global_octree_restricted_to_its_particles(tmp_path, kind, n, cap, world, port):
    mp.spawn(_worker, args=(world, port, kind, n, cap, str(tmp_path)), nprocs=world, join=True)
    total = 0
    for r in range(world):
        d = json.load(open(tmp_path / f"r{r}.json"))
        assert d["bad"] == [] and d["missing"] == [], d
        assert d["n_forced"] >= 1                      # the root at least straddles every splitter
        total += d["n_local"]
    assert total == n


def test_model_tree_equals_the_oracle_tree():
    """The numpy model of the split rule, without forced cells, is the oracle's octree (tests/test_oracle.py pins that one)."""
    P = workloads.plummer(2500)
    keys = np.sort(oracle.morton_keys(P[:, 0:3], [1, 1, 1]))
    t = oracle.Tree(keys, [1, 1, 1], 8)
    m = M.build_tree(keys, 8, M.MAX_DEPTH)
    assert len(m) == t.num_nodes
    mine = sorted((d, p, c, int(s)) for (d, p), (c, s) in m.items())
    sh = [3 * (M.MAX_DEPTH - int(d)) for d in t.depth]
    theirs = sorted((int(d), int(k) >> s, int(c), int(h)) for d, k, s, c, h in zip(t.depth, t.prefix, sh, t.leaf_count, t.has_children))
    assert mine == theirs


@pytest.mark.parametrize("kind", ["uniform", "plummer", "two_galaxies"])
def test_generate_slices_tile_the_global_set(kind):
    n = 10007
    full = workloads.generate(kind, n)
    assert np.array_equal(full, workloads.GENERATORS[kind](n))
    cuts = [0, 1, 1234, 5003, 5004, 9999, n]
    parts = [workloads.generate(kind, n, a, b - a) for a, b in zip(cuts, cuts[1:])]
    assert np.array_equal(np.concatenate(parts), full)


def _lists_worker(rank, world, port, n, cap, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P = workloads.plummer(n)
    keys = np.sort(oracle.morton_keys(P[:, 0:3], [1, 1, 1]))
    split = [0] + [int(keys[n * k // world]) for k in range(1, world)] + [M.KEY_END]
    lo = [int(np.searchsorted(keys, np.uint64(split[r]), "left")) if split[r] < M.KEY_END else n for r in range(world + 1)]
    mine = keys[lo[rank]:lo[rank + 1]]
    cells = M.straddling_cells(split)
    allc = [None] * world
    dist.all_gather_object(allc, M.local_straddle_counts(mine, split))
    glob = np.sum(np.array(allc, dtype=np.int64), axis=0)
    forced = frozenset((d, p) for (b, d, p), c in zip(cells, glob) if c > cap)
    local = M.build_tree(mine, cap, M.MAX_DEPTH, forced)
    trees = [None] * world
    dist.all_gather_object(trees, local)                                         # X3: every rank's tree (here: the whole dict)
    ref = M.build_tree(keys, cap, M.MAX_DEPTH)

    def members(cell, a, b):                                                     # global particle indices of keys[a:b) inside `cell`
        c_lo, c_hi = M.cell_range(*cell)
        i0 = a + int(np.searchsorted(keys[a:b], np.uint64(c_lo), "left"))
        i1 = a + (int(np.searchsorted(keys[a:b], np.uint64(c_hi), "left")) if c_hi < (1 << 64) else b - a)
        return i0, i1

    # the rank's own targets against every rank's tree (own first, then the imported ones: the seeds of k_traverse_init)
    pairs_p2p, covered = set(), np.zeros(lo[rank + 1] - lo[rank], np.int64)
    for s in range(world):
        m2l, p2p = M.traverse(local, trees[s])
        for (ca, cb), near in [(x, False) for x in m2l] + [(x, True) for x in p2p]:
            t0, t1 = members(ca, lo[rank], lo[rank + 1])
            s0, s1 = members(cb, lo[s], lo[s + 1])
            covered[t0 - lo[rank]:t1 - lo[rank]] += s1 - s0
            if near:
                pairs_p2p.update((i, j) for i in range(t0, t1) for j in range(s0, s1))
    # the single-GPU lists, restricted to this rank's target particles
    m2l_g, p2p_g = M.traverse(ref, ref)
    ref_p2p = set()
    for ca, cb in p2p_g:
        t0, t1 = members(ca, 0, n)
        s0, s1 = members(cb, 0, n)
        t0, t1 = max(t0, lo[rank]), min(t1, lo[rank + 1])
        if t0 < t1:
            ref_p2p.update((i, j) for i in range(t0, t1) for j in range(s0, s1))
    with open(os.path.join(tmp, f"l{rank}.json"), "w") as f:
        json.dump({"covered_once": bool(np.all(covered == n)), "p2p_equal": pairs_p2p == ref_p2p, "n_p2p": len(pairs_p2p),
                   "n_local": int(len(mine))}, f)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n,cap,world,port", [(1200, 8, 2, 29551), (900, 4, 3, 29552)])
def test_own_targets_against_all_ranks_trees_give_the_global_lists(tmp_path, n, cap, world, port):
    """The second half of the construction: traversing a rank's own targets against (own tree + every other rank's tree) covers every
    (own target particle, source particle) pair exactly once by an M2L or a P2P interaction, and the P2P particle pairs are exactly those
    of the single-GPU traversal of the global tree — which is why the P2P evaluation counts of the ranks sum to the single-GPU count."""
    mp.spawn(_lists_worker, args=(world, port, n, cap, str(tmp_path)), nprocs=world, join=True)
    total = 0
    for r in range(world):
        d = json.load(open(tmp_path / f"l{r}.json"))
        assert d["covered_once"] and d["p2p_equal"] and d["n_p2p"] > 0, d
        total += d["n_local"]
    assert total == n
