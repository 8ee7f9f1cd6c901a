"""CPU checks of tests/tools/let_volume.py, the exact locally-essential-tree accounting used in DESIGN section 10: what each rank
of a Morton-range partition reads from other ranks under the reference-rule traversal (oracle lists)."""
import importlib.util
import os

import numpy as np

import oracle
from nbody_b200 import workloads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("let_volume", os.path.join(ROOT, "tests", "tools", "let_volume.py"))
lv = importlib.util.module_from_spec(spec)
spec.loader.exec_module(lv)


def test_single_rank_needs_nothing_remote():
    d = lv.let_volume("plummer", 6000, 8, 1)
    r = d["per_rank"][0]
    assert r["particles"] == 6000 and r["remote_multipoles"] == 0 and r["shared_multipoles"] == 0 and r["halo_particles"] == 0
    assert d["replicated_exchange_bytes_per_rank"] == 0


def test_partition_tiles_and_needs_are_bounded():
    n, world, cap = 20000, 4, 8
    d = lv.let_volume("uniform", n, cap, world)
    rows = d["per_rank"]
    assert sum(x["particles"] for x in rows) == n and all(abs(x["particles"] - n // world) <= cap for x in rows)   # cuts snap to leaves
    for x in rows:
        assert 0 < x["halo_particles"] <= n - x["particles"]          # a halo exists and holds only other ranks' particles
        assert 0 < x["remote_multipoles"] < d["nodes"] and x["shared_multipoles"] <= 21 * (world - 1)   # a cut shares one node per level
        assert x["let_bytes"] < d["replicated_exchange_bytes_per_rank"] * 2


def test_halo_matches_a_brute_force_count():
    """Independent recount for one rank: the halo is every particle of a foreign leaf that appears opposite one of the rank's
    leaves in the P2P list."""
    n, world, cap = 3000, 3, 8
    d = lv.let_volume("plummer", n, cap, world)
    P = workloads.plummer(n)
    sk, _ = oracle.sort_keys(oracle.morton_keys(P[:, 0:3], (1.0, 1.0, 1.0)))
    tree = oracle.Tree(sk, (1.0, 1.0, 1.0), cap, 21)
    _, p2p = tree.traverse(0.5)
    begin, count = np.asarray(tree.leaf_index, np.int64), np.asarray(tree.leaf_count, np.int64)
    cuts = np.cumsum([0] + [x["particles"] for x in d["per_rank"]])
    for r in range(world):
        mine = lambda node: cuts[r] <= begin[node] < cuts[r + 1]
        halo = set()
        for a, b in p2p.tolist():
            if mine(a) and not mine(b):
                halo.add(b)
            if mine(b) and not mine(a):
                halo.add(a)
        assert sum(int(count[x]) for x in halo) == d["per_rank"][r]["halo_particles"]
        assert len(halo) == d["per_rank"][r]["halo_leaves"]
