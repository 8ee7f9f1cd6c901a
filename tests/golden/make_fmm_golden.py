"""Generates tests/golden/fmm_path.npz: keys, permutation, octree, interaction lists and FP64 direct-sum accelerations of
small seeded particle sets, computed by the ORACLE (oracle/oracle.cpp). These fixtures pin the restatement against drift
and give the GPU tests a committed answer that does not depend on rebuilding the oracle. The octree in them is the oracle's
(glade is absent: unpinned); the interaction lists are, entry for entry, the ones the reference's own find_interactions kernel
produces on that octree (tests/test_reference_kernels.py asserts it against tests/golden/clref_golden.npz).
Run in the authoring container: python tests/golden/make_fmm_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
import oracle
from nbody_b200 import workloads

CASES = [("uniform", 700, 8), ("plummer", 900, 8), ("two_galaxies", 800, 4), ("plummer", 1500, 32)]


def packed(pairs):
    p = pairs.astype(np.uint64)
    return np.sort(p[:, 0] << np.uint64(32) | p[:, 1])


def build_case(kind, n, cap):
    P = workloads.GENERATORS[kind](n)
    keys = oracle.morton_keys(P[:, 0:3], (1.0, 1.0, 1.0))
    sk, perm = oracle.sort_keys(keys)
    tree = oracle.Tree(sk, (1.0, 1.0, 1.0), cap, 21)
    m2l, p2p = tree.traverse(0.5)
    Ps = P[perm]
    posq = np.ascontiguousarray(np.concatenate([Ps[:, 0:3], Ps[:, 9:10]], axis=1))
    acc = oracle.direct_field(posq, None, 0.01) * (Ps[:, 9] / Ps[:, 8])[:, None]
    out = {"keys": sk, "perm": perm.astype(np.uint32), "m2l": packed(m2l), "p2p": packed(p2p), "acc": acc.astype(np.float64)}
    for name in ("depth", "prefix", "leaf_index", "leaf_count", "has_children", "child_off", "parent_off", "sibling", "geom"):
        out["tree_" + name] = np.asarray(getattr(tree, name))
    return out


def main():
    blob = {"cases": np.array([f"{k}:{n}:{c}" for k, n, c in CASES])}
    for k, n, c in CASES:
        for name, arr in build_case(k, n, c).items():
            blob[f"{k}_{n}_{c}/{name}"] = arr
    path = os.path.join(HERE, "fmm_path.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
