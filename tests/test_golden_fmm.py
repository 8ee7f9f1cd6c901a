"""The committed FMM-path fixtures (tests/golden/fmm_path.npz, written by tests/golden/make_fmm_golden.py from the oracle):
on CPU the oracle must still reproduce them bit for bit (drift guard); on a B200 the CUDA path, called through the C ABI,
must match them — bit-exact keys, permutation, octree and interaction lists, accelerations within 1e-3 RMS of the stored
FP64 direct sum. The lists in the fixture are the reference's own kernel's lists (tests/test_reference_kernels.py); the octree is
the oracle's (glade absent: unpinned), see the generator's header."""
import os

import numpy as np
import pytest

import oracle
from nbody_b200 import workloads

GOLDEN = np.load(os.path.join(os.path.dirname(__file__), "golden", "fmm_path.npz"))
CASES = [tuple(c.split(":")) for c in GOLDEN["cases"]]
TREE_FIELDS = ("depth", "prefix", "leaf_index", "leaf_count", "has_children", "child_off", "parent_off", "sibling", "geom")


def fixture(kind, n, cap):
    pre = f"{kind}_{n}_{cap}/"
    return {k[len(pre):]: GOLDEN[k] for k in GOLDEN.files if k.startswith(pre)}


def packed(pairs):
    p = pairs.astype(np.uint64)
    return np.sort(p[:, 0] << np.uint64(32) | p[:, 1])


def directed(packed_unordered):
    a, b = packed_unordered >> np.uint64(32), packed_unordered & np.uint64(0xFFFFFFFF)
    return np.unique(np.concatenate([a << np.uint64(32) | b, b << np.uint64(32) | a]))


@pytest.mark.parametrize("kind,n,cap", CASES)
def test_oracle_reproduces_the_fixture(kind, n, cap):
    n, cap = int(n), int(cap)
    g = fixture(kind, n, cap)
    P = workloads.GENERATORS[kind](n)
    sk, perm = oracle.sort_keys(oracle.morton_keys(P[:, 0:3], (1.0, 1.0, 1.0)))
    assert np.array_equal(sk, g["keys"]) and np.array_equal(perm, g["perm"])
    tree = oracle.Tree(sk, (1.0, 1.0, 1.0), cap, 21)
    for name in TREE_FIELDS:
        assert np.array_equal(np.asarray(getattr(tree, name)), g["tree_" + name]), name
    m2l, p2p = tree.traverse(0.5)
    assert np.array_equal(packed(m2l), g["m2l"]) and np.array_equal(packed(p2p), g["p2p"])
    Ps = P[perm]
    posq = np.ascontiguousarray(np.concatenate([Ps[:, 0:3], Ps[:, 9:10]], axis=1))
    acc = oracle.direct_field(posq, None, 0.01) * (Ps[:, 9] / Ps[:, 8])[:, None]
    np.testing.assert_allclose(acc, g["acc"], rtol=1e-12, atol=0)   # FP64; the summation is split over threads


@pytest.mark.gpu
@pytest.mark.parametrize("kind,n,cap", CASES)
def test_cuda_path_matches_the_fixture(kind, n, cap):
    import nbody_b200
    n, cap = int(n), int(cap)
    g = fixture(kind, n, cap)
    P = workloads.GENERATORS[kind](n)
    sim = nbody_b200.CudaSimulation([1.0, 1.0, 1.0], P, 1e-3, leaf_capacity=cap, flags=nbody_b200.FLAG_NO_INTEGRATE)
    sim.step()
    assert np.array_equal(sim.keys(), g["keys"]) and np.array_equal(sim.permutation(), g["perm"])
    t = sim.tree()
    for name in TREE_FIELDS:
        assert np.array_equal(t[name], g["tree_" + name]), name
    m2l, p2p = sim.lists()
    assert np.array_equal(packed(m2l), directed(g["m2l"])) and np.array_equal(packed(p2p), directed(g["p2p"]))
    acc = sim.accelerations().astype(np.float64)
    err = float(np.sqrt(((acc - g["acc"]) ** 2).sum() / (g["acc"] ** 2).sum()))
    assert err < 1e-3, err
    sim.close()
