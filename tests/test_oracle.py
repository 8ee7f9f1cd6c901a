"""CPU tests: the oracle against the reference's own naive path (oracle/_ref, the
unmodified /root/reference/src/naive_simulation.cpp) and the committed golden vectors,
plus internal consistency of the restated tree / traversal / FMM."""
import json
import os

import numpy as np
import pytest

import oracle
from nbody_b200 import workloads
from conftest import sorted_system, rms_rel

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def kat_two_particles():
    P = np.zeros((2, 12), np.float32)
    P[0, 8] = 1.0; P[0, 9] = 1.0
    P[1, 0:3] = (1.0, 0.5, 0.25); P[1, 8] = 2.0; P[1, 9] = 1.0
    return P


def test_naive_restatement_matches_survey_kat():
    # SURVEY 8c KAT (1), derived from the unmodified reference source
    out, t = oracle.naive_step_as_written(kat_two_particles(), -1.0, 0.001, 1)
    assert np.float32(t) == np.float32(0.001)
    np.testing.assert_allclose(out[0, 0:3], [-1.92450102e-07] * 3, rtol=1e-6)
    np.testing.assert_allclose(out[0, 4:7], [-0.000192450098] * 3, rtol=1e-6)
    np.testing.assert_allclose(out[1, 0:3], [1.00000012, 0.500000119, 0.250000089], rtol=1e-7)
    np.testing.assert_allclose(out[1, 4:7], [9.62250488e-05] * 3, rtol=1e-6)


def test_naive_restatement_matches_golden_files():
    with open(os.path.join(GOLDEN, "naive_kat.json")) as f:
        kat = json.load(f)
    assert len(kat["cases"]) >= 4
    for case in kat["cases"]:
        P = np.array(case["particles_in"], np.float32)
        out, t = oracle.naive_step_as_written(P, case["force_constant"], case["dt"], case["steps"])
        ref = np.array(case["particles_out"], np.float32)
        assert np.array_equal(out.view(np.uint32), ref.view(np.uint32)), case["name"]
        assert np.float32(t) == np.float32(case["time"])


@pytest.mark.skipif(oracle.ref_lib() is None, reason="oracle/_ref not built (no /root/reference here)")
def test_naive_restatement_bit_exact_vs_reference_binary():
    assert oracle.ref_lib().ref_particle_size() == 48
    rng = np.random.default_rng(7)
    for n, steps in ((2, 1), (17, 3), (256, 2)):
        P = np.zeros((n, 12), np.float32)
        P[:, 0:3] = rng.random((n, 3)); P[:, 4:7] = 0.1 * (rng.random((n, 3)) - 0.5)
        P[:, 8] = 1 + 9 * rng.random(n); P[:, 9] = 0.1 + 0.9 * rng.random(n)
        a, ta = oracle.naive_step_as_written(P, 1.0, 0.001, steps)
        b, tb = oracle.ref_naive_run(P, 1.0, 0.001, steps)
        assert np.array_equal(a[:, :10].view(np.uint32), b[:, :10].view(np.uint32))
        assert ta == tb


def test_morton_keys_definition():
    pos = np.array([[0, 0, 0], [0.999999, 0, 0], [0, 0.5, 0], [0, 0, 0.5], [1.5, -1, 0.25]], np.float32)
    k = oracle.morton_keys(pos, [1, 1, 1])
    assert k[0] == 0
    assert k[2] == np.uint64(2) << np.uint64(60)   # y bit of the level-1 digit
    assert k[3] == np.uint64(4) << np.uint64(60)   # z is the most significant bit of each digit
    # clamped outside the box: x -> 2^21-1, y -> 0; z = 0.25 -> bit 19 of z -> key bit 3*19+2
    x = (2 ** 21 - 1)
    exp = sum(((x >> b) & 1) << (3 * b) for b in range(21)) | (1 << 59)
    assert int(k[4]) == exp


def test_sort_is_stable():
    keys = np.array([5, 1, 5, 1, 0, 5], np.uint64)
    sk, perm = oracle.sort_keys(keys)
    assert sk.tolist() == [0, 1, 1, 5, 5, 5]
    assert perm.tolist() == [4, 1, 3, 0, 2, 5]


def test_tree_contract():
    P = workloads.plummer(5000)
    o = sorted_system(P)
    t = o["tree"]
    n = t.num_nodes
    assert t.depth[0] == 0 and t.leaf_count[0] == 5000 and t.leaf_index[0] == 0
    for i in range(n):
        if t.has_children[i]:
            assert t.leaf_count[i] > 8
            kids = i + t.child_off[i, :8]
            assert kids[0] == i + 1                                   # DFS pre-order
            assert t.leaf_count[kids].sum() == t.leaf_count[i]        # all 8 children exist, ranges tile the parent
            assert np.all(t.depth[kids] == t.depth[i] + 1)
            assert np.all(t.sibling[kids] == np.arange(8))
            assert np.all(kids + t.parent_off[kids] == i)
            nxt = np.append(kids[1:], i + t.child_off[i, 8])
            assert np.all(kids + t.child_off[kids, 8] == nxt)         # child_indices[8] = next sibling
        else:
            assert t.leaf_count[i] <= 8 or t.depth[i] == 21
            assert t.child_off[i, 8] == 1
    # every particle sits inside its leaf's cell
    leaves = np.where((t.has_children == 0) & (t.leaf_count > 0))[0]
    for i in leaves[:200]:
        p = o["posq"][t.leaf_index[i]:t.leaf_index[i] + t.leaf_count[i], :3]
        half = t.geom[i, 3] / 2
        assert np.all(np.abs(p - t.geom[i, :3]) <= half * (1 + 1e-6))


def test_tree_edge_cases():
    # fewer particles than the capacity: a childless root
    o = sorted_system(workloads.uniform_cube(5))
    assert o["tree"].num_nodes == 1 and o["tree"].has_children[0] == 0
    m2l, p2p = o["tree"].traverse(0.5)
    assert len(m2l) == 0 and p2p.tolist() == [[0, 0]]
    # coincident particles: the chain stops at max depth with an over-full leaf
    P = np.zeros((20, 12), np.float32)
    P[:, 0:3] = 0.3; P[:, 8] = 1; P[:, 9] = 1
    o = sorted_system(P)
    t = o["tree"]
    assert t.depth.max() == 21 and t.leaf_count[t.depth == 21].max() == 20
    o2 = sorted_system(P, max_depth=5)
    assert o2["tree"].depth.max() == 5


def test_traversal_covers_every_pair_exactly_once():
    P = workloads.uniform_cube(600)
    o = sorted_system(P)
    t = o["tree"]
    m2l, p2p = t.traverse(0.5)
    n = 600
    cover = np.zeros((n, n), np.int32)
    for lst in (m2l, p2p):
        for a, b in lst:
            ra = slice(t.leaf_index[a], t.leaf_index[a] + t.leaf_count[a])
            rb = slice(t.leaf_index[b], t.leaf_index[b] + t.leaf_count[b])
            cover[ra, rb] += 1
            if a != b:
                cover[rb, ra] += 1
    assert np.all(cover == 1)  # each ordered particle pair (incl. i == j inside self leaves) is handled exactly once


@pytest.mark.parametrize("kind,n,order,tol", [("uniform", 3000, 4, 5e-4), ("plummer", 3000, 4, 1e-3), ("uniform", 3000, 2, 2e-2)])
def test_fmm_oracle_converges_to_direct(kind, n, order, tol):
    P = workloads.GENERATORS[kind](n)
    o = sorted_system(P)
    o["tree"].traverse(0.5)
    g = o["tree"].fmm_field(o["posq"], order, 0.01)
    gd = oracle.direct_field(o["posq"], None, 0.01)
    assert rms_rel(g, gd) < tol


def test_reference_order0_far_field_is_inaccurate():
    # documents SURVEY D8: the reference's monopole-at-cell-centre far field cannot meet 1e-3
    P = workloads.uniform_cube(3000)
    o = sorted_system(P)
    o["tree"].traverse(0.5)
    g = o["tree"].fmm_field(o["posq"], 0, 0.01)
    gd = oracle.direct_field(o["posq"], None, 0.01)
    assert rms_rel(g, gd) > 0.1


def test_direct_step_integrators():
    P = workloads.uniform_cube(64)
    G = workloads.force_constant("uniform", 64)
    a, t = oracle.direct_step(P, G, 0.01, 1e-3, 1, 0)
    b, _ = oracle.direct_step(P, G, 0.01, 1e-3, 1, 1)
    assert np.float32(t) == np.float32(1e-3)
    np.testing.assert_array_equal(a[:, 4:7], b[:, 4:7])                       # same kick
    np.testing.assert_allclose(a[:, 0:3], P[:, 0:3] + a[:, 4:7] * np.float32(1e-3), rtol=0, atol=1e-7)   # drift with v_new
    np.testing.assert_allclose(b[:, 0:3], P[:, 0:3] + P[:, 4:7] * np.float32(1e-3), rtol=0, atol=1e-7)   # drift with v_old
