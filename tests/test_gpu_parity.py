"""GPU parity tests (B200): every stage of the CUDA path, called through the C ABI,
against the CPU oracle on the same seeded inputs.

Bars: bit-exact for integer work (Morton keys, permutation, octree topology,
interaction lists); multipoles / locals within FP32 round-off of the FP64 oracle;
accelerations within 1e-3 RMS of FP64 direct summation (BASELINE.json north_star)."""
import os
from math import factorial

import numpy as np
import pytest

import nbody_b200
import oracle
from nbody_b200 import workloads
from conftest import sorted_system, rms_rel

pytestmark = pytest.mark.gpu

ACC_TOL = 1e-3   # RMS relative acceleration error vs FP64 direct sum (north_star)
EXP_TOL = 2e-6   # multipole / local coefficients vs the FP64 oracle (FP32 round-off)


def make_sim(P, **kw):
    kw.setdefault("flags", nbody_b200.FLAG_NO_INTEGRATE)
    return nbody_b200.CudaSimulation([1.0, 1.0, 1.0], P, 1e-3, **kw)


def directed(unordered):
    u = unordered.astype(np.uint64)
    d = np.concatenate([u, u[:, ::-1]], axis=0)
    return np.unique(d[:, 0] << np.uint64(32) | d[:, 1])


def packed(pairs):
    p = pairs.astype(np.uint64)
    return np.sort(p[:, 0] << np.uint64(32) | p[:, 1])


CASES = [("uniform", 4096, 8), ("uniform", 30000, 8), ("plummer", 30000, 8), ("two_galaxies", 20000, 8),
         ("plummer", 20000, 32), ("uniform", 777, 3)]


@pytest.mark.parametrize("kind,n,cap", CASES)
def test_keys_tree_lists_bit_exact(kind, n, cap):
    P = workloads.GENERATORS[kind](n)
    sim = make_sim(P, leaf_capacity=cap)
    sim.step()
    o = sorted_system(P, capacity=cap)
    assert np.array_equal(sim.keys(), o["keys"])
    assert np.array_equal(sim.permutation(), o["perm"])          # stable sort: same order among equal keys
    t, ref = sim.tree(), o["tree"]
    assert len(t["depth"]) == ref.num_nodes
    for name in ("depth", "prefix", "leaf_index", "leaf_count", "has_children", "child_off", "parent_off", "sibling", "geom"):
        assert np.array_equal(t[name], getattr(ref, name)), name
    m2l_o, p2p_o = ref.traverse(0.5)
    m2l, p2p = sim.lists()
    assert np.array_equal(packed(m2l), directed(m2l_o))
    assert np.array_equal(packed(p2p), directed(p2p_o))
    st = sim.stats()
    assert st["m2l_interactions"] == 2 * len(m2l_o)
    cnt = ref.leaf_count.astype(np.int64)
    evals = (cnt[p2p[:, 0]] * cnt[p2p[:, 1]]).sum() - n          # ordered target<-source evaluations, i != j
    assert st["p2p_interactions"] == evals
    sim.close()


@pytest.mark.parametrize("kind,n,order,tau", [("uniform", 20000, 4, None), ("plummer", 20000, 4, None), ("plummer", 20000, 4, 0.0),
                                               ("uniform", 20000, 4, 0.0), ("plummer", 20000, 3, None), ("plummer", 20000, 3, 0.0),
                                               ("uniform", 20000, 2, None), ("plummer", 20000, 5, 0.0), ("uniform", 20000, 5, None)])
def test_expansions_and_accelerations(kind, n, order, tau):
    """tau = 0.0 (and order 2, which has no lower order): every pair at order P — the device must reproduce the oracle's
    expansions coefficient by coefficient. tau = None: the default adaptive-order M2L (pairs with ext2 < 0.13 d2 may run at
    order P-1). The traversal classifies exactly like the oracle (m2l_interactions_low), but the M2L kernel evaluates some
    low-class pairs at order P where lanes would otherwise idle, so its result must lie between the oracle's adaptive and
    full-order answers: never further from the full-order one than the adaptive oracle is."""
    P = workloads.GENERATORS[kind](n)
    sim = make_sim(P, order=order) if tau is None else make_sim(P, order=order, low_order_tau=tau)
    sim.step()
    o = sorted_system(P)
    tr = o["tree"]
    tr.traverse(0.5)
    tau_eff = float(sim.config.low_order_tau) if order >= 3 else 0.0
    g_fmm, Mo, Lo = tr.fmm_field(o["posq"], order, 0.01, want_expansions=True, low_order_tau=tau_eff)
    assert sim.stats()["m2l_interactions_low"] == tr.low_count
    if tau is None and order >= 3:
        assert 0.4 < tr.low_fraction < 0.9
    M, L = sim.expansions()
    ne = tr.leaf_count > 0
    assert rms_rel(M[ne], Mo[ne]) < EXP_TOL
    mi = [(i, j, oo - i - j) for oo in range(order + 1) for i in range(oo, -1, -1) for j in range(oo - i, -1, -1)]
    fac = np.array([factorial(i) * factorial(j) * factorial(k) for i, j, k in mi], np.float64)
    # the device keeps pure derivatives (n! x Taylor coefficient) and does not carry order 0
    scale = (o["P"][:, 9] / o["P"][:, 8])[:, None]
    acc = sim.accelerations()
    tg = np.linspace(0, n - 1, 4096).astype(np.uint32)
    gd = oracle.direct_field(o["posq"], tg, 0.01)
    err = rms_rel(acc[tg], gd * scale[tg])
    if tau_eff == 0.0:
        assert rms_rel(L[ne][:, 1:], (Lo * fac[None, :])[ne][:, 1:]) < EXP_TOL
        assert rms_rel(acc, g_fmm * scale) < 2e-6                     # same algorithm, FP32 vs FP64
        assert abs(err - rms_rel(g_fmm[tg], gd)) < 1e-5               # the GPU adds no error beyond the method's
    else:
        g_full, _, Lf = tr.fmm_field(o["posq"], order, 0.01, want_expansions=True, low_order_tau=0.0)
        Lf_d, La_d = (Lf * fac[None, :])[ne][:, 1:], (Lo * fac[None, :])[ne][:, 1:]
        assert rms_rel(L[ne][:, 1:], Lf_d) <= 1.05 * rms_rel(La_d, Lf_d) + EXP_TOL
        assert rms_rel(acc, g_full * scale) <= 1.05 * rms_rel(g_fmm, g_full) + 2e-6
        assert err <= rms_rel(g_fmm[tg], gd) + 1e-5                   # no error beyond the adaptive method's
    if order >= 4:
        assert err < ACC_TOL
    if order == 5 and tau_eff == 0.0:
        assert err < 1e-4                                             # one order more: 2.1e-4 -> below 1e-4 on the Plummer model
    sim.close()


def test_config1_uniform_4096_ten_steps_vs_direct_sum():
    """BASELINE config 1: uniform cube N=4096, 10 kick-drift steps, against the FP64 direct-sum oracle."""
    n = 4096
    P = workloads.uniform_cube(n)
    G = workloads.force_constant("uniform", n)
    sim = nbody_b200.CudaSimulation([1.0, 1.0, 1.0], P, 1e-3, force_constant=G)
    t = 0.0
    for _ in range(10):
        t = sim.step()
    out, perm = sim.particles(), sim.permutation()
    ref, tref = oracle.direct_step(P, G, 0.01, 1e-3, 10, 0)
    assert np.float32(t) == np.float32(tref)
    assert sorted(perm.tolist()) == list(range(n))
    dv_ref = ref[perm][:, 4:7] - P[perm][:, 4:7]
    assert rms_rel(out[:, 4:7] - P[perm][:, 4:7], dv_ref) < ACC_TOL      # accumulated kicks
    np.testing.assert_allclose(out[:, 0:3], ref[perm][:, 0:3], rtol=0, atol=5e-6)
    assert np.array_equal(out[:, 8:10], P[perm][:, 8:10])                   # mass / charge carried unchanged
    sim.close()


def test_explicit_euler_variant():
    n = 2000
    P = workloads.uniform_cube(n)
    G = workloads.force_constant("uniform", n)
    sim = nbody_b200.CudaSimulation([1.0, 1.0, 1.0], P, 1e-3, force_constant=G, integrator=nbody_b200.EXPLICIT_EULER)
    sim.step()
    out, perm = sim.particles(), sim.permutation()
    # x += v_old dt  (src/open_cl_simulation.cpp:602-607)
    np.testing.assert_allclose(out[:, 0:3], P[perm][:, 0:3] + P[perm][:, 4:7] * np.float32(1e-3), rtol=0, atol=1e-7)
    ref, _ = oracle.direct_step(P, G, 0.01, 1e-3, 1, 1)
    assert rms_rel(out[:, 4:7] - P[perm][:, 4:7], ref[perm][:, 4:7] - P[perm][:, 4:7]) < ACC_TOL
    sim.close()


def test_direct_mode_and_direct_field_kernel():
    n = 5000
    P = workloads.plummer(n)
    sim = make_sim(P, flags=nbody_b200.FLAG_NO_INTEGRATE | nbody_b200.FLAG_DIRECT)
    sim.step()
    o = sorted_system(P)
    gd = oracle.direct_field(o["posq"], None, 0.01)
    scale = (o["P"][:, 9] / o["P"][:, 8])[:, None]
    assert rms_rel(sim.accelerations(), gd * scale) < 5e-6
    sim.close()
    f, ms = nbody_b200.direct_field(o["posq"], o["posq"][:1000], 0.01)
    assert rms_rel(f, gd[:1000]) < 5e-6 and ms > 0
    # no softening: coincident points and the self term must not produce NaN
    posq = o["posq"].copy(); posq[1, :3] = posq[0, :3]
    f0, _ = nbody_b200.direct_field(posq, posq[:64], 0.0)
    assert np.all(np.isfinite(f0))
    assert rms_rel(f0, oracle.direct_field(posq, np.arange(64, dtype=np.uint32), 0.0)) < 2e-5


def test_edge_cases():
    # N = 1 and N below the capacity: childless root, P2P only
    for n in (1, 5, 8, 9):
        P = workloads.uniform_cube(n)
        sim = make_sim(P)
        sim.step()
        o = sorted_system(P)
        assert len(sim.tree()["depth"]) == o["tree"].num_nodes
        gd = oracle.direct_field(o["posq"], None, 0.01)
        scale = (o["P"][:, 9] / o["P"][:, 8])[:, None]
        if n > 1:
            assert rms_rel(sim.accelerations(), gd * scale) < 1e-5
        else:
            assert np.all(sim.accelerations() == 0)
        sim.close()
    # coincident particles: an over-full leaf at max depth (or at a reduced max_depth)
    P = workloads.uniform_cube(300)
    P[:40, 0:3] = (0.3, 0.3, 0.3)
    for md in (21, 6):
        sim = make_sim(P, max_depth=md, low_order_tau=0.0)  # every pair at order P: the device must equal the FP64 FMM
        sim.step()
        o = sorted_system(P, max_depth=md)
        assert np.array_equal(sim.keys(), o["keys"])
        t = sim.tree()
        assert np.array_equal(t["leaf_count"], o["tree"].leaf_count) and np.array_equal(t["depth"], o["tree"].depth)
        m2l_o, p2p_o = o["tree"].traverse(0.5)
        m2l, p2p = sim.lists()
        assert np.array_equal(packed(m2l), directed(m2l_o)) and np.array_equal(packed(p2p), directed(p2p_o))
        gd = oracle.direct_field(o["posq"], None, 0.01)
        scale = (o["P"][:, 9] / o["P"][:, 8])[:, None]
        g_fmm = o["tree"].fmm_field(o["posq"], 4, 0.01, low_order_tau=0.0)
        assert rms_rel(sim.accelerations(), g_fmm * scale) < 2e-6      # identical to the FP64 FMM over the same lists
        assert rms_rel(sim.accelerations(), gd * scale) < 6e-3         # a 300-body system with a 40-body point clump
        sim.close()
    # particles outside the root box are clamped into the boundary cells: keys, tree and lists still match the
    # oracle bit for bit (the expansions' error bound needs particles inside their cell, so no accuracy claim here)
    P = workloads.uniform_cube(300)
    P[40:44, 0:3] = [(-0.5, 0.2, 0.2), (1.5, 0.2, 0.2), (0.2, 7.0, 0.2), (0.2, 0.2, -3.0)]
    sim = make_sim(P)
    sim.step()
    o = sorted_system(P)
    assert np.array_equal(sim.keys(), o["keys"])
    assert np.array_equal(sim.tree()["child_off"], o["tree"].child_off)
    m2l_o, p2p_o = o["tree"].traverse(0.5)
    m2l, p2p = sim.lists()
    assert np.array_equal(packed(m2l), directed(m2l_o)) and np.array_equal(packed(p2p), directed(p2p_o))
    assert np.all(np.isfinite(sim.accelerations()))
    sim.close()
    # a leaf far beyond the P2P tile size (600 coincident particles): one source entry straddles several tiles
    P = workloads.uniform_cube(2000)
    P[:600, 0:3] = (0.7, 0.2, 0.4)
    sim = make_sim(P, low_order_tau=0.0)
    sim.step()
    o = sorted_system(P)
    gd = oracle.direct_field(o["posq"], None, 0.01)
    scale = (o["P"][:, 9] / o["P"][:, 8])[:, None]
    o["tree"].traverse(0.5)
    g_fmm = o["tree"].fmm_field(o["posq"], 4, 0.01, low_order_tau=0.0)
    assert rms_rel(sim.accelerations(), g_fmm * scale) < 2e-6          # straddling entries == FP64 FMM over the same lists
    assert rms_rel(sim.accelerations(), gd * scale) < 6e-3             # 30 % of the mass in one point: the method's own error
    sim.close()


def test_non_cubic_bounds_and_mac_ratio():
    n = 6000
    P = workloads.uniform_cube(n)
    P[:, 1] *= 0.5; P[:, 2] *= 0.75
    bounds = (1.0, 0.5, 0.75)
    sim = nbody_b200.CudaSimulation(list(bounds), P, 1e-3, flags=nbody_b200.FLAG_NO_INTEGRATE, mac_ratio=0.4)
    sim.step()
    o = sorted_system(P, bounds=bounds)
    assert np.array_equal(sim.keys(), o["keys"])
    t = sim.tree()
    assert np.array_equal(t["geom"], o["tree"].geom) and np.array_equal(t["child_off"], o["tree"].child_off)
    m2l_o, p2p_o = o["tree"].traverse(0.4)
    m2l, p2p = sim.lists()
    assert np.array_equal(packed(m2l), directed(m2l_o)) and np.array_equal(packed(p2p), directed(p2p_o))
    sim.close()


def test_set_particles_round_trip_and_errors():
    n = 3000
    P = workloads.plummer(n)
    sim = nbody_b200.CudaSimulation([1.0, 1.0, 1.0], P, 1e-3)
    assert np.array_equal(sim.particles()[:, :10], P[:, :10] * np.array([1, 1, 1, 0, 1, 1, 1, 0, 1, 1], np.float32))
    sim.step()
    a = sim.particles()
    sim.set_particles(P)
    sim.step()
    b = sim.particles()
    np.testing.assert_allclose(a, b, rtol=0, atol=1e-6)               # same input -> same result (up to atomic order)
    with pytest.raises(nbody_b200.NbodyCudaError):
        sim.set_particles(P[:10])
    sim.close()
    sim = nbody_b200.CudaSimulation([1.0, 1.0, 1.0], P, 1e-3)
    with pytest.raises(nbody_b200.NbodyCudaError):
        sim.keys()                                                      # no step yet
    sim.close()


def test_pool_growth_retry():
    # deliberately tiny pools: the step must grow them and still produce the right answer
    n = 20000
    P = workloads.plummer(n)
    sim = make_sim(P, pool_scale=0.05)
    sim.step()
    assert sim.stats()["retries"] >= 1
    o = sorted_system(P)
    tg = np.linspace(0, n - 1, 2048).astype(np.uint32)
    gd = oracle.direct_field(o["posq"], tg, 0.01)
    scale = (o["P"][:, 9] / o["P"][:, 8])[:, None]
    assert rms_rel(sim.accelerations()[tg], gd * scale[tg]) < ACC_TOL
    sim.close()


def test_large_uniform_1m_against_gpu_direct_subsample():
    """BASELINE config 2: uniform cube N=2^20, validated against direct summation on a subsample
    (GPU all-pairs kernel on 65,536 targets, FP64 CPU oracle on 512 of them)."""
    n = 1 << 20
    P = workloads.uniform_cube(n)
    sim = make_sim(P)
    sim.step()
    acc = sim.accelerations()
    out = sim.particles()
    posq = np.ascontiguousarray(np.concatenate([out[:, 0:3], out[:, 9:10]], axis=1))
    tg = np.linspace(0, n - 1, 65536).astype(np.int64)
    f, _ = nbody_b200.direct_field(posq, posq[tg], 0.01)
    scale = (out[:, 9] / out[:, 8])[:, None]
    assert rms_rel(acc[tg], f * scale[tg]) < ACC_TOL
    spot = tg[::128].astype(np.uint32)
    gd = oracle.direct_field(posq, spot, 0.01)
    assert rms_rel(f[::128], gd) < 2e-4                                   # FP32 accumulation over 2^20 sources
    assert rms_rel(acc[spot], gd * scale[spot]) < ACC_TOL
    # size-independent properties: sorted keys, a permutation, momentum conservation of the pair forces
    k = sim.keys()
    assert np.all(k[1:] >= k[:-1])
    assert np.array_equal(np.sort(sim.permutation()), np.arange(n, dtype=np.uint32))
    F = acc.astype(np.float64) * out[:, 8:9]                              # force = m a (G = 1, q = m)
    assert np.abs(F.sum(0)).max() / np.abs(F).sum(0).max() < 1e-4
    sim.close()


def test_full_size_plummer_16m_properties():
    """BASELINE config 3 at full size (Plummer N = 2^24) AT THE BENCHED SETTINGS (bench.py: leaf capacity 48, order 4, default
    low_order_tau) through size-independent properties: sorted keys, a valid permutation, accelerations within 1e-3 RMS of direct
    summation on a 65,536-target subsample (GPU all-pairs kernel with compensated sums, cross-checked against the FP64 oracle on
    128 targets), and vanishing net force."""
    n = 1 << 24
    P = workloads.plummer(n)
    sim = make_sim(P, leaf_capacity=48, order=4)
    assert abs(sim.config.low_order_tau - 0.13) < 1e-6
    sim.step()
    st = sim.stats()
    assert st["retries"] == 0 and st["n_leaves"] > 0 and st["m2l_interactions"] > st["m2l_interactions_low"] > 0
    k = sim.keys()
    assert np.all(k[1:] >= k[:-1])
    perm = sim.permutation()
    assert np.array_equal(np.sort(perm), np.arange(n, dtype=np.uint32))
    out = sim.particles()
    assert np.array_equal(out[:, 0:3], P[perm][:, 0:3])                   # NO_INTEGRATE: the state is only re-ordered
    acc = sim.accelerations()
    sim.close()
    posq = np.ascontiguousarray(np.concatenate([out[:, 0:3], out[:, 9:10]], axis=1))
    tg = np.linspace(0, n - 1, 65536).astype(np.int64)
    f, _ = nbody_b200.direct_field(posq, posq[tg], 0.01)
    scale = (out[:, 9] / out[:, 8])[:, None]
    assert rms_rel(acc[tg], f * scale[tg]) < ACC_TOL
    spot = tg[::512].astype(np.uint32)
    gd = oracle.direct_field(posq, spot, 0.01)
    assert rms_rel(f[::512], gd) < 2e-5
    assert rms_rel(acc[spot], gd * scale[spot]) < ACC_TOL
    F = acc.astype(np.float64) * out[:, 8:9]                              # equal masses: total force must cancel
    assert np.abs(F.sum(0)).max() / np.abs(F).sum(0).max() < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("n,nruns", [(1, 1), (5000, 2), (70001, 3), (300000, 8), (1 << 21, 16)])  # (first green on a B200: profiles/r02a_call.log)
def test_distributed_sort_pipeline_equals_stable_sort(n, nruns):
    """nbody_cuda_sort_runs = what NBODY_FLAG_DIST_SORT does on the device (slice radix sorts + pairwise merge rounds), on one GPU:
    the stable sort of all keys for any boundaries — the oracle's std::stable_sort, ties included."""
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 1 << 62, n, dtype=np.uint64)
    keys[rng.integers(0, n, n // 3)] = keys[0]                                       # heavy ties
    bound = np.concatenate([[0], np.sort(rng.integers(0, n + 1, nruns - 1)), [n]])  # includes empty runs
    k, i = nbody_b200.sort_runs(keys, bound)
    sk, perm = oracle.sort_keys(keys)
    assert np.array_equal(k, sk) and np.array_equal(i.astype(np.int64), np.asarray(perm, np.int64))

