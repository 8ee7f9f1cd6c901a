"""Partitioned multi-GPU scheme (NBODY_FLAG_PARTITIONED: every rank holds only its Morton-key range and imports a locally essential
tree, nbody_b200/csrc/let.cu) on ONE GPU through virtual ranks: the same kernels and phases as one process per GPU, the exchanges
are device copies. The W-rank run must reproduce the single-GPU run: Morton keys, tree order and permutation bit for bit (the
concatenation of the ranks in rank order IS the global tree order), the P2P work summed over the ranks exactly (the ranks' trees are
the global octree restricted to their own particles, so the interaction lists are the global ones), accelerations and trajectories
to FP32 round-off (the lists are summed in a different order), and the accelerations within 1e-3 RMS of FP64 direct summation."""
import numpy as np
import pytest

import nbody_b200
import oracle
from nbody_b200 import workloads
from conftest import rms_rel

pytestmark = pytest.mark.gpu

ACC_TOL = 1e-3


def single(P, **kw):
    return nbody_b200.CudaSimulation([1.0, 1.0, 1.0], P, 1e-3, **kw)


CASES = [("plummer", 30000, 8, 2), ("plummer", 30000, 8, 4), ("uniform", 20000, 8, 8), ("two_galaxies", 40000, 32, 4),
         ("plummer", 100000, 48, 8), ("uniform", 5000, 8, 3), ("plummer", 3000, 8, 16)]


@pytest.mark.parametrize("kind,n,cap,world", CASES)
def test_force_evaluation_matches_single_gpu(kind, n, cap, world):
    P = workloads.GENERATORS[kind](n)
    ref = single(P, leaf_capacity=cap, flags=nbody_b200.FLAG_NO_INTEGRATE)
    ref.step()
    grp = nbody_b200.VirtualGroup([1.0, 1.0, 1.0], P, 1e-3, world, leaf_capacity=cap, flags=nbody_b200.FLAG_NO_INTEGRATE)
    grp.step()
    assert sum(grp.counts()) == n
    assert np.array_equal(grp.keys(), ref.keys())
    assert np.array_equal(grp.permutation(), ref.permutation())
    assert np.array_equal(grp.particles(), ref.particles())
    st, sr = grp.stats(), ref.stats()
    assert sum(s["p2p_interactions"] for s in st) == sr["p2p_interactions"]       # same P2P lists, particle for particle
    assert sum(s["m2l_interactions"] for s in st) >= sr["m2l_interactions"]       # cells shared by several ranks count once per part
    a, a_ref = grp.accelerations().astype(np.float64), ref.accelerations().astype(np.float64)
    assert rms_rel(a, a_ref) < 1.5e-4                                             # summation order only (FP32, thousands of terms that largely cancel)
    if n <= 40000:
        out = ref.particles()
        posq = np.ascontiguousarray(np.concatenate([out[:, 0:3], out[:, 9:10]], axis=1))
        g = oracle.direct_field(posq, None, 0.01) * (out[:, 9] / out[:, 8])[:, None]
        e_grp, e_ref = rms_rel(a, g), rms_rel(a_ref, g)
        assert e_grp < ACC_TOL and abs(e_grp - e_ref) < 0.05 * e_ref                # the same error against FP64 direct summation
    assert all(s["retries"] == 0 for s in st)
    if world > 1 and n >= 20000:
        assert all(s["imported_nodes"] > 0 for s in st) and any(s["halo_particles"] > 0 for s in st)
    grp.close(); ref.close()


@pytest.mark.parametrize("kind,n,cap,world,flags", [("plummer", 60000, 16, 4, 0), ("plummer", 60000, 16, 4, nbody_b200.FLAG_STATIC_PARTITION),
                                                    ("two_galaxies", 50000, 8, 8, 0), ("uniform", 30000, 8, 2, 0)])
def test_trajectory_matches_single_gpu(kind, n, cap, world, flags):
    """Five steps with the integrator: particles migrate between the ranks and the splitters follow the measured work, yet the
    state after every step is the single-GPU state in the same order."""
    P = workloads.GENERATORS[kind](n)
    G = workloads.force_constant(kind, n)
    ref = single(P, leaf_capacity=cap, force_constant=G)
    grp = nbody_b200.VirtualGroup([1.0, 1.0, 1.0], P, 1e-3, world, leaf_capacity=cap, force_constant=G, flags=flags)
    migrated = 0
    for step in range(5):
        t_ref, t = ref.step(), grp.step()
        assert t == t_ref
        assert sum(grp.counts()) == n
        a, b, pa, pb = grp.particles(), ref.particles(), grp.permutation(), ref.permutation()
        k = grp.keys()
        assert np.all(k[1:] >= k[:-1]), step                       # the ranks' key ranges tile the key space in rank order
        assert np.array_equal(np.sort(pa), np.arange(n, dtype=np.uint32)), step   # nobody lost, nobody duplicated
        if step == 0:  # identical input: identical order and lists; later the two runs differ by FP32 round-off (the ranks sum their
            assert np.array_equal(pa, pb)  # lists in another order), and a particle 1e-7 from a cell boundary may sort differently
            assert sum(s["p2p_interactions"] for s in grp.stats()) == ref.stats()["p2p_interactions"]
        ia, ib = np.argsort(pa), np.argsort(pb)                    # by identity
        assert np.abs(a[ia, 0:3] - b[ib, 0:3]).max() < 2e-6 and np.abs(a[ia, 4:7] - b[ib, 4:7]).max() < 1e-3, step
        rel = abs(sum(s["p2p_interactions"] for s in grp.stats()) / ref.stats()["p2p_interactions"] - 1.0)
        assert rel < 1e-3, step
        migrated += sum(s["migrated_particles"] for s in grp.stats()) if step else 0
    assert migrated > 0                                            # particles did change owner
    if not flags & nbody_b200.FLAG_STATIC_PARTITION:
        assert len(set(grp.counts())) > 1  # the partition follows the work, not the particle count
    grp.close(); ref.close()


def test_owned_round_trip_and_memory_scaling():
    """get/set of a member's own particles touches nothing else, and a member's device memory follows N/W + halo, not N."""
    n = 200000
    P = workloads.plummer(n)
    bytes_w = {}
    for world in (1, 4):
        grp = nbody_b200.VirtualGroup([1.0, 1.0, 1.0], P, 1e-3, world, leaf_capacity=32)
        grp.step()
        bytes_w[world] = max(s["device_bytes"] for s in grp.stats())
        m = grp.members[world - 1]
        mine = m.owned_particles()
        assert mine.shape[0] == m.n and np.array_equal(mine, m.particles())
        first, count = m.owned_range()
        assert count == m.n and first == sum(grp.counts()[:world - 1])
        m.set_owned_particles_ptr(mine.ctypes.data, mine.shape[0])
        assert np.array_equal(m.particles(), mine)
        grp.step()
        grp.close()
    assert bytes_w[4] < 0.6 * bytes_w[1]


def test_pool_growth_inside_a_partitioned_step():
    """Tiny pools: the ranks that overflow grow their pools and repeat their part of the step while the others wait."""
    n = 40000
    P = workloads.plummer(n)
    ref = single(P, leaf_capacity=8, flags=nbody_b200.FLAG_NO_INTEGRATE)
    ref.step()
    grp = nbody_b200.VirtualGroup([1.0, 1.0, 1.0], P, 1e-3, 4, leaf_capacity=8, flags=nbody_b200.FLAG_NO_INTEGRATE, pool_scale=0.02)
    grp.step()
    assert any(s["retries"] > 0 for s in grp.stats())
    assert np.array_equal(grp.permutation(), ref.permutation())
    assert rms_rel(grp.accelerations(), ref.accelerations()) < 1.5e-4
    grp.close(); ref.close()


def test_variable_time_step_and_edge_cases():
    """The variable time step in partitioned mode (the largest acceleration is a maximum over the ranks' slots), a group of one rank,
    more ranks than make sense for the particle count (ranks that own next to nothing), and the calls a member must refuse."""
    n = 20000
    P = workloads.plummer(n)
    kw = dict(leaf_capacity=8, time_step_eta=0.01)
    ref = single(P, **kw)
    grp = nbody_b200.VirtualGroup([1.0, 1.0, 1.0], P, 1e-3, 4, **kw)
    for step in range(3):
        ref.step(); grp.step()
        a, b = grp.members[0].time_step(), ref.time_step()
        assert abs(a["last"] / b["last"] - 1) < 1e-4 and abs(a["next"] / b["next"] - 1) < 1e-3 and abs(a["acc_max"] / b["acc_max"] - 1) < 1e-3, (step, a, b)
        assert all(m.time_step()["next"] == a["next"] for m in grp.members)           # every rank derives the same step, bit for bit
    assert grp.members[0].time_step()["next"] < 1e-3                                  # the rule did take over
    with pytest.raises(nbody_b200.NbodyCudaError):
        grp.members[1].step()                                                        # members step through the group
    grp.close(); ref.close()
    for world, n2 in ((1, 3000), (16, 300)):
        P2 = workloads.uniform_cube(n2)
        r2 = single(P2, flags=nbody_b200.FLAG_NO_INTEGRATE)
        r2.step()
        g2 = nbody_b200.VirtualGroup([1.0, 1.0, 1.0], P2, 1e-3, world, flags=nbody_b200.FLAG_NO_INTEGRATE)
        g2.step()
        assert sum(g2.counts()) == n2 and np.array_equal(g2.permutation(), r2.permutation())
        assert sum(s["p2p_interactions"] for s in g2.stats()) == r2.stats()["p2p_interactions"]
        assert rms_rel(g2.accelerations(), r2.accelerations()) < 1.5e-4
        g2.close(); r2.close()
    with pytest.raises(nbody_b200.NbodyCudaError):
        nbody_b200.VirtualGroup([1.0, 1.0, 1.0], P, 1e-3, 2, flags=nbody_b200.FLAG_DIRECT)   # the all-pairs path is single-GPU only
