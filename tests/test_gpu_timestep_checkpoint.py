"""GPU tests of SURVEY 8f ranks 2 and 4 through the C ABI: the variable time step against the oracle's FP64 direct-sum
restatement of the same rule, caller-driven step changes, and checkpoint -> restart (bit-exact where the arithmetic is
deterministic, i.e. the all-pairs path; to FP32 round-off on the FMM path, whose list order depends on kernel timing)."""
import numpy as np
import pytest

import nbody_b200
import oracle
from nbody_b200 import workloads

pytestmark = pytest.mark.gpu


def by_identity(sim):
    """state in the order of the constructor's array"""
    P = sim.particles()
    out = np.empty_like(P)
    out[sim.permutation()] = P
    return out


@pytest.mark.parametrize("flags,tol_dt,tol_x", [(nbody_b200.FLAG_DIRECT, 2e-5, 2e-6), (0, 1e-3, 2e-5)])
def test_variable_time_step_follows_the_oracle(flags, tol_dt, tol_x):
    n, steps, eta = 3000, 6, 0.02
    P = workloads.uniform_cube(n)
    G = workloads.force_constant("uniform", n)
    Pref, tref, dts, amax = oracle.direct_step_adaptive(P, G, 0.01, 1e-3, eta, 0.0, 0.0, steps)
    assert np.all(dts[1:] < 1e-3)                       # the rule is active from the second step on
    sim = nbody_b200.CudaSimulation([1.0, 1.0, 1.0], P, 1e-3, force_constant=G, flags=flags, time_step_eta=eta)
    assert sim.time_step() == {"next": pytest.approx(1e-3), "last": 0.0, "acc_max": 0.0}
    got_dt, got_amax, t = [], [], 0.0
    for _ in range(steps):
        t = sim.step()
        ts = sim.time_step()
        got_dt.append(ts["last"]); got_amax.append(ts["acc_max"])
        # the library's next step is the host rule applied to the maximum it reports, bit for bit
        assert np.float32(ts["next"]) == np.float32(nbody_b200.next_time_step(ts["acc_max"], config=sim.config))
    np.testing.assert_allclose(got_amax, amax, rtol=2 * tol_dt)
    np.testing.assert_allclose(got_dt, dts, rtol=tol_dt)
    acc = np.float32(0)
    for d in got_dt:
        acc = np.float32(acc + np.float32(d))
    assert np.float32(t) == acc and abs(t - tref) < steps * 1e-3 * tol_dt + 1e-7   # FP32 accumulation of the steps actually taken
    assert sim.sim_time() == (pytest.approx(t), steps)
    assert np.abs(by_identity(sim)[:, 0:3] - Pref[:, 0:3]).max() < tol_x
    # max |a| reported == max over the accelerations the caller can read
    a = sim.accelerations().astype(np.float64)
    assert abs(np.sqrt((a * a).sum(1).max()) / got_amax[-1] - 1) < 1e-6
    sim.close()


def test_time_step_bounds_and_fixed_step_is_untouched():
    n = 2000
    P = workloads.plummer(n)
    sim = nbody_b200.CudaSimulation([1.0, 1.0, 1.0], P, 1e-3, time_step_eta=1e-4, time_step_min=4e-4, time_step_max=8e-4)
    sim.step()
    assert sim.time_step()["last"] == np.float32(1e-3) and sim.time_step()["next"] == np.float32(4e-4)   # clamped from below
    sim.close()
    sim = nbody_b200.CudaSimulation([1.0, 1.0, 1.0], P, 1e-3, time_step_eta=100.0, time_step_max=8e-4)
    sim.step()
    assert sim.time_step()["next"] == np.float32(8e-4)                                                  # clamped from above
    sim.close()
    sim = nbody_b200.CudaSimulation([1.0, 1.0, 1.0], P, 1e-3)                                           # eta = 0: the reference's fixed step
    for k in range(3):
        t = sim.step()
    ts = sim.time_step()
    assert ts == {"next": np.float32(1e-3), "last": np.float32(1e-3), "acc_max": 0.0} and abs(t - 0.003) < 1e-7
    sim.close()


def test_caller_driven_time_steps_match_the_oracle():
    n = 2048
    P = workloads.uniform_cube(n)
    G = workloads.force_constant("uniform", n)
    sim = nbody_b200.CudaSimulation([1.0, 1.0, 1.0], P, 1e-3, force_constant=G, flags=nbody_b200.FLAG_DIRECT)
    Pref, tsum = P.copy(), np.float32(0)
    for dt in (1e-3, 2.5e-4, 2e-3, 2e-3):
        sim.set_time_step(dt)
        t = sim.step()
        Pref, _ = oracle.direct_step(Pref, G, 0.01, dt, 1, 0)
        tsum = np.float32(tsum + np.float32(dt))
        assert np.float32(t) == tsum and sim.time_step()["last"] == np.float32(dt)
    assert np.abs(by_identity(sim)[:, 0:7] - Pref[:, 0:7]).max() < 5e-6
    with pytest.raises(nbody_b200.NbodyCudaError):
        sim.set_time_step(0.0)
    with pytest.raises(nbody_b200.NbodyCudaError):
        sim.set_time_step(float("nan"))
    sim.close()


def test_checkpoint_restart_is_bit_exact_on_the_deterministic_path(tmp_path):
    n = 4096
    P = workloads.uniform_cube(n)
    G = workloads.force_constant("uniform", n)
    kw = dict(force_constant=G, flags=nbody_b200.FLAG_DIRECT, time_step_eta=0.02)
    full = nbody_b200.CudaSimulation([1.0, 1.0, 1.0], P, 1e-3, **kw)
    for _ in range(5):
        t_full = full.step()
    first = nbody_b200.CudaSimulation([1.0, 1.0, 1.0], P, 1e-3, **kw)
    for _ in range(2):
        first.step()
    path = str(tmp_path / "run.ckp")
    first.save_checkpoint(path)
    saved_state, saved_perm, saved_ts = first.particles(), first.permutation(), first.time_step()
    first.close()
    hdr, Q, o = nbody_b200.checkpoint_read(path)                  # the file holds exactly what the ABI returned
    assert np.array_equal(Q, saved_state) and np.array_equal(o, saved_perm) and hdr.steps_done == 2
    assert hdr.next_time_step == np.float32(saved_ts["next"]) and hdr.last_acc_max == np.float32(saved_ts["acc_max"])
    second = nbody_b200.CudaSimulation.from_checkpoint(path)
    assert second.n == n and second.sim_time() == (pytest.approx(hdr.time), 2) and second.time_step() == saved_ts
    assert np.array_equal(second.particles(), saved_state) and np.array_equal(second.permutation(), saved_perm)
    assert second.config.flags == nbody_b200.FLAG_DIRECT and second.config.time_step_eta == np.float32(0.02)
    for _ in range(3):
        t_second = second.step()
    assert np.float32(t_second) == np.float32(t_full) and second.sim_time()[1] == 5
    assert np.array_equal(second.particles(), full.particles())   # bit for bit, velocities included
    assert np.array_equal(second.permutation(), full.permutation())
    assert second.time_step() == full.time_step()
    full.close(); second.close()


def test_checkpoint_restart_on_the_fmm_path_and_with_a_new_configuration(tmp_path):
    n = 20000
    P = workloads.plummer(n)
    full = nbody_b200.CudaSimulation([1.0, 1.0, 1.0], P, 1e-3, leaf_capacity=16)
    for _ in range(4):
        full.step()
    first = nbody_b200.CudaSimulation([1.0, 1.0, 1.0], P, 1e-3, leaf_capacity=16)
    first.step(); first.step()
    path = str(tmp_path / "fmm.ckp")
    first.save_checkpoint(path)
    first.close()
    second = nbody_b200.CudaSimulation.from_checkpoint(path)
    assert second.config.leaf_capacity == 16
    with pytest.raises(nbody_b200.NbodyCudaError):               # no step taken by this object yet: nothing to export
        second.keys()
    second.step(); t = second.step()
    assert abs(t - 0.004) < 1e-7 and second.sim_time()[1] == 4
    a, b = by_identity(full), by_identity(second)
    kick = np.abs(a[:, 4:7] - P[:, 4:7]).max()
    assert np.abs(a[:, 0:3] - b[:, 0:3]).max() < 1e-6 and np.abs(a[:, 4:7] - b[:, 4:7]).max() < 1e-4 * kick
    assert np.array_equal(a[:, 8:10], b[:, 8:10])                # masses and charges travel unchanged
    second.close(); full.close()
    # continue the same file under another configuration (order 3, capacity 8, half the step)
    third = nbody_b200.CudaSimulation.from_checkpoint(path, order=3, leaf_capacity=8, time_step=5e-4)
    assert third.config.order == 3 and third.config.leaf_capacity == 8 and third.time_step()["next"] == np.float32(5e-4)
    t = third.step()
    assert abs(t - 0.0025) < 1e-7 and np.isfinite(third.particles()).all()
    third.close()
