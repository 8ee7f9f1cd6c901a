"""CPU tests of the host-side parts of SURVEY 8f ranks 2 and 4: the checkpoint file format (written and read through the
C ABI's host-only entry points and re-parsed here independently with numpy, following the layout documented in
include/nbody_cuda.h) and the variable-time-step rule (product arithmetic == oracle arithmetic, bit for bit).
No device is needed for any of it; restoring a simulation from a file is a GPU test (test_gpu_timestep_checkpoint.py)."""
import ctypes as C
import os
import struct

import numpy as np
import pytest

import nbody_b200
import oracle
from nbody_b200 import workloads

GOLDEN = 0x9E3779B97F4A7C15
MASK = (1 << 64) - 1


def layout_checksum(payload: bytes) -> int:
    """sum_i (w_i + golden) (2i+1) mod 2^64 over the little-endian 64-bit words of the zero-padded payload."""
    if len(payload) % 8:
        payload += b"\0" * (8 - len(payload) % 8)
    w = np.frombuffer(payload, dtype="<u8")
    i = np.arange(w.size, dtype=np.uint64)
    with np.errstate(over="ignore"):
        return int(((w + np.uint64(GOLDEN)) * (np.uint64(2) * i + np.uint64(1))).sum(dtype=np.uint64))


def particles(n, seed=1):
    rng = np.random.default_rng(seed)
    P = rng.random((n, 12), dtype=np.float32)
    P[:, 3] = P[:, 7] = P[:, 10] = P[:, 11] = 0
    return P


@pytest.mark.parametrize("n", [1, 2, 7, 1000])   # odd counts: the permutation does not fill a whole 64-bit word
def test_checkpoint_file_layout_and_round_trip(tmp_path, n):
    P = particles(n)
    orig = np.random.default_rng(2).permutation(n).astype(np.uint32)
    cfg = nbody_b200.default_config(order=3, leaf_capacity=32, time_step=2e-3, time_step_eta=0.05)
    path = str(tmp_path / "state.ckp")
    nbody_b200.checkpoint_write(path, P, orig, time=0.125, steps_done=17, next_time_step=7.5e-4, config=cfg)
    assert not os.path.exists(path + ".partial")
    # the C reader
    hdr, Q, o = nbody_b200.checkpoint_read(path)
    assert np.array_equal(P, Q) and np.array_equal(o, orig)
    assert (hdr.n_particles, hdr.steps_done, hdr.time) == (n, 17, 0.125)
    assert hdr.next_time_step == np.float32(7.5e-4) and hdr.config.order == 3 and hdr.config.leaf_capacity == 32
    assert hdr.config.time_step_eta == np.float32(0.05)
    # an independent reader that follows the documented layout
    raw = open(path, "rb").read()
    hb = C.sizeof(nbody_b200.CheckpointHeader)
    assert hb == 152 and len(raw) == hb + n * 48 + n * 4
    magic, version, header_bytes, n_file, steps = struct.unpack_from("<QIIQQ", raw, 0)
    time_, next_dt, last_dt, last_amax, checksum = struct.unpack_from("<ffffQ", raw, 32)
    assert raw[:8] == b"NBDYCKP1" and magic == nbody_b200.CHECKPOINT_MAGIC and version == 2 and header_bytes == hb
    assert (n_file, steps, time_) == (n, 17, 0.125) and next_dt == np.float32(7.5e-4) and last_dt == 0 and last_amax == 0
    assert np.array_equal(np.frombuffer(raw, "<f4", n * 12, hb).reshape(n, 12), P)
    assert np.array_equal(np.frombuffer(raw, "<u4", n, hb + n * 48), orig)
    # version 2: the sum runs over the header with its checksum field (bytes 48..55) zeroed, then the payload
    assert checksum == layout_checksum(raw[:48] + b"\0" * 8 + raw[56:]) == hdr.checksum
    cfg_file = nbody_b200.Config.from_buffer_copy(raw[56:56 + C.sizeof(nbody_b200.Config)])
    assert cfg_file.order == 3 and cfg_file.time_step == np.float32(2e-3)


def test_version_1_files_are_still_read(tmp_path):
    # version 1 (written by this library before the header came under the checksum): the sum covers the two arrays only
    path = str(tmp_path / "v2.ckp")
    P = particles(50)
    nbody_b200.checkpoint_write(path, P, time=0.5, steps_done=3)
    raw = bytearray(open(path, "rb").read())
    raw[8:12] = struct.pack("<I", 1)
    raw[48:56] = struct.pack("<Q", layout_checksum(bytes(raw[152:])))
    old = str(tmp_path / "v1.ckp")
    open(old, "wb").write(bytes(raw))
    hdr, Q, o = nbody_b200.checkpoint_read(old)
    assert hdr.version == 1 and hdr.time == 0.5 and hdr.steps_done == 3 and np.array_equal(P, Q)
    raw[152 + 7] ^= 0x20
    open(old, "wb").write(bytes(raw))
    with pytest.raises(nbody_b200.NbodyCudaError):
        nbody_b200.checkpoint_read(old)


def test_checkpoint_default_permutation_is_identity(tmp_path):
    path = str(tmp_path / "a.ckp")
    nbody_b200.checkpoint_write(path, particles(33))
    hdr, _, o = nbody_b200.checkpoint_read(path)
    assert np.array_equal(o, np.arange(33, dtype=np.uint32)) and hdr.next_time_step == np.float32(1e-3) and hdr.steps_done == 0


def test_checkpoint_rejects_damaged_files(tmp_path):
    lib = nbody_b200.load_library()
    path = str(tmp_path / "a.ckp")
    nbody_b200.checkpoint_write(path, particles(100))
    raw = bytearray(open(path, "rb").read())

    def expect_error(data, needle):
        bad = str(tmp_path / "bad.ckp")
        open(bad, "wb").write(bytes(data))
        with pytest.raises(nbody_b200.NbodyCudaError) as e:
            nbody_b200.checkpoint_read(bad)
        assert needle in str(e.value), str(e.value)

    flipped = bytearray(raw); flipped[152 + 48 * 40 + 5] ^= 0x10           # one bit of one particle
    expect_error(flipped, "corrupt")
    flipped = bytearray(raw); flipped[-1] ^= 0x01                          # last byte of the permutation
    expect_error(flipped, "corrupt")
    swapped = bytearray(raw); swapped[152:152 + 48], swapped[152 + 48:152 + 96] = raw[152 + 48:152 + 96], raw[152:152 + 48]
    expect_error(swapped, "corrupt")                                        # the checksum is order-sensitive
    for at in (32, 36, 24, 56 + 8):                                         # time, next step, steps done, a configuration field
        flipped = bytearray(raw); flipped[at] ^= 0x01
        expect_error(flipped, "corrupt")                                    # (version 2: the header is under the checksum too)
    expect_error(raw[:-4], "truncated")
    expect_error(raw + b"\0" * 8, "trailing")
    expect_error(b"XXXXXXXX" + raw[8:], "magic")
    expect_error(raw[:40], "shorter")
    newer = bytearray(raw); newer[8] = 9
    expect_error(newer, "version")
    with pytest.raises(nbody_b200.NbodyCudaError):
        nbody_b200.checkpoint_info(str(tmp_path / "does_not_exist.ckp"))
    # too small an output buffer
    out = np.empty((10, 12), np.float32)
    assert lib.nbody_cuda_checkpoint_read(os.fsencode(path), out.ctypes.data_as(C.c_void_p), None, 10) == 1
    assert b"capacity" in lib.nbody_cuda_last_error()
    # argument validation of the writer
    hdr = nbody_b200.CheckpointHeader()
    hdr.config = nbody_b200.default_config()
    hdr.n_particles = 0
    assert lib.nbody_cuda_checkpoint_write(os.fsencode(path), C.byref(hdr), out.ctypes.data_as(C.c_void_p), None) == 1
    assert lib.nbody_cuda_checkpoint_write(None, C.byref(hdr), out.ctypes.data_as(C.c_void_p), None) == 1


def test_checkpoint_load_needs_a_device(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    path = str(tmp_path / "a.ckp")
    nbody_b200.checkpoint_write(path, particles(10))
    with pytest.raises(nbody_b200.NbodyCudaError) as e:
        nbody_b200.CudaSimulation.from_checkpoint(path)
    assert "no CPU fallback" in str(e.value)


def test_time_step_rule_matches_the_oracle_bit_for_bit():
    rng = np.random.default_rng(5)
    for _ in range(400):
        eta = float(np.float32(rng.choice([0.0, 0.01, 0.05, 0.3])))
        soft = float(np.float32(rng.choice([0.0, 0.01, 0.1])))
        dt0 = float(np.float32(10.0 ** rng.uniform(-4, -1)))
        lo = float(np.float32(rng.choice([0.0, dt0 * 0.1])))
        hi = float(np.float32(rng.choice([0.0, dt0 * 0.5, dt0 * 4])))
        if hi and lo > hi:
            lo = 0.0
        amax = float(np.float32(rng.choice([0.0, np.inf, np.nan, 10.0 ** rng.uniform(-3, 6)])))
        depth = int(rng.integers(1, 22))
        length = soft if soft > 0 else float(np.float32(2.0) * np.float32(2.0) ** -depth)
        got = nbody_b200.next_time_step(amax, bounds=[2.0, 1.0, 1.0], time_step=dt0, softening=soft, max_depth=depth,
                                        time_step_eta=eta, time_step_min=lo, time_step_max=hi)
        want = oracle.next_time_step(eta, length, amax, dt0, lo, hi)
        assert np.float32(got).tobytes() == np.float32(want).tobytes(), (eta, soft, dt0, lo, hi, amax, depth, got, want)


def test_time_step_rule_values():
    f = nbody_b200.next_time_step
    assert f(100.0, time_step=1e-3) == np.float32(1e-3)                                    # eta = 0: the reference's fixed step
    assert abs(f(400.0, time_step=1.0, time_step_eta=0.1) / (0.1 * np.sqrt(0.01 / 400.0)) - 1) < 1e-6
    assert f(1e-9, time_step=1e-3, time_step_eta=0.1) == np.float32(1e-3)                  # capped by time_step
    assert f(1e-9, time_step=1e-3, time_step_eta=0.1, time_step_max=5e-3) == np.float32(5e-3)
    assert f(1e12, time_step=1e-3, time_step_eta=0.1, time_step_min=1e-5) == np.float32(1e-5)
    for bad in (0.0, float("nan"), float("inf"), -1.0):
        assert f(bad, time_step=1e-3, time_step_eta=0.1) == np.float32(1e-3)


def test_time_step_config_validation():
    lib = nbody_b200.load_library()
    h = C.c_void_p()
    P = particles(4)
    for kw in ({"time_step_eta": -0.1}, {"time_step_min": -1.0}, {"time_step_min": 2e-3, "time_step_max": 1e-3},
               {"time_step_eta": 0.1, "time_step": 0.0}, {"time_step_eta": float("nan")}):
        cfg = nbody_b200.default_config(**kw)
        assert lib.nbody_cuda_create(C.byref(cfg), P.ctypes.data_as(C.c_void_p), 4, C.byref(h)) == 1, kw
        assert b"time_step" in lib.nbody_cuda_last_error()
    assert lib.nbody_cuda_set_time_step(None, 1e-3) == 1 and lib.nbody_cuda_get_time(None, None, None) == 1


def test_oracle_adaptive_run_is_consistent():
    from nbody_b200 import workloads
    n = 300
    P = workloads.uniform_cube(n)
    G = workloads.force_constant("uniform", n)
    Q, t, dts, amax = oracle.direct_step_adaptive(P, G, 0.01, 1e-3, 0.02, 0.0, 0.0, 6)
    assert dts[0] == np.float32(1e-3) and np.all(dts[1:] < 1e-3) and np.all(amax > 0)
    for s in range(5):
        assert dts[s + 1] == np.float32(oracle.next_time_step(0.02, 0.01, float(amax[s]), 1e-3))
    acc = np.float32(0)
    for d in dts:
        acc = np.float32(acc + d)
    assert np.float32(t) == acc                                     # FP32 accumulation like src/naive_simulation.cpp:44
    # with the rule switched off the run is the fixed-step run, bit for bit
    Qf, tf, dtf, _ = oracle.direct_step_adaptive(P, G, 0.01, 1e-3, 0.0, 0.0, 0.0, 3)
    Qr, tr = oracle.direct_step(P, G, 0.01, 1e-3, 3, 0)
    assert np.array_equal(Qf, Qr) and tf == tr and np.all(dtf == np.float32(1e-3))


def test_load_refuses_identities_that_are_not_a_permutation(tmp_path):
    """nbody_cuda_checkpoint_write accepts any identity array; nbody_cuda_checkpoint_load must not hand out-of-range or
    duplicated identities to callers that index with them. The check runs before any device is touched."""
    P = workloads.uniform_cube(64)
    for bad in (np.zeros(64, np.uint32), np.arange(64, dtype=np.uint32) + 1, np.r_[np.arange(63), 1000].astype(np.uint32)):
        path = str(tmp_path / "bad.ckp")
        nbody_b200.checkpoint_write(path, P, bad)
        with pytest.raises(nbody_b200.NbodyCudaError) as e:
            nbody_b200.CudaSimulation.from_checkpoint(path)
        assert "not a permutation" in str(e.value)


def test_denormal_softening_is_rejected():
    with pytest.raises(nbody_b200.NbodyCudaError) as e:
        nbody_b200.CudaSimulation([1, 1, 1], workloads.uniform_cube(8), 1e-3, softening=1e-30)
    assert "denormal" in str(e.value)


@pytest.mark.parametrize("field,value,needle", [
    ("order", 1, "order must be"), ("order", 6, "order must be"), ("max_depth", 0, "max_depth"), ("max_depth", 22, "max_depth"),
    ("leaf_capacity", 0, "leaf_capacity"), ("softening", -1.0, "softening"), ("mac_ratio", 0.0, "mac_ratio"), ("integrator", 2, "integrator"),
    ("low_order_tau", -0.1, "low_order_tau"), ("time_step_eta", -1.0, "time_step_eta"), ("abi_version", 99, "abi_version"),
])
def test_configuration_errors_are_reported_before_any_device_is_touched(field, value, needle):
    """nbody_cuda_create validates its configuration first (api.cu: validate), so every argument error reaches the caller as
    NBODY_ERR_INVALID with its message even on a box without a GPU — where every VALID configuration fails with "no CUDA device"."""
    with pytest.raises(nbody_b200.NbodyCudaError) as e:
        nbody_b200.CudaSimulation([1, 1, 1], workloads.uniform_cube(8), 1e-3, **{field: value})
    assert needle in str(e.value) and "no CUDA device" not in str(e.value), str(e.value)


def test_bounds_and_time_step_rules():
    for bounds, dt, kw, needle in (([1.0, 0.0, 1.0], 1e-3, {}, "bounds"), ([1, 1, 1], 1e-3, dict(time_step_min=2e-3, time_step_max=1e-3), "min <= max"),
                                   ([1, 1, 1], 0.0, dict(time_step_eta=0.1), "time_step > 0")):
        with pytest.raises(nbody_b200.NbodyCudaError) as e:
            nbody_b200.CudaSimulation(bounds, workloads.uniform_cube(8), dt, **kw)
        assert needle in str(e.value), str(e.value)


def test_checkpoint_reader_survives_mutated_files(tmp_path):
    """1500 random mutations of a valid file (bit flips, truncation, trailing bytes, random words in the header, zeroed runs): the
    reader never crashes, and accepts a file only if it is byte-identical to the original (the header is under the checksum too)."""
    rng = np.random.default_rng(5)
    path = str(tmp_path / "a.ckp")
    nbody_b200.checkpoint_write(path, particles(37), rng.permutation(37).astype(np.uint32), time=0.5, steps_done=3)
    raw = open(path, "rb").read()
    bad = str(tmp_path / "m.ckp")
    rejected = 0
    for _ in range(1500):
        b = bytearray(raw)
        kind = rng.integers(0, 5)
        if kind == 0:
            for _ in range(rng.integers(1, 4)):
                b[rng.integers(0, len(b))] ^= 1 << rng.integers(0, 8)
        elif kind == 1:
            b = b[:rng.integers(0, len(b))]
        elif kind == 2:
            b += bytes(rng.integers(0, 256, rng.integers(1, 64)).astype(np.uint8))
        elif kind == 3:
            at = rng.integers(0, 152 - 8)
            b[at:at + 8] = bytes(rng.integers(0, 256, 8).astype(np.uint8))
        else:
            at = rng.integers(0, len(b) - 16)
            b[at:at + 16] = bytes(16)
        open(bad, "wb").write(bytes(b))
        try:
            nbody_b200.checkpoint_read(bad)
            assert bytes(b) == raw
        except nbody_b200.NbodyCudaError:
            rejected += 1
    assert rejected > 1400
