"""CPU test of the distributed sort's merge rounds (nbody_b200/csrc/merge_path.h compiled for the host with g++, driven tile by
tile and thread by thread the way sort.cu:k_merge_runs does): slice-wise sorted runs merged pairwise must equal the stable sort
of all keys — the contract of the replicated radix sort and of the oracle — for any number of ranks, empty slices, heavy ties,
and real Morton keys after a step's motion."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def lib():
    src = os.path.join(HERE, "host", "merge_host.cpp")
    so = os.path.join(HERE, "host", "libmerge_host.so")
    hdr = os.path.join(HERE, "..", "nbody_b200", "csrc", "merge_path.h")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wall", "-Werror", src, "-o", so])
    L = C.CDLL(so)
    L.merge_all_runs.argtypes = [np.ctypeslib.ndpointer(np.uint64, flags="C"), np.ctypeslib.ndpointer(np.uint32, flags="C"), C.c_uint64,
                                 np.ctypeslib.ndpointer(np.uint32, flags="C"), C.c_int]
    return L


def dist_sort(lib, keys, part):
    """what NBODY_FLAG_DIST_SORT computes: per-slice stable sorts (the radix sort of a slice), then the merge rounds"""
    keys = np.ascontiguousarray(keys, np.uint64)
    k = keys.copy()
    v = np.arange(len(keys), dtype=np.uint32)
    for r in range(len(part) - 1):
        sl = slice(part[r], part[r + 1])
        o = np.argsort(keys[sl], kind="stable")
        k[sl] = keys[sl][o]
        v[sl] = (np.arange(part[r], part[r + 1], dtype=np.uint32))[o]
    rounds = lib.merge_all_runs(k, v, len(keys), np.ascontiguousarray(part, np.uint32), len(part) - 1)
    return k, v, rounds


def check(lib, keys, part):
    keys = np.asarray(keys, np.uint64)
    ref = np.argsort(keys, kind="stable")
    k, v, rounds = dist_sort(lib, keys, part)
    assert np.array_equal(v.astype(np.int64), ref) and np.array_equal(k, keys[ref])
    assert rounds == int(np.ceil(np.log2(max(len(part) - 1, 1))))


@pytest.mark.parametrize("seed", range(9))
def test_random_keys_and_partitions(lib, seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 40000))
    W = int(rng.integers(1, 17))
    keys = rng.integers(0, [1 << 62, 50, 3][seed % 3], n, dtype=np.uint64)      # wide keys, many ties, almost all ties
    part = np.concatenate([[0], np.sort(rng.integers(0, n + 1, W - 1)), [n]])   # includes empty slices
    check(lib, keys, part)


def test_edges(lib):
    T = lib.merge_tile_size()
    check(lib, [5], [0, 1])
    check(lib, [5, 5, 5, 5], [0, 1, 2, 3, 4])
    check(lib, np.arange(2 * T)[::-1], [0, T, 2 * T])                            # every B before every A, tile-aligned
    check(lib, np.arange(3 * T + 7), [0, T + 3, 3 * T + 7])                      # already sorted
    check(lib, np.zeros(5 * T + 1), [0, 1, T, T, 3 * T + 5, 5 * T + 1])          # all ties, an empty run, odd run count
    check(lib, np.r_[np.arange(T) * 2, np.arange(T) * 2 + 1], [0, T, 2 * T])     # perfect interleave
    check(lib, np.full(100, (1 << 63) - 1), [0, 50, 100])                        # the largest 63-bit key


def test_morton_keys_after_a_step(lib):
    """the real input: Morton keys of a Plummer model in the previous step's order, after one step of motion"""
    import oracle
    from nbody_b200 import workloads
    n = 60000
    P = workloads.plummer(n)
    k0 = oracle.morton_keys(P[:, 0:3], (1., 1., 1.))
    _, perm = oracle.sort_keys(k0)
    Ps = P[perm]
    moved = Ps[:, 0:3] + 5e-3 * Ps[:, 4:7] / max(np.abs(Ps[:, 4:7]).max(), 1e-30) * 3.0
    k1 = oracle.morton_keys(np.ascontiguousarray(np.clip(moved, 0, 0.999999).astype(P.dtype)), (1., 1., 1.))
    for W in (2, 4, 8):
        check(lib, k1, np.linspace(0, n, W + 1).astype(np.int64))
