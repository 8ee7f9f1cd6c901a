"""The numpy model of the proposed distributed sort (tests/models/slice_sort_model.py) must return
exactly the stable sort of all keys — the contract of the product's replicated radix sort and of the oracle — for any
partition, any boundary keys, with ties, empty slices, and on real Morton keys after a time step's motion."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "models"))
import slice_sort_model as m

import oracle
from nbody_b200 import workloads


def check(keys, part, K=None):
    keys = np.asarray(keys, np.uint64)
    ref = np.argsort(keys, kind="stable")
    k, i, n_em = m.slice_sort(keys, part, K)
    assert np.array_equal(i.astype(np.int64), ref) and np.array_equal(k, keys[ref])
    if K is not None:                                            # the sort-the-emigrants + 2-way-merge formulation agrees
        k2, i2, n2 = m.slice_sort_two_way(keys, part, K)
        assert np.array_equal(i2.astype(np.int64), ref) and np.array_equal(k2, keys[ref]) and n2 == n_em
    return n_em


@pytest.mark.parametrize("seed", range(6))
def test_random_keys_partitions_and_boundaries(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 3000))
    W = int(rng.integers(1, 9))
    keys = rng.integers(0, [1 << 62, 50, 7][seed % 3], n, dtype=np.uint64)          # wide keys, many ties, almost all ties
    part = np.concatenate([[0], np.sort(rng.integers(0, n + 1, W - 1)), [n]])       # includes empty slices
    check(keys, part)                                                               # no knowledge: everything but run 0 emigrates
    K = np.empty(W + 1, np.uint64)
    K[0], K[W] = 0, np.iinfo(np.uint64).max
    K[1:W] = np.sort(rng.integers(0, int(keys.max()) + 2, W - 1, dtype=np.uint64))
    check(keys, part, K)                                                            # arbitrary boundary keys never affect the result


def test_nearly_sorted_input_has_few_emigrants_and_jumpers_are_handled():
    rng = np.random.default_rng(1)
    n, W = 20000, 8
    prev = np.sort(rng.integers(0, 1 << 40, n, dtype=np.uint64))
    part = np.arange(W + 1) * n // W
    K = m.boundary_keys(prev, part)
    keys = prev.copy()
    moved = rng.choice(n, 300, replace=False)
    keys[moved] = keys[moved] + rng.integers(0, 1 << 20, 300, dtype=np.uint64)       # small drifts, mostly inside the slice
    jump = rng.choice(n, 20, replace=False)
    keys[jump] = rng.integers(0, 1 << 40, 20, dtype=np.uint64)                       # a plane crossing: lands anywhere in key space
    n_em = check(keys, part, K)
    assert n_em < 400                                                                # ~ the particles that left their interval, not N


def test_on_morton_keys_after_one_time_step():
    n, W = 30000, 8
    P = workloads.plummer(n)
    k0 = oracle.morton_keys(P[:, 0:3], (1.0, 1.0, 1.0))
    sk, perm = oracle.sort_keys(k0)
    P = P[perm]
    part = np.arange(W + 1) * n // W
    K = m.boundary_keys(sk, part)
    P[:, 0:3] += np.float32(1e-3) * P[:, 4:7]                                        # drift by one dt
    k1 = oracle.morton_keys(P[:, 0:3], (1.0, 1.0, 1.0))
    ok, operm = oracle.sort_keys(k1)
    ks, idx, n_em = m.slice_sort(k1, part, K)
    assert np.array_equal(ks, ok) and np.array_equal(idx.astype(np.uint32), operm)   # == the oracle's stable sort
    assert n_em < 0.2 * n
    print(f"emigrants after one step: {n_em} of {n}")
