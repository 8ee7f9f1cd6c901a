"""Executable specification (numpy, CPU) of the distributed sort proposed in DESIGN.md section 10, item 1: "slice sort +
emigrant merge". Not product code and not on any product path: it exists so that the CUDA version can be written against a
model whose every step is checked (tests/test_slice_sort_model.py) against the definition of the result, a stable sort of
ALL keys, which is what the replicated radix sort computes today (sort.cu) and what the oracle's std::stable_sort defines.

Setting (comm.cu): every rank holds the full particle array in the tree order of the previous step; rank r owns the slice
[part[r], part[r+1]). A stable sort by key equals a sort by the composite (key, previous index), which has no ties.

  1. rank r sorts ITS slice by (key, index)                                  -> run r            (N/W elements per rank)
  2. boundary keys K[0..W] (any non-decreasing values with K[0] = 0, K[W] = +inf; efficiency, not correctness, depends on
     them: the keys at the slice starts of the previous step make ~99 % of a run "in place"): the in-place part of run r is
     [lo_r, hi_r) = keys in [K[r], K[r+1]); what lies before lo_r and after hi_r are the run's emigrants (a prefix and a suffix)
  3. all-gather the runs (12 B per particle); every rank derives lo/hi of every run by 2 W binary searches
  4. final position of an element x = number of elements smaller than x
       in-place x of run r at local index j:  sum_{s<r} (hi_s - lo_s) + (j - lo_r) + sum over the 2 W emigrant pieces of lower_bound(piece, x)
       emigrant x:                            sum over the 2 W emigrant pieces of lower_bound(piece, x)
                                              + sum_{s<t} (hi_s - lo_s) + lower_bound(in-place part of run t, x),  t = interval of x's key
     (in a kernel the emigrant-piece searches of in-place elements are done once per tile of consecutive elements: first and last
     element of the tile bracket a range of each piece that is almost always empty)."""
import numpy as np


def composite(keys, idx):
    """(key, previous index) as one sortable structured value"""
    c = np.empty(len(keys), dtype=[("k", np.uint64), ("i", np.uint64)])
    c["k"], c["i"] = keys, idx
    return c


def lower_bound(sorted_comp, x_comp):
    return np.searchsorted(sorted_comp, x_comp, side="left")


def local_runs(keys, part):
    """step 1: per rank, (sorted keys, previous global indices) of its slice"""
    runs = []
    for r in range(len(part) - 1):
        sl = slice(part[r], part[r + 1])
        k = keys[sl]
        order = np.argsort(k, kind="stable")
        runs.append((k[order], (np.arange(part[r], part[r + 1], dtype=np.uint64))[order]))
    return runs


def split_runs(runs, K):
    """step 2/3: lo_r, hi_r of every run for the boundary keys K[0..W]"""
    lo = np.array([np.searchsorted(k, K[r], side="left") for r, (k, _) in enumerate(runs)], dtype=np.int64)
    hi = np.array([np.searchsorted(k, K[r + 1], side="left") if r + 1 < len(runs) else len(k) for r, (k, _) in enumerate(runs)], dtype=np.int64)
    hi = np.maximum(hi, lo)
    return lo, hi


def merge_positions(runs, K):
    """step 4: final position of every element of every run; returns (positions per run, number of emigrants)"""
    W = len(runs)
    lo, hi = split_runs(runs, K)
    comps = [composite(k, i) for k, i in runs]
    inplace_before = np.concatenate([[0], np.cumsum(hi - lo)])
    pieces = [c[:lo[r]] for r, c in enumerate(comps)] + [c[hi[r]:] for r, c in enumerate(comps)]    # 2 W sorted emigrant pieces
    n_emigrants = int(sum(len(p) for p in pieces))
    pos = []
    for r, c in enumerate(comps):
        p = np.empty(len(c), np.int64)
        j = np.arange(len(c))
        emig_lt = sum(lower_bound(piece, c) for piece in pieces)      # for emigrants this includes their own piece: elements before them
        inpl = (j >= lo[r]) & (j < hi[r])
        p[inpl] = inplace_before[r] + (j[inpl] - lo[r]) + emig_lt[inpl]
        em = ~inpl
        if em.any():
            x = c[em]
            t = np.clip(np.searchsorted(K, x["k"], side="right") - 1, 0, W - 1)      # interval of the emigrant's key
            cnt = np.empty(len(x), np.int64)
            for tt in np.unique(t):
                sel = t == tt
                cnt[sel] = inplace_before[tt] + lower_bound(comps[tt][lo[tt]:hi[tt]], x[sel])
            p[em] = emig_lt[em] + cnt
        pos.append(p)
    return pos, n_emigrants


def slice_sort(keys, part, K=None):
    """The whole scheme: returns (sorted keys, permutation) == stable argsort of all keys, and the emigrant count."""
    keys = np.asarray(keys, np.uint64)
    part = np.asarray(part, np.int64)
    W = len(part) - 1
    if K is None:   # what a first step would use: nothing is known, every key is an emigrant of interval 0 except run 0's
        K = np.array([0] + [np.iinfo(np.uint64).max] * W, dtype=np.uint64)
    runs = local_runs(keys, part)
    pos, n_em = merge_positions(runs, np.asarray(K, np.uint64))
    out_k = np.empty(len(keys), np.uint64)
    out_i = np.empty(len(keys), np.uint64)
    for (k, i), p in zip(runs, pos):
        out_k[p], out_i[p] = k, i
    return out_k, out_i, n_em


def slice_sort_two_way(keys, part, K):
    """The formulation a kernel would use when emigrants are many (7.5 % of N per step on the Plummer benchmark at dt = 1e-3,
    independent of N: the slices meet in the core and a particle moves 0.18 core radii per step): compact the emigrants of all
    runs in rank order (prefix of run 0, suffix of run 0, prefix of run 1, ...), radix-sort them by key (stable, so ties keep
    ascending previous index), and 2-way merge them with the concatenated in-place parts, which are globally sorted already:
        position of A[j] = j + lower_bound(B, A[j]),   position of B[j] = j + lower_bound(A, B[j])   (composite compares)."""
    keys = np.asarray(keys, np.uint64)
    part = np.asarray(part, np.int64)
    runs = local_runs(keys, part)
    lo, hi = split_runs(runs, np.asarray(K, np.uint64))
    A = np.concatenate([composite(k[lo[r]:hi[r]], i[lo[r]:hi[r]]) for r, (k, i) in enumerate(runs)])
    E = np.concatenate([np.concatenate([composite(k[:lo[r]], i[:lo[r]]), composite(k[hi[r]:], i[hi[r]:])]) for r, (k, i) in enumerate(runs)])
    B = E[np.argsort(E["k"], kind="stable")]
    assert np.array_equal(np.sort(A, order=("k", "i")), A)       # in-place parts of successive intervals: sorted as they stand
    assert np.array_equal(np.sort(B, order=("k", "i")), B)       # stable by key == sorted by (key, previous index)
    out = np.empty(len(keys), dtype=A.dtype)
    out[np.arange(len(A)) + lower_bound(B, A)] = A
    out[np.arange(len(B)) + lower_bound(A, B)] = B
    return out["k"], out["i"], len(B)


def boundary_keys(prev_sorted_keys, part):
    """K[r] = key at the start of slice r in the PREVIOUS step's sorted order (K[0] = 0, K[W] = max)"""
    W = len(part) - 1
    K = np.empty(W + 1, np.uint64)
    K[0], K[W] = 0, np.iinfo(np.uint64).max
    for r in range(1, W):
        K[r] = prev_sorted_keys[min(part[r], len(prev_sorted_keys) - 1)]
    return np.maximum.accumulate(K)
