"""The lane arithmetic of the leaf kernel's tile fill (nbody_b200/csrc/leaf.cu `stage`), restated per lane in
tests/models/leaf_fill_model.py, must deliver exactly the concatenated particle ranges of the source list, tile by tile:
for the default bitmap / 16-byte cp.async variant and for the experimental one-bulk-copy-per-run variant
(-DNBODY_LEAF_BULK=1), on random chains and on the real P2P lists of a Plummer model from the oracle."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "models"))
import leaf_fill_model as m


def check(segments):
    want = m.expected_tiles(segments)
    n = {}
    for bulk in (False, True):
        got, n[bulk] = m.stream(segments, bulk)
        assert len(got) == len(want)
        for (fill, tile), w in zip(got, want):
            assert fill == len(w) and np.array_equal(tile, w)
    flat = np.concatenate([np.arange(x, x + y) for seg in segments for x, y in np.asarray(seg).reshape(-1, 2)] or [np.zeros(0, int)])
    assert np.array_equal(np.concatenate([w for w in want] or [np.zeros(0, int)]), flat)     # nothing lost or repeated over the whole chain
    return n


def random_chain(rng, max_count, contiguous_prob):
    segments, pos = [], 0
    for _ in range(int(rng.integers(1, 5))):
        seg = []
        for _ in range(int(rng.integers(1, 90))):
            if rng.random() > contiguous_prob:
                pos += int(rng.integers(1, 1000))          # a gap: the next entry starts a new run
            cnt = int(rng.integers(1, max_count + 1))
            seg.append((pos, cnt))
            pos += cnt
        segments.append(seg)
    return segments


@pytest.mark.parametrize("seed", range(12))
def test_random_chains(seed):
    rng = np.random.default_rng(seed)
    check(random_chain(rng, [3, 48, 48, 700][seed % 4], [0.0, 0.5, 0.9, 0.7][seed % 4]))


def test_edge_cases():
    check([[(5, 1)]])                                                   # one particle
    check([[(0, 256)]])                                                 # exactly one tile
    check([[(0, 257)]])                                                 # one entry straddling
    check([[(0, 5000)]])                                                # one entry across 20 tiles
    check([[(10, 255), (265, 1), (300, 1)]])                            # an entry ending exactly at the tile end
    check([[(i * 8, 8) for i in range(32)]])                            # 32 contiguous entries = exactly one tile, one run
    check([[(i * 8, 8) for i in range(33)]])                            # ... and one more entry in the next fetch
    check([[(i * 100, 9) for i in range(64)]])                          # 32 entries < one tile: short tile, next fetch
    check([[(0, 300)], [(300, 300)], [(7, 1)]])                         # tiles never cross segments
    check([[(i * 2, 1) for i in range(200)]])                           # all runs of one particle
    n = check([[(1000 + i * 48, 48) for i in range(32)]])               # a fully merged list: 6 tiles
    assert n[True] == 6 and n[False] == 32 * 48                         # one bulk copy per tile instead of 256 16-byte copies


def test_real_p2p_lists_of_a_plummer_model():
    """Source lists of real leaves (oracle traversal, Plummer 20k, capacity 48): in particle order, where adjacent leaves merge
    into runs (tests/tools/p2p_list_structure.py), and shuffled, where nothing merges."""
    import oracle
    from nbody_b200 import workloads
    P = workloads.plummer(20000)
    sk, _ = oracle.sort_keys(oracle.morton_keys(P[:, 0:3], (1., 1., 1.)))
    t = oracle.Tree(sk, (1., 1., 1.), 48, 21)
    _, p2p = t.traverse(0.5)
    cnt, begin = np.asarray(t.leaf_count, np.int64), np.asarray(t.leaf_index, np.int64)
    a = np.concatenate([p2p[:, 0], p2p[:, 1]]).astype(np.int64)
    b = np.concatenate([p2p[:, 1], p2p[:, 0]]).astype(np.int64)
    key = np.unique(a * len(cnt) + b)
    a, b = key // len(cnt), key % len(cnt)
    order = np.lexsort((begin[b], a))
    a, b = a[order], b[order]
    cuts = np.flatnonzero(np.diff(a)) + 1
    rng = np.random.default_rng(0)
    total = {False: 0, True: 0}
    for src in np.split(b, cuts)[::7][:150]:
        entries = [(int(begin[j]), int(cnt[j])) for j in src if cnt[j]]
        if not entries:
            continue
        half = len(entries) // 2
        got = check([entries[:half], entries[half:]] if half else [entries])   # two segments, as the device pool chains them
        total[False] += got[False]
        total[True] += got[True]
        check([[entries[k] for k in rng.permutation(len(entries))]])
    assert total[True] * 10 < total[False]                                      # the point of the bulk variant: >10x fewer copy instructions
