"""Executable specification of the partitioned tree build (nbody_b200/csrc/let.cu, tree.cu: node_splits), in numpy.

Rank r owns the particles whose Morton key lies in [split[r], split[r+1]). A cell splits when it holds more than `cap`
particles. A cell whose key range contains a splitter strictly inside ("straddling" cell) holds particles of
several ranks: the ranks exchange their local counts of those cells (at most one per splitter and level) and split them by the
GLOBAL count. Claim (checked by tests/test_partitioned_model.py against the oracle's tree): every rank's tree is then the global
octree restricted to the cells that hold its own particles — same cells, same has_children."""
import numpy as np

MAX_DEPTH = 21
KEY_END = 1 << 63


def cell_range(depth, prefix):
    sh = 3 * (MAX_DEPTH - depth)
    return prefix << sh, (prefix + 1) << sh


def count_in(keys_sorted, lo, hi):
    return int(np.searchsorted(keys_sorted, np.uint64(hi) if hi < (1 << 64) else np.uint64((1 << 64) - 1), "left") -
               np.searchsorted(keys_sorted, np.uint64(lo), "left")) if hi <= KEY_END else 0


def straddling_cells(split):
    """{(depth, prefix)} of the cells that contain a splitter strictly inside their key range (k_let_straddle)."""
    cells = []
    for b in range(1, len(split) - 1):
        K = int(split[b])
        for d in range(MAX_DEPTH + 1):
            sh = 3 * (MAX_DEPTH - d)
            lo = (K >> sh) << sh
            if K != lo and K < KEY_END:
                cells.append((b, d, K >> sh))
    return cells


def local_straddle_counts(keys_sorted, split):
    return [count_in(keys_sorted, *cell_range(d, p)) for (_, d, p) in straddling_cells(split)]


def build_tree(keys_sorted, cap, max_depth, forced=frozenset()):
    """Cells of the octree of `keys_sorted` as {(depth, prefix): (count, has_children)}; a split cell has all 8 children.
    `forced`: cells that split whatever their local count (global count > cap), provided they hold a local particle."""
    nodes = {}
    frontier = [(0, 0)]
    while frontier:
        nxt = []
        for (d, p) in frontier:
            c = count_in(keys_sorted, *cell_range(d, p))
            split = d < max_depth and (c > cap or (c > 0 and (d, p) in forced))
            nodes[(d, p)] = (c, split)
            if split:
                nxt.extend((d + 1, (p << 3) | k) for k in range(8))
        frontier = nxt
    return nodes
