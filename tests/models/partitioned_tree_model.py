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


# ---- the traversal over (target tree, source tree): the rule of src/interaction.cl:22-99 as nbody_b200/csrc/traverse.cu organises it
# (directed, by target), for two trees that may belong to different ranks ----------------------------------------------------------
def cell_geometry(depth, prefix):
    """Centre (x, y, z) and edge of cell (depth, prefix) in the unit box; digit = x | y << 1 | z << 2, most significant digit first."""
    ix = iy = iz = 0
    for level in range(depth):
        digit = (prefix >> (3 * (depth - 1 - level))) & 7
        ix = ix << 1 | (digit & 1); iy = iy << 1 | (digit >> 1 & 1); iz = iz << 1 | (digit >> 2 & 1)
    size = np.float32(1.0) / np.float32(1 << depth)
    half = size * np.float32(0.5)
    return (np.float32(ix) * size + half, np.float32(iy) * size + half, np.float32(iz) * size + half), size


def mac_accept(ca, cb):
    """can_approx of src/interaction.cl:64-82 in FP32 (cells of different trees at the same place have d2 = 0: never accepted)."""
    (ax, ay, az), sa = cell_geometry(*ca)
    (bx, by, bz), sb = cell_geometry(*cb)
    dx, dy, dz = bx - ax, by - ay, bz - az
    d2 = np.float32(dx * dx) + np.float32(dy * dy) + np.float32(dz * dz)
    ext = np.float32(sa + sb)
    ext2 = np.float32(np.float32(np.float32(0.75) * ext) * ext)
    return bool(d2 > 0 and np.float32(ext2 / d2) < np.float32(0.25))


def traverse(target_tree, source_tree):
    """Directed lists for the targets of `target_tree` against the sources of `source_tree`: ([(target cell, source cell)] M2L, P2P)."""
    m2l, p2p = [], []
    stack = [((0, 0), (0, 0))]
    while stack:
        A, B = stack.pop()
        ca_list = [(A[0] + 1, (A[1] << 3) | k) for k in range(8)] if target_tree[A][1] else [A]
        cb_list = [(B[0] + 1, (B[1] << 3) | k) for k in range(8)] if source_tree[B][1] else [B]
        for ca in ca_list:
            if target_tree[ca][0] == 0:
                continue
            for cb in cb_list:
                if source_tree[cb][0] == 0:
                    continue
                if mac_accept(ca, cb):
                    m2l.append((ca, cb))
                elif target_tree[ca][1] or source_tree[cb][1]:
                    stack.append((ca, cb))
                else:
                    p2p.append((ca, cb))
    return m2l, p2p
