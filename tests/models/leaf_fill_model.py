"""Executable specification (numpy, CPU) of the leaf kernel's tile fill, nbody_b200/csrc/leaf.cu `stage` (both variants):
the lane arithmetic that cuts 256-particle tiles off a leaf's source list and decides which global particle lands in
which shared-memory slot. Not product code and not on any product path: it restates the per-lane expressions of the
kernel one for one (same names) so that tests/test_leaf_fill_model.py can check them against the definition of the result
— the concatenation of the entries' particle ranges, cut into tiles of 256 — for arbitrary segment chains, including the
experimental TMA variant (-DNBODY_LEAF_BULK=1), which has not run on hardware yet.

A source list is a chain of segments; a segment is an array of entries {first particle, count}. The cursor is
(segment, e0 = first unconsumed entry of the segment, skip = particles of entry e0 already consumed). `fetch` hands lane l
entry e0 + l (or {0, 0} past the segment's end); `stage` cuts ONE tile off the flat particle range of those (up to) 32
entries and advances the cursor; a tile never crosses a segment boundary (the last tile of a segment is short)."""
import numpy as np

TILE = 256
LANES = 32


def shfl(v, src):
    """__shfl_sync(v, src): src may be a scalar or a per-lane array"""
    return v[np.asarray(src) & 31]


def stage(ent_x, ent_y, seg_cnt, e0, skip, bulk):
    """One call of the kernel's `stage` lambda for the 32 entries in (ent_x, ent_y).
    Returns (tile, fill, e0, skip, copies): tile[f] = global particle index copied to slot f (-1 = untouched),
    copies = list of (dst slot, src particle, count) issued (one per 16-byte cp.async or one per bulk copy)."""
    lane = np.arange(LANES, dtype=np.int64)
    nvalid = min(32, seg_cnt - e0)
    sk = np.where(lane == 0, skip, 0)
    v = np.where(lane < nvalid, ent_y - sk, 0)
    inc = np.cumsum(v)                                               # the shuffle-up inclusive scan
    full = (lane < nvalid) & (inc <= TILE)
    nfull = int(full.sum())
    assert np.array_equal(full, lane < nfull), "entries that end inside the tile are a prefix of the lanes"
    avail = int(inc[nvalid - 1])
    used = int(inc[(nfull - 1) & 31])
    fill = min(avail, TILE)
    if nfull < nvalid:
        skip = (0 if nfull else skip) + (fill - (used if nfull else 0))
    else:
        skip = 0
    e0 += nfull
    tile = np.full(TILE, -1, dtype=np.int64)
    copies = []
    if bulk:
        prev_end = np.concatenate([[0], (ent_x + ent_y)[:-1]])       # __shfl_up by 1 (lane 0 keeps its own value: unused)
        head = (lane < nvalid) & ((lane == 0) | (ent_x != prev_end))
        heads = int(sum(1 << int(l) for l in lane[head]))
        start = inc - v
        tx = 0
        for l in range(LANES):
            above = heads >> (l + 1) if l < 31 else 0
            nxt = l + ((above & -above).bit_length()) if above else nvalid    # lane + __ffs(above)
            run_end = min(int(start[nxt & 31]), avail)
            stop = min(run_end if nxt < nvalid else avail, TILE)
            if head[l] and start[l] < stop:
                cnt = int(stop - start[l])
                copies.append((int(start[l]), int(ent_x[l] + sk[l]), cnt))
                tx += 16 * cnt
        assert tx == 16 * fill, "the bytes the mbarrier expects are the bytes the bulk copies deliver"
    else:
        src0 = ent_x + sk - (inc - v)
        pc = 0
        for i in range(TILE // 32):
            if 32 * i >= fill:
                break
            word = 0
            for l in range(LANES):
                if full[l] and (inc[l] >> 5) == i:
                    word |= 1 << int(inc[l] & 31)
            for l in range(LANES):
                f = 32 * i + l
                le_mask = (2 << l) - 1
                e = pc + bin(word & le_mask).count("1")
                s0 = int(src0[e & 31])
                if f < fill:
                    copies.append((f, s0 + f, 1))
            pc += bin(word).count("1")
    for dst, src, cnt in copies:
        assert np.all(tile[dst:dst + cnt] == -1), "no slot is written twice"
        tile[dst:dst + cnt] = np.arange(src, src + cnt)
    return tile, fill, e0, skip, copies


def stream(segments, bulk):
    """Walk a whole chain the way the kernel does (fetch / stage until the chain ends).
    Returns (list of (fill, tile[:fill]) per tile, number of copy instructions issued)."""
    tiles, ncopies = [], 0
    for seg in segments:
        seg = np.asarray(seg, dtype=np.int64).reshape(-1, 2)
        e0, skip = 0, 0
        while e0 < len(seg):
            ent = np.zeros((LANES, 2), dtype=np.int64)
            chunk = seg[e0:e0 + LANES]
            ent[:len(chunk)] = chunk
            tile, fill, e0, skip, copies = stage(ent[:, 0], ent[:, 1], len(seg), e0, skip, bulk)
            tiles.append((fill, tile[:fill].copy()))
            assert np.all(tile[fill:] == -1)
            ncopies += len(copies)
    return tiles, ncopies


def expected_tiles(segments):
    """The definition: per segment, the concatenated particle ranges of its entries, cut every 256 particles... except that a
    batch of 32 entries shorter than a tile also ends a tile (the kernel never mixes two fetches in one tile)."""
    out = []
    for seg in segments:
        seg = np.asarray(seg, dtype=np.int64).reshape(-1, 2)
        e0, skip = 0, 0
        while e0 < len(seg):
            flat = []
            for j, (x, y) in enumerate(seg[e0:e0 + LANES]):
                lo = skip if j == 0 else 0
                flat.append(np.arange(x + lo, x + y))
            ends = np.cumsum([len(f) for f in flat])
            flat = np.concatenate(flat)
            take = min(len(flat), TILE)
            out.append(flat[:take])
            whole = int(np.searchsorted(ends, take, side="right"))    # entries wholly consumed
            if whole < len(ends):
                before = int(ends[whole - 1]) if whole else 0
                skip = (skip if whole == 0 else 0) + (take - before)
            else:
                skip = 0
            e0 += whole
    return out
