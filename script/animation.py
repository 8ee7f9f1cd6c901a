#!/usr/bin/env python3
"""Animate a particles.csv written by the driver (examples/nbody_main.cpp, or the reference's src/main.cpp:88-95):
one row per step, ``time,x0,y0,z0,x1,y1,z1,...``.

Python-3 counterpart of the reference's script/animation.py:9-47 (Python 2, interactive matplotlib 3-D scatter over the
unit cube). Differences, all forced by scale and by running on GPU boxes without a display:
  * headless by default: the frames are rasterised with numpy and written as ONE animated PNG (APNG: plays in any
    browser; other viewers show the first frame) using only zlib; no matplotlib, no display;
  * ``--show`` opens the reference's interactive matplotlib window instead, when matplotlib and a display exist;
  * large runs are subsampled (``--max-particles``, evenly strided, the same particles in every frame) and frames can
    be strided (``--every``), because a 16M-particle row is 400 MB of text.

    python script/animation.py particles.csv --out particles.png [--size 512] [--max-particles 20000] [--every 1]
"""
import argparse
import struct
import sys
import zlib

import numpy as np


def read_csv(path, max_particles=20000, every=1):
    """-> (times [F], positions [F, K, 3]); K <= max_particles evenly strided columns, every `every`-th row."""
    times, frames, stride, count = [], [], None, None
    with open(path, "r") as f:
        for r, line in enumerate(f):
            line = line.strip()
            if not line:                      # the reference skips empty rows too (script/animation.py:17-18)
                continue
            if r % max(every, 1):
                continue
            row = np.array(line.split(","), dtype=np.float64)   # whole row at once: ~20x faster than csv.reader + float()
            if (row.size - 1) % 3:
                raise ValueError(f"{path}: row {r} has {row.size - 1} coordinates, not a multiple of 3")
            n = (row.size - 1) // 3
            if count is None:
                count = n
                stride = max(1, -(-n // max(max_particles, 1)))
            elif n != count:
                raise ValueError(f"{path}: row {r} has {n} particles, the first row had {count}")
            times.append(row[0])
            frames.append(row[1:].reshape(n, 3)[::stride].astype(np.float32))
    if not frames:
        raise ValueError(f"{path}: no data rows")
    return np.array(times), np.stack(frames)


def view_matrix(azim_deg=-60.0, elev_deg=30.0):
    """Rotation taking world (x, y, z) to (right, up, towards the viewer); mplot3d's default view angles."""
    a, e = np.radians(azim_deg), np.radians(elev_deg)
    to_viewer = np.array([np.cos(e) * np.cos(a), np.cos(e) * np.sin(a), np.sin(e)])
    right = np.array([-np.sin(a), np.cos(a), 0.0])
    up = np.cross(to_viewer, right)
    return np.stack([right, up, to_viewer])


def project(points, bounds, size, view):
    """Orthographic projection of points in [0, bounds) to pixel coordinates of a size x size image (+ depth in [0, 1])."""
    c = (np.asarray(points, np.float64) / np.asarray(bounds, np.float64) - 0.5) @ view.T     # cube centred on the origin
    scale = size / 1.8                                                                      # the cube's diagonal (1.73) fits
    px = size / 2 + c[..., 0] * scale
    py = size / 2 - c[..., 1] * scale
    return px, py, np.clip(c[..., 2] / 1.74 + 0.5, 0.0, 1.0)


def draw_box(img, bounds, size, view):
    corners = np.array([[i, j, k] for i in (0, 1) for j in (0, 1) for k in (0, 1)], np.float64) * np.asarray(bounds, np.float64)
    for a in range(8):
        for b in range(a + 1, 8):
            if np.count_nonzero(corners[a] != corners[b]) != 1:
                continue                                                                    # an edge joins corners differing in one axis
            t = np.linspace(0.0, 1.0, 2 * size)[:, None]
            px, py, _ = project(corners[a] * (1 - t) + corners[b] * t, bounds, size, view)
            ix, iy = np.round(px).astype(int), np.round(py).astype(int)
            ok = (ix >= 0) & (ix < size) & (iy >= 0) & (iy < size)
            img[iy[ok], ix[ok]] = np.maximum(img[iy[ok], ix[ok]], 70)


def render_frame(points, bounds=(1.0, 1.0, 1.0), size=512, view=None):
    """One greyscale frame (uint8 [size, size]): box edges + particles splatted as 2x2 dots, nearer ones brighter,
    overlapping ones accumulating so that dense regions (a Plummer core) saturate to white."""
    view = view_matrix() if view is None else view
    img = np.zeros((size, size), np.uint8)
    draw_box(img, bounds, size, view)
    px, py, depth = project(points, bounds, size, view)
    acc = np.zeros((size, size), np.float32)
    w = (0.35 + 0.65 * depth).astype(np.float32)
    for dx in (0, 1):
        for dy in (0, 1):
            ix, iy = np.floor(px).astype(int) + dx, np.floor(py).astype(int) + dy
            ok = (ix >= 0) & (ix < size) & (iy >= 0) & (iy < size)
            np.add.at(acc, (iy[ok], ix[ok]), w[ok])
    lit = np.clip(90.0 + 165.0 * np.minimum(acc, 2.0) / 2.0, 0, 255).astype(np.uint8)
    return np.where(acc > 0, np.maximum(lit, img), img)


def _chunk(tag, data):
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)


def write_apng(path, frames, delay_ms=40):
    """frames: uint8 [F, H, W] greyscale -> animated PNG (PNG 1.2 + APNG chunks acTL / fcTL / fdAT), zlib only."""
    frames = np.asarray(frames, np.uint8)
    nf, h, w = frames.shape
    out = [b"\x89PNG\r\n\x1a\n", _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 0, 0, 0, 0)),
           _chunk(b"acTL", struct.pack(">II", nf, 0))]
    seq = 0
    for k in range(nf):
        raw = np.concatenate([np.zeros((h, 1), np.uint8), frames[k]], axis=1).tobytes()     # filter type 0 on every scanline
        data = zlib.compress(raw, 6)
        out.append(_chunk(b"fcTL", struct.pack(">IIIIIHHBB", seq, w, h, 0, 0, delay_ms, 1000, 0, 0)))
        seq += 1
        if k == 0:
            out.append(_chunk(b"IDAT", data))                                               # frame 0 doubles as the still image
        else:
            out.append(_chunk(b"fdAT", struct.pack(">I", seq) + data))
            seq += 1
    out.append(_chunk(b"IEND", b""))
    with open(path, "wb") as f:
        f.write(b"".join(out))


def show_interactive(times, pos, bounds):
    """The reference's window (script/animation.py:24-44): matplotlib 3-D scatter, axes fixed to the bounds."""
    import matplotlib.pyplot as plt
    import matplotlib.animation as anim
    fig = plt.figure()
    ax = fig.add_subplot(projection="3d")
    for setter, label, b in ((ax.set_xlim3d, "X", bounds[0]), (ax.set_ylim3d, "Y", bounds[1]), (ax.set_zlim3d, "Z", bounds[2])):
        setter([0.0, b])
    ax.set_xlabel("X"); ax.set_ylabel("Y"); ax.set_zlabel("Z")
    points = ax.scatter(pos[0, :, 0], pos[0, :, 1], pos[0, :, 2], s=2)

    def animate(i):
        points._offsets3d = (pos[i, :, 0], pos[i, :, 1], pos[i, :, 2])
        ax.set_title(f"t = {times[i]:.5g}")
    keep = anim.FuncAnimation(fig, animate, interval=10, frames=pos.shape[0])  # noqa: F841 (must stay referenced)
    plt.show()


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("csv", help="particles.csv: one row per step, time,x0,y0,z0,...")
    ap.add_argument("--out", default="particles.png", help="animated PNG to write (default particles.png)")
    ap.add_argument("--size", type=int, default=512, help="frame edge in pixels")
    ap.add_argument("--max-particles", type=int, default=20000, help="subsample to at most this many particles")
    ap.add_argument("--every", type=int, default=1, help="use every k-th row")
    ap.add_argument("--bounds", type=float, nargs=3, default=[1.0, 1.0, 1.0], help="simulation box (src/main.cpp:24)")
    ap.add_argument("--delay-ms", type=int, default=40)
    ap.add_argument("--show", action="store_true", help="interactive matplotlib window instead of a file")
    args = ap.parse_args(argv)
    times, pos = read_csv(args.csv, args.max_particles, args.every)
    if args.show:
        try:
            show_interactive(times, pos, args.bounds)
            return 0
        except ImportError:
            print("matplotlib is not installed: writing the animated PNG instead", file=sys.stderr)
    view = view_matrix()
    frames = np.stack([render_frame(p, args.bounds, args.size, view) for p in pos])
    write_apng(args.out, frames, args.delay_ms)
    print(f"{args.out}: {pos.shape[0]} frames of {pos.shape[1]} particles, t = {times[0]:.6g} .. {times[-1]:.6g}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
