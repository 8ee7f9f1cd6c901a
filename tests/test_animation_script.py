"""CPU tests of script/animation.py, the Python-3 counterpart of the reference's script/animation.py:9-47: it reads the
CSV rows the drivers write (src/main.cpp:88-95: time,x0,y0,z0,...), subsamples, renders headless and writes a valid APNG."""
import importlib.util
import os
import struct
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("animation", os.path.join(ROOT, "script", "animation.py"))
animation = importlib.util.module_from_spec(spec)
spec.loader.exec_module(animation)


def write_csv(path, times, pos):
    with open(path, "w") as f:
        for t, p in zip(times, pos):
            f.write(",".join([repr(float(t))] + [repr(float(v)) for v in p.reshape(-1)]) + "\n")
            f.write("\n")                                   # blank rows are skipped like the reference does


def test_read_csv_subsample_and_stride(tmp_path):
    rng = np.random.default_rng(0)
    times = np.arange(1, 7) * 1e-3
    pos = rng.random((6, 1000, 3))
    path = str(tmp_path / "particles.csv")
    write_csv(path, times, pos)
    t, p = animation.read_csv(path, max_particles=10**6)
    np.testing.assert_allclose(t, times)
    np.testing.assert_allclose(p, pos, rtol=1e-6)
    t, p = animation.read_csv(path, max_particles=100, every=1)
    assert p.shape == (6, 100, 3)
    np.testing.assert_allclose(p, pos[:, ::10], rtol=1e-6)   # the same particles in every frame
    open(str(tmp_path / "bad.csv"), "w").write("0.001,1,2,3,4\n")
    with pytest.raises(ValueError):
        animation.read_csv(str(tmp_path / "bad.csv"))
    open(str(tmp_path / "ragged.csv"), "w").write("0.001,1,2,3\n0.002,1,2,3,4,5,6\n")
    with pytest.raises(ValueError):
        animation.read_csv(str(tmp_path / "ragged.csv"))
    open(str(tmp_path / "empty.csv"), "w").write("\n\n")
    with pytest.raises(ValueError):
        animation.read_csv(str(tmp_path / "empty.csv"))


def test_projection_and_frame():
    v = animation.view_matrix()
    np.testing.assert_allclose(v @ v.T, np.eye(3), atol=1e-12)                  # a rotation
    np.testing.assert_allclose(np.cross(v[0], v[1]), v[2], atol=1e-12)          # right-handed: right x up = towards the viewer
    px, py, d = animation.project(np.array([[0.5, 0.5, 0.5]]), (1, 1, 1), 256, v)
    assert abs(px[0] - 128) < 1e-9 and abs(py[0] - 128) < 1e-9 and abs(d[0] - 0.5) < 1e-9
    corners = np.array([[i, j, k] for i in (0, 2) for j in (0, 1) for k in (0, 1)], float)
    px, py, d = animation.project(corners, (2, 1, 1), 256, v)                  # every corner of a non-cubic box is inside the frame
    assert px.min() >= 0 and px.max() <= 255 and py.min() >= 0 and py.max() <= 255
    _, py_top, _ = animation.project(np.array([[0.5, 0.5, 1.0]]), (1, 1, 1), 256, v)
    assert py_top[0] < 128                                                      # +z is up on the screen
    empty = animation.render_frame(np.zeros((0, 3)), size=128)
    one = animation.render_frame(np.array([[0.5, 0.5, 0.5]]), size=128)
    assert empty.dtype == np.uint8 and empty.shape == (128, 128) and empty.max() == 70      # only the box edges
    assert one[64, 64] > 90 and (one != empty).sum() <= 4                       # a 2x2 dot at the centre
    far_outside = animation.render_frame(np.array([[50.0, -30.0, 9.0]]), size=128)
    assert np.array_equal(far_outside, empty)                                   # escaped particles are clipped, not wrapped


def parse_png(raw):
    assert raw[:8] == b"\x89PNG\r\n\x1a\n"
    chunks, off = [], 8
    while off < len(raw):
        n, = struct.unpack_from(">I", raw, off)
        tag, data = raw[off + 4:off + 8], raw[off + 8:off + 8 + n]
        crc, = struct.unpack_from(">I", raw, off + 8 + n)
        assert crc == zlib.crc32(tag + data) & 0xFFFFFFFF, tag
        chunks.append((tag, data))
        off += 12 + n
    return chunks


def test_end_to_end_writes_a_valid_apng(tmp_path, capsys):
    rng = np.random.default_rng(1)
    pos = rng.random((5, 300, 3))
    csv, out = str(tmp_path / "particles.csv"), str(tmp_path / "anim.png")
    write_csv(csv, np.arange(1, 6) * 1e-3, pos)
    assert animation.main([csv, "--out", out, "--size", "96", "--max-particles", "150"]) == 0
    assert "5 frames of 150 particles" in capsys.readouterr().out
    chunks = parse_png(open(out, "rb").read())
    tags = [t for t, _ in chunks]
    assert tags[0] == b"IHDR" and tags[-1] == b"IEND" and tags.count(b"fcTL") == 5 and tags.count(b"IDAT") == 1 and tags.count(b"fdAT") == 4
    w, h, depth, colour = struct.unpack(">IIBB", chunks[0][1][:10])
    assert (w, h, depth, colour) == (96, 96, 8, 0)
    assert struct.unpack(">II", dict(chunks)[b"acTL"]) == (5, 0)
    # sequence numbers run 0..8 over fcTL and fdAT, and every frame inflates to h * (w + 1) bytes equal to the rendered frame
    seqs, frames = [], []
    for tag, data in chunks:
        if tag == b"fcTL":
            seqs.append(struct.unpack(">I", data[:4])[0])
        elif tag == b"fdAT":
            seqs.append(struct.unpack(">I", data[:4])[0])
            frames.append(zlib.decompress(data[4:]))
        elif tag == b"IDAT":
            frames.append(zlib.decompress(data))
    assert seqs == list(range(9))
    view = animation.view_matrix()
    for k, raw in enumerate(frames):
        img = np.frombuffer(raw, np.uint8).reshape(96, 97)
        assert np.all(img[:, 0] == 0)
        assert np.array_equal(img[:, 1:], animation.render_frame(pos[k, ::2].astype(np.float32), (1, 1, 1), 96, view))
