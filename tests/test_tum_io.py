"""TUM dataset plumbing: association / trajectory parsing, trajectory formatting, PNG round trips
(CPU), and the vors_track CLI clone end to end on a synthetic TUM-layout dataset (GPU)."""
import os
import subprocess

import numpy as np
import pytest

from vors_b200 import synth, tum

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "visual-odometry-rs_b200", "bin", "vors_track")


def test_parse_associations_like_the_reference():
    # examples/README.md:26-31 format
    text = ("# depth_timestamp depth_file_path rgb_timestamp rgb_file_path\n"
            "1305031102.160407 depth/1305031102.160407.png 1305031102.175304 rgb/1305031102.175304.png\n"
            "1305031102.226738 depth/1305031102.226738.png 1305031102.211214 rgb/1305031102.211214.png\n")
    a = tum.parse_associations(text)
    assert a == [(1305031102.160407, "depth/1305031102.160407.png", 1305031102.175304, "rgb/1305031102.175304.png"),
                 (1305031102.226738, "depth/1305031102.226738.png", 1305031102.211214, "rgb/1305031102.211214.png")]
    with pytest.raises(ValueError):
        tum.parse_associations("1.0 depth/a.png\n")
    with pytest.raises(ValueError):
        tum.parse_associations("\n")  # a blank line is neither a comment nor an association (tum_rgbd.rs:111-118)


def test_trajectory_format_and_parse_round_trip():
    # tum_rgbd.rs:76-86: Rust `{}` prints shortest round-trip digits, no exponent, no trailing ".0"
    line = tum.frame_to_string(1305031098.6659, np.array([1.3563, 0.6305, 1.6380, 0.0, 0.0, 0.0, 1.0], np.float32))
    assert line == "1305031098.6659 1.3563 0.6305 1.638 0 0 0 1"
    assert tum.frame_to_string(0.5, np.array([1e-7, -2.5, 3, 0, 0, 0, 1], np.float32)).split()[1] == "0.0000001"
    (ts, p), = tum.parse_trajectory("# ground truth trajectory\n" + line + "\n")
    assert ts == 1305031098.6659 and np.allclose(p, [1.3563, 0.6305, 1.638, 0, 0, 0, 1])


def test_png_round_trip(tmp_path):
    scene, frames, _ = synth.make_sequence(seed=5, n_frames=2, rows=48, cols=64)
    path = tum.write_dataset(str(tmp_path), frames)
    assoc = tum.parse_associations(open(path).read())
    assert len(assoc) == 2
    d = tum.read_depth_png(os.path.join(tmp_path, assoc[1][1]))
    g = tum.read_gray_png(os.path.join(tmp_path, assoc[1][3]))
    assert np.array_equal(d, frames[1][1]) and np.array_equal(g, frames[1][0])


PROBE = os.path.join(ROOT, "visual-odometry-rs_b200", "bin", "png_probe")


def _probe(mode, png_path, tmp_path):
    """The CLI's own PNG reader (tools/png_reader.h) through the host-only probe command."""
    if not os.path.exists(PROBE):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "visual-odometry-rs_b200"), "-s", "bin/png_probe"])
    raw = os.path.join(tmp_path, "probe.raw")
    res = subprocess.run([PROBE, mode, png_path, raw], capture_output=True, text=True, timeout=60)
    if res.returncode != 0:
        raise ValueError(res.stderr.strip())
    rows, cols = [int(v) for v in res.stdout.split()]
    return np.fromfile(raw, dtype=np.uint16 if mode == "depth" else np.uint8).reshape(rows, cols)


def _luma(rgb):
    r, g, b = [rgb[..., k].astype(np.float32) for k in range(3)]
    return (np.float32(0.2126) * r + np.float32(0.7152) * g + np.float32(0.0722) * b).astype(np.uint8)


ADAM7 = ((0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2))


def _write_png(path, samples, bit_depth, color_type, interlace=0, plte=None):
    """An independent PNG encoder for the test (PIL writes neither interlaced nor sub-byte gray files): `samples` is
    [rows, cols, channels] of integers < 2^bit_depth; scanlines are packed most significant bit first, big-endian samples,
    Adam7 passes when `interlace`, and the five filter types are used in rotation so that the reader's unfiltering is
    exercised in every pass."""
    import struct
    import zlib

    def chunk(kind, body):
        return struct.pack(">I", len(body)) + kind + body + struct.pack(">I", zlib.crc32(kind + body) & 0xFFFFFFFF)

    samples = np.asarray(samples)
    h, w, ch = samples.shape
    bpp = max(1, ch * bit_depth // 8)

    def pack(rowvals):  # [pw, ch] -> bytes
        if bit_depth == 16:
            return rowvals.astype(">u2").tobytes()
        if bit_depth == 8:
            return rowvals.astype(np.uint8).tobytes()
        bits = "".join(format(int(v), f"0{bit_depth}b") for v in rowvals.reshape(-1))
        bits += "0" * (-len(bits) % 8)
        return int(bits, 2).to_bytes(len(bits) // 8, "big") if bits else b""

    def paeth(a, b, c):
        p = a + b - c
        pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
        return a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)

    raw = bytearray()
    line_no = 0
    for (x0, y0, dx, dy) in (ADAM7 if interlace else ((0, 0, 1, 1),)):
        sub = samples[y0::dy, x0::dx]
        if sub.shape[0] == 0 or sub.shape[1] == 0:
            continue
        prev = None
        for r in sub:
            cur = pack(r)
            up = prev if prev is not None else bytes(len(cur))
            f = line_no % 5
            line_no += 1
            out = bytearray([f])
            for i, v in enumerate(cur):
                a_ = cur[i - bpp] if i >= bpp else 0
                b_ = up[i]
                c_ = up[i - bpp] if i >= bpp else 0
                pred = (0, a_, b_, (a_ + b_) // 2, paeth(a_, b_, c_))[f]
                out.append((v - pred) & 0xFF)
            raw += out
            prev = cur
    data = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, bit_depth, color_type, 0, 0, 1 if interlace else 0))
    if plte is not None:
        data += chunk(b"PLTE", bytes(plte))
    z = zlib.compress(bytes(raw))  # two IDAT chunks: the compressed stream may be split anywhere
    data += chunk(b"IDAT", z[:len(z) // 2]) + chunk(b"IDAT", z[len(z) // 2:]) + chunk(b"IEND", b"")
    open(path, "wb").write(data)


def test_cli_png_reader_decodes_what_the_reference_decoder_accepts(tmp_path):
    """tools/png_reader.h on every PNG flavour `image::open().to_luma()` / `read_png_16bits` can meet: gray, gray+alpha,
    RGB, RGBA at 8 and 16 bits, gray at 1 / 2 / 4 bits, palettes at 1 / 2 / 4 / 8 bits, each plain and Adam7-interlaced, on
    odd sizes (ragged and empty interlace passes), with all five scanline filters.  Files come from PIL where PIL can write
    them and from the test's own encoder otherwise; PIL's decoder is the judge of both."""
    from PIL import Image

    rng = np.random.default_rng(7)
    n = 0
    for (h, w) in ((37, 53), (1, 1), (2, 3), (9, 8), (5, 4)):
        yy, xx = np.mgrid[0:h, 0:w]
        base = ((np.sin(xx / 3.0) + np.cos(yy / 5.0)) * 60 + 128 + rng.integers(-5, 6, (h, w))).clip(0, 255).astype(np.uint8)
        rgb = np.stack([base, np.roll(base, 1, 1), 255 - base], -1)
        rgba = np.concatenate([rgb, rng.integers(0, 256, (h, w, 1), dtype=np.uint8)], -1)
        la = np.stack([base, 255 - base], -1)
        d16 = (base.astype(np.uint16) * 257) ^ rng.integers(0, 65536, (h, w), dtype=np.uint16)
        rgb16 = rng.integers(0, 65536, (h, w, 3)).astype(np.uint16)

        def path_of(name):
            return os.path.join(tmp_path, f"{name}_{h}x{w}.png")

        # --- written by PIL (its encoder picks the scanline filters adaptively), not interlaced
        for name, img, mode, want in (("l", Image.fromarray(base), "gray", base), ("rgb", Image.fromarray(rgb), "gray", _luma(rgb)),
                                      ("rgba", Image.fromarray(rgba), "gray", _luma(rgb)), ("la", Image.fromarray(la), "gray", base),
                                      ("d16", Image.fromarray(d16), "depth", d16), ("d16g", Image.fromarray(d16), "gray", (d16 >> 8).astype(np.uint8))):
            img.save(path_of(name))
            assert np.array_equal(_probe(mode, path_of(name), tmp_path), want), (name, h, w)
            n += 1
        for colors in (256, 16, 2):
            Image.fromarray(rgb).quantize(colors=colors).save(path_of(f"p{colors}"))
            want = _luma(np.asarray(Image.open(path_of(f"p{colors}")).convert("RGB")))
            assert np.array_equal(_probe("gray", path_of(f"p{colors}"), tmp_path), want), (colors, h, w)
            n += 1
        # --- written by the test's encoder, plain and interlaced; PIL must read them the same way
        for il in (0, 1):
            cases = [("gray8", base[..., None], 8, 0, None, "gray", base), ("rgb8", rgb, 8, 2, None, "gray", _luma(rgb)),
                     ("rgba8", rgba, 8, 6, None, "gray", _luma(rgb)), ("la8", la, 8, 4, None, "gray", base),
                     ("gray16", d16[..., None], 16, 0, None, "depth", d16), ("rgb16", rgb16, 16, 2, None, "gray", _luma((rgb16 >> 8).astype(np.uint8)))]
            for bits in (1, 2, 4):
                v = rng.integers(0, 1 << bits, (h, w))
                cases.append((f"gray{bits}", v[..., None], bits, 0, None, "gray", (v * 255 // ((1 << bits) - 1)).astype(np.uint8)))
            for bits in (1, 2, 4, 8):
                v = rng.integers(0, 1 << bits, (h, w))
                plte = rng.integers(0, 256, ((1 << bits), 3)).astype(np.uint8)
                cases.append((f"pal{bits}", v[..., None], bits, 3, plte.reshape(-1).tolist(), "gray", _luma(plte[v])))
            for name, smp, bits, ctype, plte, mode, want in cases:
                f = path_of(f"{name}_il{il}")
                _write_png(f, smp, bits, ctype, interlace=il, plte=plte)
                assert np.array_equal(_probe(mode, f, tmp_path), want), (name, il, h, w)
                n += 1
                # PIL on the same file (it keeps 16-bit RGB as 8-bit, like the reader's "high byte")
                im = Image.open(f)
                assert im.info.get("interlace", 0) == il
                if name == "gray16":
                    assert np.array_equal(np.asarray(im).astype(np.uint16), d16)
                elif ctype in (0, 4):  # gray (+ alpha): PIL expands sub-byte samples the same way
                    assert np.array_equal(np.asarray(im.convert("L")), want), (name, il, h, w)
                elif name != "rgb16":
                    assert np.array_equal(_luma(np.asarray(im.convert("RGB"))), want), (name, il, h, w)
    assert n >= 5 * (9 + 2 * 13)


def test_cli_png_reader_rejects_damaged_files(tmp_path):
    from PIL import Image

    good = os.path.join(tmp_path, "good.png")
    Image.fromarray(np.arange(48, dtype=np.uint8).reshape(6, 8), "L").save(good)
    data = bytearray(open(good, "rb").read())
    at = data.index(b"IDAT") + 6
    data[at] ^= 0x40  # one flipped bit inside the compressed stream: the chunk CRC no longer matches
    bad = os.path.join(tmp_path, "bad.png")
    open(bad, "wb").write(data)
    with pytest.raises(ValueError, match="CRC"):
        _probe("gray", bad, tmp_path)
    open(bad, "wb").write(bytes(data[:-12]))  # IEND cut off
    with pytest.raises(ValueError, match="corrupt PNG"):
        _probe("gray", bad, tmp_path)
    open(bad, "wb").write(b"not a png at all")
    with pytest.raises(ValueError, match="not a PNG"):
        _probe("gray", bad, tmp_path)
    with pytest.raises(ValueError, match="16-bit gray"):
        _probe("depth", good, tmp_path)  # helper::read_png_16bits only takes 16-bit gray
    # a palette index beyond PLTE
    _write_png(bad, np.array([[[0], [5]]]), 8, 3, plte=[1, 2, 3, 4, 5, 6])
    with pytest.raises(ValueError, match="palette index"):
        _probe("gray", bad, tmp_path)


def test_cli_png_reader_agrees_with_the_python_twin_on_a_tum_dataset(tmp_path):
    """The files the GPU end-to-end test feeds the CLI (written by tum.write_dataset) decode to the same arrays through the
    CLI's reader and through the Python twin's (vors_b200/tum.py)."""
    scene, frames, _ = synth.make_sequence(seed=6, n_frames=2, rows=48, cols=64)
    for rgb in (False, True):
        root = os.path.join(tmp_path, f"rgb{int(rgb)}")
        os.makedirs(root)
        path = tum.write_dataset(root, frames, rgb=rgb)
        for a in tum.parse_associations(open(path).read()):
            dp, cp = os.path.join(root, a[1]), os.path.join(root, a[3])
            assert np.array_equal(_probe("depth", dp, tmp_path), tum.read_depth_png(dp))
            assert np.array_equal(_probe("gray", cp, tmp_path), tum.read_gray_png(cp))


@pytest.mark.gpu
@pytest.mark.parametrize("rgb", [False, True])
def test_vors_track_cli_matches_python_tracker_and_oracle(tmp_path, oracle, rgb):
    import vors_b200 as vb

    assert os.path.exists(CLI), "build with make -C visual-odometry-rs_b200"
    scene, frames, poses = synth.make_sequence(seed=60, n_frames=6, rows=480, cols=640, step_v=0.015, step_w=0.01)
    path = tum.write_dataset(str(tmp_path), frames, rgb=rgb)
    res = subprocess.run([CLI, "fr1", path], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    cli = tum.parse_trajectory(res.stdout)
    assert len(cli) == 5 and res.stderr.count("Optical_flow:") == 5
    if rgb:
        # R=G=B inputs go through the f32 luma formula (0.2126+0.7152+0.0722 rounds below 1 for some values)
        frames = [(tum.read_gray_png(os.path.join(tmp_path, a[3])), f[1]) for a, f in zip(tum.parse_associations(open(path).read()), frames)]
    py = tum.run_tracker("fr1", path)
    assert [l.split() for l in py] == [l.split() for l in res.stdout.strip().splitlines()]  # same library, same text
    cfg = oracle.default_config(nb_levels=6, **tum.INTRINSICS["fr1"])
    assoc = tum.parse_associations(open(path).read())
    ot = oracle.Tracker(cfg, assoc[0][0], frames[0][1], assoc[0][2], frames[0][0])
    for k in range(1, 6):
        ot.track(assoc[k][0], frames[k][1], assoc[k][2], frames[k][0])
        ts, p = ot.current_frame()
        assert cli[k - 1][0] == ts
        ang, dist = oracle.pose_error(cli[k - 1][1], p.as_array())
        assert ang <= 1e-4 and dist <= 1e-4, (k, ang, dist)


def test_vors_track_cli_argument_errors(tmp_path):
    if not os.path.exists(CLI):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "visual-odometry-rs_b200"), "-s"])
    res = subprocess.run([CLI], capture_output=True, text=True)
    assert "Usage: ./vors_track [fr1|fr2|fr3|icl] associations_file" in res.stderr and res.stdout == ""
    res = subprocess.run([CLI, "fr9", str(tmp_path / "nope.txt")], capture_output=True, text=True)
    assert "Unknown camera id: fr9" in res.stderr
