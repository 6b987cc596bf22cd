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
