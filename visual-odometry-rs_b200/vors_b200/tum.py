"""TUM RGB-D dataset plumbing around the tracker: what src/dataset/tum_rgbd.rs, src/misc/helper.rs
and src/bin/vors_track.rs do around `Tracker::track`.  Host-side I/O only — no numerics of the hot path live here."""
from __future__ import annotations

import os

import numpy as np

DEPTH_SCALE = 5000.0  # src/dataset/tum_rgbd.rs:15
# src/dataset/tum_rgbd.rs:23-51 (principal point, focal)
INTRINSICS = {
    "icl": dict(cx=319.5, cy=239.5, fx=481.20, fy=-480.00, skew=0.0),
    "fr1": dict(cx=318.643040, cy=255.313989, fx=517.306408, fy=516.469215, skew=0.0),
    "fr2": dict(cx=325.141442, cy=249.701764, fx=520.908620, fy=521.007327, skew=0.0),
    "fr3": dict(cx=320.106653, cy=247.632132, fx=535.433105, fy=539.212524, skew=0.0),
}


def parse_associations(text: str):
    """tum_rgbd::parse::associations (tum_rgbd.rs:97-145): comment lines start with '#'; every other line is
    `depth_timestamp depth_path rgb_timestamp rgb_path`; anything else raises ("Parsing error")."""
    out = []
    for line in text.splitlines():
        if line.startswith("#"):
            continue
        parts = line.split()
        if len(parts) < 4:
            raise ValueError("Parsing error")
        try:
            out.append((float(parts[0]), parts[1], float(parts[2]), parts[3]))
        except ValueError as e:
            raise ValueError("Parsing error") from e
    return out


def parse_trajectory(text: str):
    """tum_rgbd::parse::trajectory (tum_rgbd.rs:102-104): `timestamp tx ty tz qx qy qz qw` -> (ts, [7]) list."""
    out = []
    for line in text.splitlines():
        if line.startswith("#"):
            continue
        v = [float(x) for x in line.split()]
        if len(v) != 8:
            raise ValueError("Parsing error")
        q = np.array(v[4:8])
        q /= np.linalg.norm(q)  # UnitQuaternion::from_quaternion (tum_rgbd.rs:192)
        out.append((v[0], np.concatenate([v[1:4], q])))
    return out


def _rust_float(x, dtype):
    # Rust `{}`: shortest round-trip digits, positional notation, no trailing ".0"
    return np.format_float_positional(dtype(x), unique=True, trim="-")


def frame_to_string(timestamp: float, pose7) -> str:
    """`Frame::to_string` (tum_rgbd.rs:76-86): `timestamp tx ty tz qx qy qz qw`, f64 timestamp and f32 pose."""
    return " ".join([_rust_float(timestamp, np.float64)] + [_rust_float(v, np.float32) for v in pose7])


def read_depth_png(path: str) -> np.ndarray:
    """helper::read_png_16bits (helper.rs:13-36): 16-bit gray PNG -> u16 [rows, cols]."""
    import cv2

    d = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if d is None or d.dtype != np.uint16 or d.ndim != 2:
        raise ValueError(f"{path}: expected a 16-bit gray PNG")
    return d


def read_gray_png(path: str) -> np.ndarray:
    """image::open(path).to_luma() (vors_track.rs:143): gray passes through; RGB -> f32 BT.709 luma, truncated
    (recalled behaviour of image 0.19)."""
    import cv2

    im = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if im is None:
        raise ValueError(f"cannot read {path}")
    if im.dtype == np.uint16:
        im = (im >> 8).astype(np.uint8)
    if im.ndim == 2:
        return im
    b, g, r = [im[..., k].astype(np.float32) for k in range(3)]  # OpenCV stores BGR
    return (np.float32(0.2126) * r + np.float32(0.7152) * g + np.float32(0.0722) * b).astype(np.uint8)


def write_dataset(root: str, frames, timestamps=None, rgb: bool = False) -> str:
    """Write [(gray u8, depth u16), ...] as a TUM-layout dataset (rgb/*.png, depth/*.png, associations.txt in the
    format of examples/README.md:26-31) and return the associations path."""
    import cv2

    os.makedirs(os.path.join(root, "rgb"), exist_ok=True)
    os.makedirs(os.path.join(root, "depth"), exist_ok=True)
    lines = ["# depth_timestamp depth_file_path rgb_timestamp rgb_file_path"]
    for k, (gray, depth) in enumerate(frames):
        ts = timestamps[k] if timestamps is not None else 1305031102.0 + k / 30.0
        dts, cts = ts, ts + 0.011
        dn, cn = f"depth/{dts:.6f}.png", f"rgb/{cts:.6f}.png"
        cv2.imwrite(os.path.join(root, dn), np.ascontiguousarray(depth, np.uint16))
        img = np.repeat(gray[..., None], 3, -1) if rgb else gray
        cv2.imwrite(os.path.join(root, cn), np.ascontiguousarray(img, np.uint8))
        lines.append(f"{dts:.6f} {dn} {cts:.6f} {cn}")
    path = os.path.join(root, "associations.txt")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    return path


def run_tracker(camera_id: str, associations_path: str, out=None, **config_overrides):
    """`my_run` of src/bin/vors_track.rs:26-67 through the Python mirror; returns the trajectory lines."""
    import vors_b200 as vb

    with open(associations_path) as f:
        assoc = parse_associations(f.read())
    parent = os.path.dirname(os.path.abspath(associations_path))
    kw = dict(nb_levels=6, candidates_diff_threshold=7, depth_scale=DEPTH_SCALE, idepth_variance=1e-4, **INTRINSICS[camera_id])
    kw.update(config_overrides)
    cfg = vb.Config(**kw)

    def read(a):
        return read_depth_png(os.path.join(parent, a[1])), read_gray_png(os.path.join(parent, a[3]))

    depth, gray = read(assoc[0])
    tracker = cfg.init(assoc[0][0], depth, assoc[0][2], gray)
    lines = []
    for a in assoc[1:]:
        depth, gray = read(a)
        tracker.track(a[0], depth, a[2], gray)
        ts, pose = tracker.current_frame()
        lines.append(frame_to_string(ts, pose.as_array()))
        if out is not None:
            print(lines[-1], file=out)
    return lines
