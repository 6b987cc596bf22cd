"""Generate the frozen golden vectors under tests/golden/ FROM THE CPU ORACLE (oracle/vors_oracle.cpp).

The reference ships no fixtures for this path and cannot be run here (Rust toolchain absent), so these vectors pin
the ORACLE's behaviour ("parity unpinned" against the reference itself, see oracle/vors_oracle.h): seeded synthetic
RGB-D inputs -> pyramid / gradient checksums, the level-0 candidate mask, candidate counts, the LM decision trace and
the final pose, for the reference configuration (coarse-to-fine candidates, adaptive LM) and the benchmark
configuration (dense candidates, 10 fixed rounds).

    python tests/golden/make_golden.py        # rewrites tests/golden/pair_96x128.npz
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "visual-odometry-rs_b200"))

from oracle import oracle_py as O  # noqa: E402
from vors_b200 import synth  # noqa: E402

ROWS, COLS, LEVELS, SEED = 96, 128, 4, 4242


def crc(a):
    return np.uint32(zlib.crc32(np.ascontiguousarray(a).tobytes()))


def build():
    scene, f0, f1, pose1 = synth.make_pair(seed=SEED, rows=ROWS, cols=COLS, max_v=0.03, max_w=0.02, holes=2)
    out = dict(gray0=f0[0], depth0=f0[1], gray1=f1[0], depth1=f1[1], gt_pose=np.concatenate(pose1),
               intrinsics=np.array([scene.fx, scene.fy, scene.cx, scene.cy, 0.0], np.float64), levels=np.int32(LEVELS))
    pyr = O.mean_pyramid(f0[0], LEVELS)
    gx, gy, g2 = O.gradients_tracker(pyr)
    out["pyr_crc"] = np.array([crc(p) for p in pyr])
    out["gx_crc"] = np.array([crc(g) for g in gx])
    out["gy_crc"] = np.array([crc(g) for g in gy])
    out["g2_crc"] = np.array([crc(g) for g in g2])
    for name, kw in (("c2f", dict()), ("dense", dict(candidate_mode=1, fixed_iters=10))):
        cfg = O.default_config(nb_levels=LEVELS, **synth.scene_config_kwargs(scene), **kw)
        tr = O.Tracker(cfg, 0.0, f0[1], 0.0, f0[0])
        kf = tr.keyframe()
        out[f"{name}_mask0"] = np.packbits(kf.mask0())
        out[f"{name}_n_points"] = np.array([kf.n_points(l) for l in range(LEVELS)], np.int32)
        st, stats, trace = tr.track(1.0, f1[1], 1.0, f1[0], trace_cap=512)
        out[f"{name}_status"] = np.int32(st)
        out[f"{name}_trace"] = np.array([(r.level, r.iter, r.energy, r.n_inside, r.lm_coef, r.accepted) for r in trace], np.float64)
        out[f"{name}_pose"] = tr.current_frame()[1].as_array()
        out[f"{name}_flow"] = np.float32(stats.optical_flow)
        out[f"{name}_keyframe_changed"] = np.int32(stats.keyframe_changed)
    return out


if __name__ == "__main__":
    O.build()
    data = build()
    path = os.path.join(HERE, "pair_96x128.npz")
    np.savez_compressed(path, **data)
    print("wrote", path, os.path.getsize(path), "bytes")
