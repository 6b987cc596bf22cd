import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "visual-odometry-rs_b200"))
import numpy as np
import vors_b200 as vb
from oracle import oracle_py as O
from vors_b200 import synth
scene, frames, poses = synth.make_sequence(seed=77, n_frames=12, rows=480, cols=640)
base = dict(nb_levels=5, candidate_mode=1, fixed_iters=10, **synth.scene_config_kwargs(scene))
t = vb.Config(**base).init(0.0, frames[0][1], 0.0, frames[0][0])
o32 = O.Tracker(O.default_config(**base), 0.0, frames[0][1], 0.0, frames[0][0])
o64 = O.Tracker(O.default_config(**base), 0.0, frames[0][1], 0.0, frames[0][0])
for k in range(1, 12):
    g, d = frames[k]
    st = t.track(float(k), d, float(k), g)
    O.lib().ref_set_accum_f64(0); o32.track(float(k), d, float(k), g)
    O.lib().ref_set_accum_f64(1); o64.track(float(k), d, float(k), g); O.lib().ref_set_accum_f64(0)
    pg, p32, p64 = t.current_frame()[1].as_array(), o32.current_frame()[1].as_array(), o64.current_frame()[1].as_array()
    gt = np.concatenate(poses[k])
    fmt = lambda e: f"({e[0]:.1e} rad, {e[1]:.1e} m)"
    print(k, "gpu-o32", fmt(O.pose_error(pg, p32)), "gpu-o64", fmt(O.pose_error(pg, p64)), "o32-o64", fmt(O.pose_error(p32, p64)),
          "| vs GT: gpu", fmt(O.pose_error(pg, gt)), "o32", fmt(O.pose_error(p32, gt)), "o64", fmt(O.pose_error(p64, gt)), "kf", st.keyframe_changed)
