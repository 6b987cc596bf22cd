"""Per-phase device timers of a single-stream track() (needs a library built with -DVORS_TIMING=1, see build_variants.sh)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "visual-odometry-rs_b200"))
import numpy as np
import vors_b200 as vb
from vors_b200 import synth
rows, cols, L = 480, 640, 6
scene, frames, poses = synth.make_sequence(seed=77, n_frames=6, rows=rows, cols=cols)
cfg = vb.Config(nb_levels=L, **synth.scene_config_kwargs(scene))
bt = vb.BatchTracker(cfg, [0.0], frames[0][1][None], [0.0], frames[0][0][None])
for k in range(1, 6):
    print("frame", k, flush=True)
    bt.track([float(k)], frames[k][1][None], [float(k)], frames[k][0][None])
