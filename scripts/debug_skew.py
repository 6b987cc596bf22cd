import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "visual-odometry-rs_b200"))
import numpy as np
import vors_b200 as vb
from oracle import oracle_py as O
from vors_b200 import synth
for intr in (dict(), dict(skew=0.7), dict(skew=-1.3, fx=300.0), dict(fy=-480.0, fx=481.2, cx=159.5, cy=119.5)):
    for mode in (0, 1):
        scene, f0, f1, _ = synth.make_pair(seed=81, rows=240, cols=320, max_v=0.02, max_w=0.01)
        kw = synth.scene_config_kwargs(scene); kw.update(intr); kw.update(nb_levels=4, candidate_mode=mode)
        cfg, ocfg = vb.Config(**kw), O.default_config(**kw)
        kf = vb.Keyframe(cfg, f0[1], f0[0]); okf = O.Keyframe(ocfg, f0[1], f0[0])
        pyr1 = O.mean_pyramid(f1[0], 4)
        m = O.se3_exp([0.004, -0.002, 0.003, 0.001, 0.002, -0.001])
        for l in (3, 2, 1, 0):
            e, n, g, H = kf.align_pass(l, pyr1[l], vb.Pose.from_arrays(m.t, m.q))
            e64, n64, g64, H64 = okf.eval(l, pyr1[l], m, 1)
            print(intr, "mode", mode, "lvl", l, "n", n, n64, "relE %.2e" % (abs(e - e64) / abs(e64)), "relg %.2e" % (np.abs(g - g64).max() / np.abs(g64).max()),
                  "relH %.2e" % (np.abs(H - H64).max() / np.abs(H64).max()))
