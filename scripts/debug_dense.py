import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "visual-odometry-rs_b200"))
import numpy as np
import vors_b200 as vb
from oracle import oracle_py as O
from vors_b200 import synth
scene, f0, f1, pose1 = synth.make_pair(seed=31, rows=240, cols=320)
d = synth.scene_config_kwargs(scene); d.update(nb_levels=4, candidate_mode=1)
for team in (0, 1, 4):
    d["team_size"] = team
    cfg, ocfg = vb.Config(**d), O.default_config(**d)
    kf = vb.Keyframe(cfg, f0[1], f0[0]); okf = O.Keyframe(ocfg, f0[1], f0[0])
    ident = O.Pose.identity()
    e, n, g, H = kf.align_pass(0, f1[0], vb.Pose.identity())
    e64, n64, g64, H64 = okf.eval(0, f1[0], ident, 1)
    e32, n32, g32, H32 = okf.eval(0, f1[0], ident, 0)
    I0 = f0[0].astype(np.float64); I1 = f1[0].astype(np.float64)
    valid = f0[1] != 0
    r = (I1 - I0)[:238, :318][valid[:238, :318]]
    print("team", team, "gpu", e, n, "oracle64", e64, n64, "oracle32", e32, n32, "numpy(strict inside)", (r**2).mean(), r.size)
    print("   relH", np.abs(H-H64).max()/np.abs(H64).max(), "relg", np.abs(g-g64).max()/np.abs(g64).max())
