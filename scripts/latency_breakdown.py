"""Where a single-stream track() call spends its time (BatchTracker of 1, reference configuration)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "visual-odometry-rs_b200"))
import numpy as np
import vors_b200 as vb
from vors_b200 import synth
for (rows, cols, L, kw) in [(480, 640, 6, {}), (480, 640, 5, dict(candidate_mode=1, fixed_iters=10)), (1080, 1920, 6, {})]:
    scene, frames, poses = synth.make_sequence(seed=77, n_frames=16, rows=rows, cols=cols)
    for team in (0, 1, 4, 16, 64):
        cfg = vb.Config(nb_levels=L, team_size=team, **synth.scene_config_kwargs(scene), **kw)
        bt = vb.BatchTracker(cfg, [0.0], frames[0][1][None], [0.0], frames[0][0][None])
        tm, wall, passes = [], [], []
        for k in range(1, 16):
            t0 = time.perf_counter()
            st, stats = bt.track([float(k)], frames[k][1][None], [float(k)], frames[k][0][None])
            wall.append(time.perf_counter() - t0); tm.append(bt.last_timing()); passes.append(stats[0].n_passes)
        med = lambda key: float(np.median([t[key] for t in tm[3:]]))
        print(f"{rows}x{cols} L{L} {kw} team={team}: wall {np.median(wall[3:])*1e3:.3f} ms | upload {med('upload_ms'):.3f} pyramid {med('pyramid_ms'):.3f} align {med('align_ms'):.3f} keyframe {med('keyframe_ms'):.3f} | passes {np.median(passes)} n0 {stats[0].n_points[0]}")
