"""Diagnosis (GPU box): 296 replicas of one bench stream for a few steps: every replica must produce bit-identical poses."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "visual-odometry-rs_b200"))
import numpy as np, torch
import bench
import vors_b200 as vb
cfg = bench.CONFIGS[2]
N = int(sys.argv[1]) if len(sys.argv) > 1 else 296
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 5
gray, depth, _, scene = bench.make_streams(cfg, 1, 6, 424200, torch.device("cuda", 0))
g = gray[:, 0].cpu().numpy(); d = depth[:, 0].cpu().numpy()
kw = bench.tracker_kwargs(cfg, scene)
ts = np.zeros(N)
bt = vb.BatchTracker(vb.Config(device=0, **kw), ts, np.repeat(d[:1], N, 0), ts, np.repeat(g[:1], N, 0), layout=vb.ROW_MAJOR)
for k in range(1, STEPS + 1):
    st, stats = bt.track(ts + k, np.repeat(d[k:k+1], N, 0), ts + k, np.repeat(g[k:k+1], N, 0))
    _, poses = bt.current_frames()
    same = np.all(poses == poses[0], axis=1)
    print(f"step {k}: {int(same.sum())}/{N} replicas identical to replica 0; team {bt.last_launch_shape()}; max |diff| {np.abs(poses - poses[0]).max():.2e}; passes {sorted(set(s.n_passes for s in stats))}")
