"""Diagnosis (GPU box): one dense 640x480 stream of the bench's own data (bench.make_streams, stream index argv[1]), fixed 10 LM
rounds / level, team_size 1: per frame and level the final energy and iteration decisions of the GPU against the oracle with
f64 sums, and the pose distance.  VORS_NO_TILED=1 selects the generic records for the same comparison."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "visual-odometry-rs_b200"))
import numpy as np, torch
import bench
import vors_b200 as vb
from oracle import oracle_py as O

stream = int(sys.argv[1]) if len(sys.argv) > 1 else 5
n_frames = int(sys.argv[2]) if len(sys.argv) > 2 else 5
cfg = bench.CONFIGS[2]
gray, depth, _, scene = bench.make_streams(cfg, stream + 1, 12, 100000, torch.device("cuda", 0))
gray, depth = gray[:, stream].cpu().numpy(), depth[:, stream].cpu().numpy()
kw = bench.tracker_kwargs(cfg, scene)
t = vb.Config(team_size=1, **kw).init(0.0, depth[0], 0.0, gray[0])
t.set_tracing(True)
O.lib().ref_set_accum_f64(1)
ot = O.Tracker(O.default_config(**kw), 0.0, depth[0], 0.0, gray[0])
for k in range(1, n_frames + 1):
    g, d = gray[k], depth[k]
    st = t.track(float(k), d, float(k), g)
    _, ost, otr = ot.track(float(k), d, float(k), g, trace_cap=512)
    tr = t.last_trace()
    ang, dist = O.pose_error(t.current_frame()[1].as_array(), ot.current_frame()[1].as_array())
    same = sum(int((a.level, a.iter, a.accepted) == (b.level, b.iter, b.accepted)) for a, b in zip(tr, otr))
    line = f"frame {k}: pose diff {ang:.2e} rad {dist:.2e} m; flow {st.optical_flow:.3f}/{ost.optical_flow:.3f}; decisions {same}/{len(otr)};"
    for l in range(4, -1, -1):
        ge, oe = st.energy[l], ost.energy[l]
        line += f" L{l} E {ge:.5f}/{oe:.5f} ({(ge-oe)/oe:+.1e})"
    print(line)
    if os.environ.get("DIAG_LEVEL") and k == int(os.environ.get("DIAG_FRAME", "3")):
        L = int(os.environ["DIAG_LEVEL"])
        for a, b in zip(tr, otr):
            if b.level == L or a.level == L:
                print(f"   L{a.level}/{b.level} it {a.iter}/{b.iter}: gpu E {a.energy:.6f} n {a.n_inside} acc {a.accepted} lam {a.lm_coef:g} | oracle E {b.energy:.6f} n {b.n_inside} acc {b.accepted} lam {b.lm_coef:g}")
    shown = 0
    for a, b in zip(tr, otr):
        if (a.level, a.iter, a.accepted) != (b.level, b.iter, b.accepted) or abs(a.energy - b.energy) > 1e-5 * abs(b.energy) or a.n_inside != b.n_inside:
            print(f"   diff: level {b.level} iter {b.iter}: gpu E {a.energy:.6f} n {a.n_inside} acc {a.accepted} lam {a.lm_coef:g} | oracle E {b.energy:.6f} n {b.n_inside} acc {b.accepted} lam {b.lm_coef:g}")
            shown += 1
            if shown >= 6:
                break
