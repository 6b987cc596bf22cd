"""Diagnosis (GPU box): one evaluation of level LVL of frame F of bench stream S at the oracle's own level-(LVL+1) model,
tiled GPU records against the oracle; on a mismatch the keyframe depth is masked to sub-rectangles of tiles to localise it."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "visual-odometry-rs_b200"))
import numpy as np, torch
import bench
import vors_b200 as vb
from oracle import oracle_py as O

S, F, LVL = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
cfg = bench.CONFIGS[2]
gray, depth, _, scene = bench.make_streams(cfg, S + 1, 12, 100000, torch.device("cuda", 0))
gray, depth = gray[:, S].cpu().numpy(), depth[:, S].cpu().numpy()
kw = bench.tracker_kwargs(cfg, scene)
ocfg = O.default_config(**kw)
O.lib().ref_set_accum_f64(1)
ot = O.Tracker(ocfg, 0.0, depth[0], 0.0, gray[0])
for k in range(1, F):
    ot.track(float(k), depth[k], float(k), gray[k])
model = O.pose_mul(O.pose_inverse(ot.current_frame()[1]), ot.keyframe_pose())
okf = O.Keyframe(ocfg, depth[0], gray[0])
pyr = O.mean_pyramid(gray[F], 5)
for l in range(4, LVL, -1):
    st, model, _, _, _ = okf.iterative_solve(ocfg, l, pyr[l], model)
vm = vb.Pose.from_arrays(model.t, model.q)

def compare(dmask, tag):
    d = np.where(dmask, depth[0], 0).astype(np.uint16)
    kf = vb.Keyframe(vb.Config(**kw), d, gray[0])
    okf2 = O.Keyframe(ocfg, d, gray[0])
    e, n, g, H = kf.align_pass(LVL, pyr[LVL], vm)
    eo, no, go, Ho = okf2.eval(LVL, pyr[LVL], model, 1)
    bad = (n != no) or abs(e - eo) > 1e-5 * abs(eo)
    print(f"{tag}: gpu E {e:.6f} n {n} | oracle E {eo:.6f} n {no} | sumr2 diff {e*n - eo*no:+.2f} {'MISMATCH' if bad else 'ok'}")
    return bad

full = np.ones(depth[0].shape, bool)
if compare(full, "full"):
    sc = 2 ** LVL
    R, C = depth[0].shape
    th, tw = 32 * sc, 12 * sc  # one level-LVL tile in level-0 pixels
    # by tile row, then by tile column
    rows = range(0, R, th); cols = range(0, C, tw)
    for r0 in rows:
        m = np.zeros_like(full); m[r0:r0 + th, :] = True
        if compare(m, f"tile row y0={r0//sc}"):
            for c0 in cols:
                m2 = np.zeros_like(full); m2[r0:r0 + th, c0:c0 + tw] = True
                compare(m2, f"   tile y0={r0//sc} x0={c0//sc}")

# ---- whole level loop in one launch (vors_align) from the same prior, against per-level oracle solves
kf = vb.Keyframe(vb.Config(**kw), depth[0], gray[0])
prior = O.pose_mul(O.pose_inverse(ot.current_frame()[1]), ot.keyframe_pose())
st, out, stats, trace = kf.align(gray[F], vb.Pose.from_arrays(prior.t, prior.q), trace_cap=512)
m = prior
for l in range(4, -1, -1):
    sto, m2, nit, en, otr = okf.iterative_solve(ocfg, l, pyr[l], m)
    gtr = [r for r in trace if r.level == l]
    print(f"level {l}: iter0 gpu E {gtr[0].energy:.6f} n {gtr[0].n_inside} | oracle E {otr[0].energy:.6f} n {otr[0].n_inside}; final gpu {stats.energy[l]:.6f} oracle {en:.6f}")
    m = m2
# ---- one level at a time (vors_align_level), fed with the oracle's models
m = prior
for l in range(4, -1, -1):
    sto, m2, nit, en, otr = okf.iterative_solve(ocfg, l, pyr[l], m)
    stg, outg, nitg, eng, gtr = kf.align_level(l, pyr[l], vb.Pose.from_arrays(m.t, m.q))
    print(f"level {l} alone: iter0 gpu E {gtr[0].energy:.6f} n {gtr[0].n_inside} | oracle E {otr[0].energy:.6f} n {otr[0].n_inside}; final gpu {eng:.6f} oracle {en:.6f}")
    m = m2
