"""Debug helper: print GPU and oracle LM traces / pass results side by side (run on the GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "visual-odometry-rs_b200"))
import numpy as np
import vors_b200 as vb
from oracle import oracle_py as O
from vors_b200 import synth

def cfgs(scene, **kw):
    d = synth.scene_config_kwargs(scene); d.update(kw)
    return vb.Config(**d), O.default_config(**d)

scene, f0, f1, pose1 = synth.make_pair(seed=31, rows=240, cols=320)
for mode in (0, 1):
    cfg, ocfg = cfgs(scene, nb_levels=4, candidate_mode=mode)
    kf = vb.Keyframe(cfg, f0[1], f0[0]); okf = O.Keyframe(ocfg, f0[1], f0[0])
    pyr1 = O.mean_pyramid(f1[0], 4)
    for l in (3, 1, 0):
        st, out, it, en, tr = kf.align_level(l, pyr1[l], vb.Pose.identity())
        ost, oout, oit, oen, otr = okf.iterative_solve(ocfg, l, pyr1[l], O.Pose.identity())
        print(f"mode {mode} level {l}: gpu it={it} en={en} | oracle it={oit} en={oen} | pose err {O.pose_error(out.as_array(), oout.as_array())}")
        for k in range(max(len(tr), len(otr))):
            a = tr[k] if k < len(tr) else None; b = otr[k] if k < len(otr) else None
            fa = f"{a.iter:2d} E={a.energy:.6f} n={a.n_inside} lam={a.lm_coef:.1e} acc={a.accepted}" if a else "-"
            fb = f"{b.iter:2d} E={b.energy:.6f} n={b.n_inside} lam={b.lm_coef:.1e} acc={b.accepted}" if b else "-"
            print("   ", fa, " | ", fb)

scene, f0, f1, pose1 = synth.make_pair(seed=21, rows=240, cols=320, holes=2)
cfg, ocfg = cfgs(scene, nb_levels=4)
kf = vb.Keyframe(cfg, f0[1], f0[0]); okf = O.Keyframe(ocfg, f0[1], f0[0])
pyr1 = O.mean_pyramid(f1[0], 4)
rng = np.random.default_rng(5)
models = [np.zeros(6), np.concatenate([pose1[0] * 0.5, [0.002, -0.001, 0.003]]), rng.uniform(-0.05, 0.05, 6), np.array([0.5, 0.2, -0.3, 0.1, 0.2, -0.1])]
for l in range(4):
    for mi, xi in enumerate(models):
        m = O.se3_exp(xi); vm = vb.Pose.from_arrays(m.t, m.q)
        e, n, g, H = kf.align_pass(l, pyr1[l], vm)
        e64, n64, g64, H64 = okf.eval(l, pyr1[l], m, 1)
        e32, n32, g32, H32 = okf.eval(l, pyr1[l], m, 0)
        print(f"pass l={l} model={mi}: n {n} / {n64}; E {e:.6f} / {e64:.6f} / f32 {e32:.6f}; relE {abs(e-e64)/abs(e64):.2e} relg {np.abs(g-g64).max()/np.abs(g64).max():.2e} relH {np.abs(H-H64).max()/np.abs(H64).max():.2e} | f32-vs-f64 relE {abs(e32-e64)/abs(e64):.2e} relH {np.abs(H32-H64).max()/np.abs(H64).max():.2e}")
