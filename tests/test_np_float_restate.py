"""The oracle's FLOAT path against a second, independent restatement (tests/np_restate.py, numpy float32, written from the
reference's Rust and nalgebra 0.17's arithmetic): bit-for-bit on Jacobians, warp, interpolation, one evaluation (E, g, H with
the reference's sequential f32 sums), the damped Cholesky solve, se3::exp, and decision-for-decision on a whole LM trace.
This is what pins rows J, M, N, O, P of SURVEY.md §8a beyond a single restatement (the Rust reference cannot run here)."""
import numpy as np
import pytest

import np_restate as R
from vors_b200 import synth

F = np.float32


def _obs(oracle, okf, lvl, image_pyr):
    xy, idepth, jac = okf.points(lvl)
    tmpl = okf.image(lvl)
    pyr0 = [okf.image(l) for l in range(okf.levels)]
    gx, gy, _ = oracle.gradients_tracker(pyr0)
    xy = xy.astype(np.int64)
    intr = okf.intrinsics(lvl)  # fx fy cx cy skew
    my_jac = R.warp_jacobians(intr, xy, idepth, gx[lvl][xy[:, 1], xy[:, 0]], gy[lvl][xy[:, 1], xy[:, 0]])
    return dict(xy=xy, idepth=idepth, intr=intr, image=image_pyr[lvl], template=tmpl, jac=my_jac), jac


def _iso(p):
    return np.array(list(p.t), F), np.array(list(p.q), F)


@pytest.fixture(scope="module")
def pair(oracle):
    scene, f0, f1, pose1 = synth.make_pair(seed=3100, rows=96, cols=128, holes=1)
    kw = dict(nb_levels=3, **synth.scene_config_kwargs(scene))
    cfg = oracle.default_config(**kw)
    okf = oracle.Keyframe(cfg, f0[1], f0[0])
    pyr1 = oracle.mean_pyramid(f1[0], 3)
    return cfg, okf, pyr1, pose1


def test_jacobians_bit_exact(oracle, pair):
    cfg, okf, pyr1, _ = pair
    for lvl in range(3):
        obs, ojac = _obs(oracle, okf, lvl, pyr1)
        assert obs["jac"].shape == ojac.shape and np.array_equal(obs["jac"], ojac), f"level {lvl}"
    # with skew and a negative focal length (ICL intrinsics have fy < 0, src/dataset/tum_rgbd.rs:23-27)
    rng = np.random.default_rng(4)
    for intr in ([481.2, -480.0, 319.5, 239.5, 0.0], [500.0, 510.0, 300.0, 250.0, 1.5]):
        intr = np.array(intr, F)
        xy = rng.integers(0, 600, (200, 2))
        rho = rng.uniform(0.1, 3.0, 200).astype(F)
        gx, gy = rng.integers(-255, 256, 200), rng.integers(-255, 256, 200)
        mine = R.warp_jacobians(intr, xy, rho, gx, gy)
        for i in range(200):
            out = np.zeros(6, F)
            oracle.lib().ref_warp_jacobian_at(float(gx[i]), float(gy[i]), float(xy[i, 0]), float(xy[i, 1]), float(rho[i]), intr, out)
            assert np.array_equal(out, mine[i]), (i, out, mine[i])


def test_warp_and_interpolate_bit_exact(oracle, pair):
    import ctypes as C
    cfg, okf, pyr1, pose1 = pair
    rng = np.random.default_rng(11)
    intr = okf.intrinsics(0)
    img = pyr1[0]
    img_cm = np.ascontiguousarray(img.T).reshape(-1)
    for trial in range(4):
        m = oracle.se3_exp(rng.uniform(-0.08, 0.08, 6))
        model = _iso(m)
        xy = np.stack([rng.integers(0, img.shape[1], 300), rng.integers(0, img.shape[0], 300)], 1)
        rho = rng.uniform(0.2, 2.0, 300).astype(F)
        u, v = R.warp(model, xy[:, 0], xy[:, 1], rho, intr)
        inside, val = R.interpolate(u, v, img)
        for i in range(300):
            uv = np.zeros(2, F)
            oracle.lib().ref_warp(C.byref(m), float(xy[i, 0]), float(xy[i, 1]), float(rho[i]), intr, uv)
            assert uv[0] == u[i] and uv[1] == v[i], (i, uv, u[i], v[i])
            o = C.c_float()
            ok = oracle.lib().ref_interpolate(float(u[i]), float(v[i]), img_cm, img.shape[0], img.shape[1], C.byref(o))
            assert bool(ok) == bool(inside[i])
            if ok:
                assert F(o.value) == val[i]


def test_one_evaluation_bit_exact(oracle, pair):
    """eval_energy + compute_eval_data (lm_optimizer.rs:68-107) with the reference's sequential f32 sums."""
    cfg, okf, pyr1, pose1 = pair
    rng = np.random.default_rng(7)
    models = [np.zeros(6), np.concatenate([pose1[0] * 0.5, [0.002, -0.001, 0.003]]), rng.uniform(-0.05, 0.05, 6),
              np.array([0.3, 0.1, -0.2, 0.1, 0.15, -0.1])]  # the last one pushes many candidates outside
    for lvl in range(3):
        obs, _ = _obs(oracle, okf, lvl, pyr1)
        for xi in models:
            m = oracle.se3_exp(xi)
            e, n, g, H = okf.eval(lvl, pyr1[lvl], m, 0)
            pre = R.eval_energy(obs, _iso(m))
            ev = R.compute_eval_data(obs, _iso(m), pre)
            assert n == len(pre[1])
            assert F(e) == ev["energy"] or (np.isnan(e) and np.isnan(ev["energy"])), (lvl, e, ev["energy"])
            assert np.array_equal(g, ev["gradient"]), (lvl, g, ev["gradient"])
            assert np.array_equal(H, ev["hessian"]), lvl


def test_se3_exp_and_cholesky_bit_exact(oracle):
    rng = np.random.default_rng(3)
    for i in range(200):
        xi = rng.uniform(-0.5, 0.5, 6) * (1e-3 if i % 3 == 0 else 1.0)  # both branches of the Taylor switch
        p = oracle.se3_exp(xi)
        t, q = R.se3_exp(xi)
        assert np.array_equal(np.array(list(p.t), F), t) and np.array_equal(np.array(list(p.q), F), q), (i, xi)
    for i in range(100):
        A = rng.normal(size=(6, 8)).astype(F)
        H = (A @ A.T).astype(F)
        H = ((H + H.T) * F(0.5)).astype(F)
        g = rng.normal(size=6).astype(F)
        x = np.zeros(6, F)
        ok = oracle.lib().ref_cholesky_solve6(np.ascontiguousarray(H.reshape(-1)), g, x)
        mine = R.cholesky_solve6(H, g)
        assert ok == 1 and mine is not None and np.array_equal(x, mine), i
    bad = np.zeros((6, 6), F)  # zero Hessian (no candidate inside): the decomposition must fail
    assert R.cholesky_solve6(bad, np.ones(6, F)) is None
    assert oracle.lib().ref_cholesky_solve6(bad.reshape(-1), np.ones(6, F), np.zeros(6, F)) == 0


@pytest.mark.parametrize("fixed_iters", [0, 6])
def test_lm_trace_decision_for_decision(oracle, pair, fixed_iters):
    """A whole `iterative_solve` per level (math/optimizer.rs:57-70, lm_optimizer.rs:113-192): every energy, inside count,
    damping coefficient and accept / reject decision, and the final model, bit-for-bit."""
    cfg, okf, pyr1, _ = pair
    cfg.fixed_iters = fixed_iters
    model = oracle.Pose.identity()
    for lvl in (2, 1, 0):
        obs, _ = _obs(oracle, okf, lvl, pyr1)
        st, out, n_iter, energy, trace = okf.iterative_solve(cfg, lvl, pyr1[lvl], model)
        mst, mout, mn_iter, mtrace = R.iterative_solve(obs, _iso(model), fixed_iters)
        assert st == mst == 0
        assert len(trace) == len(mtrace) and n_iter == mn_iter, (lvl, len(trace), len(mtrace))
        for a, b in zip(trace, mtrace):
            assert (a.iter, a.n_inside, a.accepted) == (b[0], b[2], b[4]), (lvl, a.iter)
            assert F(a.energy) == b[1] and F(a.lm_coef) == b[3], (lvl, a.iter, a.energy, b[1])
        assert np.array_equal(np.array(list(out.t), F), mout[0]) and np.array_equal(np.array(list(out.q), F), mout[1]), lvl
        model = out
    cfg.fixed_iters = 0


def test_tracker_sequence_bit_exact(oracle):
    """Tracker::track (inverse_compositional.rs:170-240) over a short sequence with keyframe switches: prior, level loop, pose
    composition, optical flow and the keyframe decision, against the numpy restatement driving the same keyframe data."""
    scene, frames, _ = synth.make_sequence(seed=4200, n_frames=6, rows=96, cols=128, step_v=0.02, step_w=0.01)
    kw = dict(nb_levels=3, **synth.scene_config_kwargs(scene))
    cfg = oracle.default_config(**kw)
    tr = oracle.Tracker(cfg, 0.0, frames[0][1], 0.0, frames[0][0])
    okf = oracle.Keyframe(cfg, frames[0][1], frames[0][0])
    ident = (np.zeros(3, F), np.array([0, 0, 0, 1], F))
    state = dict(keyframe_pose=ident, current_frame_pose=ident)
    switches = 0
    for k in range(1, 6):
        gray, depth = frames[k]
        _, st, _ = tr.track(float(k), depth, float(k), gray)
        pyr = oracle.mean_pyramid(gray, 3)
        levels = [_obs(oracle, okf, l, pyr)[0] for l in range(3)]
        ok, flow = R.tracker_track(state, levels, levels[-1])
        assert ok == (st.status == 0)
        assert F(st.optical_flow) == flow, (k, st.optical_flow, flow)
        p = tr.current_frame()[1]
        assert np.array_equal(np.array(list(p.t), F), state["current_frame_pose"][0]), k
        assert np.array_equal(np.array(list(p.q), F), state["current_frame_pose"][1]), k
        change = bool(flow >= F(1.0))
        assert change == bool(st.keyframe_changed)
        if change:
            switches += 1
            okf = oracle.Keyframe(cfg, depth, gray)
            state["keyframe_pose"] = state["current_frame_pose"]
    assert switches >= 1
