"""The C++ oracle against an independent numpy restatement (integer stages, bit-exact) and against
self-consistency properties (Jacobian vs finite differences of the warp, pose recovery, tracker semantics)."""
import ctypes as C

import numpy as np
import pytest

import np_restate as NP
from vors_b200 import synth

SHAPES = [(48, 64), (37, 53), (2, 2), (3, 7), (64, 33), (480, 640)]


def _img(rng, shape, kind):
    if kind == "noise":
        return rng.integers(0, 256, shape, dtype=np.uint8)
    if kind == "const":
        return np.full(shape, 1, np.uint8)  # benches/mean_pyramid.rs:10
    if kind == "extreme":
        return (rng.integers(0, 2, shape) * 255).astype(np.uint8)
    yy, xx = np.mgrid[0:shape[0], 0:shape[1]]
    return ((np.sin(xx / 5.0) + np.cos(yy / 7.0)) * 60 + 128).astype(np.uint8)


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("kind", ["noise", "const", "extreme", "smooth"])
def test_mean_pyramid_and_gradients_bit_exact(oracle, shape, kind):
    rng = np.random.default_rng(hash((shape, kind)) % 2 ** 32)
    img = _img(rng, shape, kind)
    for L in (1, 3, 6, 12):
        pyr = oracle.mean_pyramid(img, L)
        ref = NP.mean_pyramid(img, L)
        assert len(pyr) == len(ref) == len(oracle.pyramid_shapes(*shape, L))
        for a, b in zip(pyr, ref):
            assert np.array_equal(a, b)
    pyr = oracle.mean_pyramid(img, 6)
    if min(pyr[0].shape) >= 2:
        gx, gy, g2 = oracle.gradients_tracker(pyr)
        rx, ry, r2 = NP.gradients_tracker(pyr)
        for l in range(len(pyr)):
            assert np.array_equal(gx[l], rx[l]) and np.array_equal(gy[l], ry[l]) and np.array_equal(g2[l], r2[l])
        # squared norm never wraps u16
        assert max(int(g.max()) for g in g2) <= 65025


def test_centered_gradient_truncates_toward_zero(oracle):
    img = np.zeros((3, 3), np.uint8)
    img[1, 0], img[1, 2] = 4, 1  # (1 - 4) / 2 = -1 (trunc), floor would give -2
    img[0, 1], img[2, 1] = 0, 5  # 5 / 2 = 2
    gx = np.zeros(9, np.int16)
    gy = np.zeros(9, np.int16)
    oracle.lib().ref_gradient_centered(np.ascontiguousarray(img.T).reshape(-1), 3, 3, gx, gy)
    assert gx.reshape(3, 3).T[1, 1] == -1 and gy.reshape(3, 3).T[1, 1] == 2
    assert np.count_nonzero(gx) == 1 and np.count_nonzero(gy) == 1  # border stays 0


@pytest.mark.parametrize("shape", [(48, 64), (37, 53), (96, 130), (480, 640)])
@pytest.mark.parametrize("thresh", [0, 7, 300])
def test_coarse_to_fine_bit_exact(oracle, shape, thresh):
    rng = np.random.default_rng(shape[0] * 1000 + thresh)
    for kind in ("noise", "smooth", "const"):
        img = _img(rng, shape, kind)
        pyr = oracle.mean_pyramid(img, 5)
        _, _, g2 = oracle.gradients_tracker(pyr)
        masks = oracle.c2f_select(thresh, g2)
        ref = NP.c2f_select(thresh, g2)
        for a, b in zip(masks, ref):
            assert np.array_equal(a, b)
        # coarsest all true; every selected parent yields 1 or 2 children; unselected parents none
        assert masks[-1].all()
        for l in range(len(masks) - 1):
            hr, hc = masks[l + 1].shape
            m = masks[l][:2 * hr, :2 * hc]
            cnt = m[0::2, 0::2].astype(int) + m[1::2, 0::2] + m[0::2, 1::2] + m[1::2, 1::2]
            assert np.all((cnt >= 1) & (cnt <= 2) | ~masks[l + 1])
            assert np.all(cnt[~masks[l + 1]] == 0)
            assert not masks[l][2 * hr:, :].any() and not masks[l][:, 2 * hc:].any()


def test_example_gradient_recipe_differs_from_tracker_recipe(oracle):
    # squared_norm_direct divides the un-truncated sum by 4
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (24, 32), dtype=np.uint8)
    out = np.zeros(img.size, np.uint16)
    oracle.lib().ref_squared_norm_direct(np.ascontiguousarray(img.T).reshape(-1), 24, 32, out)
    m = img.astype(np.int32)
    ref = np.zeros(img.shape, np.int32)
    ref[1:-1, 1:-1] = ((m[1:-1, 2:] - m[1:-1, :-2]) ** 2 + (m[2:, 1:-1] - m[:-2, 1:-1]) ** 2) // 4
    assert np.array_equal(out.reshape(32, 24).T, ref.astype(np.uint16))


def _cfg(oracle, scene, **kw):
    d = synth.scene_config_kwargs(scene)
    d.update(kw)
    return oracle.default_config(**d)


@pytest.fixture(scope="module")
def small_pair():
    return synth.make_pair(seed=7, rows=120, cols=160, max_v=0.02, max_w=0.01)


def test_idepth_pyramid_and_extract_order(oracle, small_pair):
    scene, f0, f1, _ = small_pair
    depth = f0[1].copy()
    depth[10:30, 20:50] = 0  # unknown depth hole
    cfg = _cfg(oracle, scene, nb_levels=4)
    kf = oracle.Keyframe(cfg, depth, f0[0])
    mask0 = kf.mask0()
    d0, w0 = kf.idepth_map(0)
    known = mask0 & (depth != 0)
    assert np.array_equal(~np.isnan(d0), known)
    assert np.allclose(d0[known], np.float32(5000.0) / depth[known].astype(np.float32), rtol=0, atol=0)
    assert np.all(w0[known] == np.float32(1e-4))
    for l in range(1, 4):
        dl, wl = kf.idepth_map(l)
        dp, wp = kf.idepth_map(l - 1)
        hr, hc = dl.shape
        kids_d = np.stack([dp[0:2 * hr:2, 0:2 * hc:2], dp[1:2 * hr:2, 0:2 * hc:2], dp[0:2 * hr:2, 1:2 * hc:2], dp[1:2 * hr:2, 1:2 * hc:2]])
        kids_w = np.stack([wp[0:2 * hr:2, 0:2 * hc:2], wp[1:2 * hr:2, 0:2 * hc:2], wp[0:2 * hr:2, 1:2 * hc:2], wp[1:2 * hr:2, 1:2 * hc:2]])
        assert np.array_equal(~np.isnan(dl), (~np.isnan(kids_d)).any(0))
        wsum = kids_w.sum(0)
        mean = np.nansum(kids_d.astype(np.float64) * kids_w, 0) / np.where(wsum > 0, wsum, 1)
        k = ~np.isnan(dl)
        assert np.allclose(dl[k], mean[k], rtol=2e-6)
        assert np.allclose(wl[k], wsum[k], rtol=1e-6)
    for l in range(4):
        xy, idepth, jac = kf.points(l)
        dl, _ = kf.idepth_map(l)
        # extract_z: column-major scan, (x = col, y = row)
        cols, rows = np.nonzero(~np.isnan(dl.T))
        assert np.array_equal(xy[:, 0], cols) and np.array_equal(xy[:, 1], rows)
        assert np.array_equal(idepth, dl[rows, cols])


def test_idepth_pyramid_bit_for_bit_against_numpy_restatement(oracle, small_pair):
    """Rows F, G, H: the oracle's inverse-depth pyramid (from_depth, halve + fuse + strategy_dso_mean) and extract_z against
    the independent f32 numpy restatement in np_restate.py - every value and weight bit for bit, on a depth map with holes whose
    odd borders produce 1-, 2-, 3- and 4-child blocs, for the Tracker's coarse-to-fine mask and for the dense extension."""
    import np_restate as R

    scene, f0, _, _ = small_pair
    depth = f0[1].copy()
    depth[10:31, 20:51] = 0
    depth[40, 100] = depth[40, 101] = depth[41, 100] = 0
    depth[77:, 3] = 0
    for mode in (0, 1):
        cfg = _cfg(oracle, scene, nb_levels=5, candidate_mode=mode)
        kf = oracle.Keyframe(cfg, depth, f0[0])
        mask = kf.mask0() if mode == 0 else np.ones(depth.shape, bool)
        levels = R.idepth_pyramid(depth, mask, cfg.depth_scale, cfg.idepth_variance, 5)
        assert len(levels) == kf.levels == 5
        child_counts = set()
        for l, (d, v) in enumerate(levels):
            od, ov = kf.idepth_map(l)
            assert np.array_equal(np.isnan(od), np.isnan(d)), (mode, l)
            assert np.array_equal(od.view(np.uint32)[~np.isnan(d)], d.view(np.uint32)[~np.isnan(d)]), (mode, l)
            assert np.array_equal(ov.view(np.uint32), v.view(np.uint32)), (mode, l)
            xy, z = R.extract_z(d)
            oxy, oz, _ = kf.points(l, with_jac=False)
            assert np.array_equal(oxy, xy) and np.array_equal(oz.view(np.uint32), z.view(np.uint32)), (mode, l)
            if l:
                pd = levels[l - 1][0]
                hr, hc = d.shape
                child_counts |= set(np.unique(sum((~np.isnan(pd[a:2 * hr:2, b:2 * hc:2])).astype(int) for a in (0, 1) for b in (0, 1))))
        assert child_counts >= {0, 1, 2, 3, 4} if mode == 1 else child_counts >= {0, 1, 2}, child_counts


def _similar_numpy(ds, vs):
    """inverse_depth.rs:105-152 restated independently (f32, the reference's operation order); returns (d, v) or None."""
    f = np.float32
    ds, vs = [f(x) for x in ds], [f(x) for x in vs]
    n = len(ds)
    if n == 1:
        return ds[0], f(2.0) * vs[0]
    if n == 2:
        nd = (ds[0] * vs[1] + ds[1] * vs[0]) / (vs[0] + vs[1])
        nv = (vs[0] + vs[1]) / f(2.0)
    elif n == 3:
        v12, v13, v23 = vs[0] * vs[1], vs[0] * vs[2], vs[1] * vs[2]
        nd = (ds[0] * v23 + ds[1] * v13 + ds[2] * v12) / (v12 + v13 + v23)
        nv = f(2.0) * (vs[0] + vs[1] + vs[2]) / f(9.0)
    else:
        v123, v234, v341, v412 = vs[0] * vs[1] * vs[2], vs[1] * vs[2] * vs[3], vs[2] * vs[3] * vs[0], vs[3] * vs[0] * vs[1]
        nd = (ds[0] * v234 + ds[1] * v341 + ds[2] * v412 + ds[3] * v123) / (v123 + v234 + v341 + v412)
        nv = (vs[0] + vs[1] + vs[2] + vs[3]) / f(8.0)
    return (nd, nv) if all((d - nd) * (d - nd) < nv for d in ds) else None


def test_statistically_similar_fusion_matches_independent_restatement(oracle, small_pair):
    """The reference's other fusion strategy (inverse_depth.rs:105-152), selectable with idepth_fusion = 1: a bloc whose known
    children disagree by more than one fused standard deviation is Discarded (not a candidate).  Checked bloc by bloc against
    a numpy restatement on a scene with a depth hole and a depth step (so that 1-, 2-, 3-, 4-child and discarded blocs occur)."""
    scene, f0, _, _ = small_pair
    depth = f0[1].copy()
    depth[10:31, 20:51] = 0                                    # hole with odd borders: 2- and 3-child blocs
    depth[40, 100] = depth[40, 101] = depth[41, 100] = 0       # a bloc with a single known child
    depth[60:, 81:] = (depth[60:, 81:].astype(np.float32) * 1.6).astype(np.uint16)  # depth step: dissimilar blocs
    cfg = _cfg(oracle, scene, nb_levels=4, candidate_mode=1, idepth_fusion=1, idepth_variance=1e-4)
    kf = oracle.Keyframe(cfg, depth, f0[0])
    seen = {1: 0, 2: 0, 3: 0, 4: 0, "discarded": 0}
    for l in range(1, 4):
        dl, wl = kf.idepth_map(l)
        dp, wp = kf.idepth_map(l - 1)
        for r in range(dl.shape[0]):
            for c in range(dl.shape[1]):
                kids = [(dp[2 * r + dr, 2 * c + dc], wp[2 * r + dr, 2 * c + dc]) for dc, dr in ((0, 0), (0, 1), (1, 0), (1, 1))]  # a, b, c, d
                kids = [k for k in kids if not np.isnan(k[0])]
                if not kids:
                    assert np.isnan(dl[r, c])
                    continue
                want = _similar_numpy([k[0] for k in kids], [k[1] for k in kids])
                if want is None:
                    assert np.isnan(dl[r, c]) and wl[r, c] == 0
                    seen["discarded"] += 1
                else:
                    assert dl[r, c] == want[0] and wl[r, c] == want[1], (l, r, c)
                    seen[len(kids)] += 1
    assert all(v > 0 for v in seen.values()), seen
    # dso_mean (the Tracker's strategy) never discards: more candidates above level 0
    kf0 = oracle.Keyframe(_cfg(oracle, scene, nb_levels=4, candidate_mode=1), depth, f0[0])
    assert kf.n_points(1) < kf0.n_points(1)


def test_scharr_option_of_the_oracle(oracle):
    """gradient_operator = 1 (an extension; the reference has no Scharr): the oracle's operator against a numpy restatement,
    including the truncating division and the zero border, on noise and on a ramp (where Scharr equals the centred difference)."""
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (37, 53), dtype=np.uint8)
    gx, gy = oracle.gradient_scharr(img)
    I = img.astype(np.int32)
    sx = 3 * (I[:-2, 2:] - I[:-2, :-2]) + 10 * (I[1:-1, 2:] - I[1:-1, :-2]) + 3 * (I[2:, 2:] - I[2:, :-2])
    sy = 3 * (I[2:, :-2] - I[:-2, :-2]) + 10 * (I[2:, 1:-1] - I[:-2, 1:-1]) + 3 * (I[2:, 2:] - I[:-2, 2:])
    trunc = lambda v: (np.sign(v) * (np.abs(v) // 32)).astype(np.int16)
    wx = np.zeros_like(gx); wy = np.zeros_like(gy)
    wx[1:-1, 1:-1] = trunc(sx); wy[1:-1, 1:-1] = trunc(sy)
    assert np.array_equal(gx, wx) and np.array_equal(gy, wy)
    ramp = (np.arange(53, dtype=np.int32)[None, :] * 2 + np.arange(37, dtype=np.int32)[:, None]).astype(np.uint8)
    rx, ry = oracle.gradient_scharr(ramp)
    assert np.all(rx[1:-1, 1:-1] == 2) and np.all(ry[1:-1, 1:-1] == 1)
    assert not rx[0].any() and not rx[:, 0].any() and not ry[-1].any() and not ry[:, -1].any()


def test_huber_option_of_the_oracle(oracle, small_pair):
    """huber_delta (an extension; the reference is plain L2): a huge delta reproduces the L2 evaluation exactly, a small
    one lowers the energy and shrinks g and H (weights <= 1), and H stays symmetric; the default is L2."""
    scene, f0, f1, _ = small_pair
    img1 = f1[0].copy()
    img1[30:60, 50:90] = 255 - img1[30:60, 50:90]  # occluder: large residuals
    m = oracle.se3_exp([0.003, -0.002, 0.002, 0.001, 0.002, -0.001])
    out = {}
    for hd in (0.0, 1e9, 6.0):
        kf = oracle.Keyframe(_cfg(oracle, scene, nb_levels=3, huber_delta=hd), f0[1], f0[0])
        out[hd] = [kf.eval(0, img1, m, acc) for acc in (0, 1)]
    assert oracle.default_config().huber_delta == 0.0
    for acc in (0, 1):
        e0, n0, g0, H0 = out[0.0][acc]
        e9, n9, g9, H9 = out[1e9][acc]
        e6, n6, g6, H6 = out[6.0][acc]
        assert n0 == n9 == n6
        assert e0 == e9 and np.array_equal(g0, g9) and np.array_equal(H0, H9)
        assert e6 < 0.8 * e0
        assert np.all(np.diag(H6) < np.diag(H0)) and np.allclose(H6, H6.T, rtol=1e-6)


def test_jacobian_matches_finite_differences_of_warp(oracle, small_pair):
    """J = grad(T) . d(warp)/d(xi) at xi = 0 (inverse_compositional.rs:313-341): check the geometric part
    against central differences of lm_optimizer.rs:213-219's warp composed with se3::exp."""
    scene, f0, _, _ = small_pair
    cfg = _cfg(oracle, scene, nb_levels=3)
    kf = oracle.Keyframe(cfg, f0[1], f0[0])
    k5 = kf.intrinsics(1)
    xy, idepth, _ = kf.points(1)
    L = oracle.lib()
    h = 1e-3
    rng = np.random.default_rng(0)
    for p in rng.choice(len(xy), 40, replace=False):
        x, y, z = float(xy[p, 0]), float(xy[p, 1]), float(idepth[p])
        Ju = np.zeros(6, np.float32)
        Jv = np.zeros(6, np.float32)
        L.ref_warp_jacobian_at(1.0, 0.0, x, y, z, k5, Ju)  # gu = 1, gv = 0 -> du/dxi
        L.ref_warp_jacobian_at(0.0, 1.0, x, y, z, k5, Jv)
        fd = np.zeros((2, 6))
        for a in range(6):
            xi = np.zeros(6)
            xi[a] = h
            uvp = np.zeros(2, np.float32)
            uvm = np.zeros(2, np.float32)
            L.ref_warp(C.byref(oracle.se3_exp(xi)), x, y, z, k5, uvp)
            L.ref_warp(C.byref(oracle.se3_exp(-xi)), x, y, z, k5, uvm)
            fd[:, a] = (uvp.astype(np.float64) - uvm) / (2 * h)
        scale = max(1.0, np.abs(fd).max())
        assert np.allclose(Ju, fd[0], atol=2e-2 * scale), (Ju, fd[0])
        assert np.allclose(Jv, fd[1], atol=2e-2 * scale), (Jv, fd[1])


def test_eval_f32_and_f64_agree_and_gradient_descends(oracle, small_pair):
    scene, f0, f1, _ = small_pair
    cfg = _cfg(oracle, scene, nb_levels=3)
    kf = oracle.Keyframe(cfg, f0[1], f0[0])
    pyr1 = oracle.mean_pyramid(f1[0], 3)
    ident = oracle.Pose.identity()
    for l in range(3):
        e0, n0, g0, H0 = kf.eval(l, pyr1[l], ident, 0)
        e1, n1, g1, H1 = kf.eval(l, pyr1[l], ident, 1)
        assert n0 == n1 and n0 > 0
        assert np.isclose(e0, e1, rtol=1e-4)
        assert np.allclose(g0, g1, rtol=1e-3, atol=1e-3 * np.abs(g1).max())
        assert np.allclose(H0, H1, rtol=1e-3)
        assert np.allclose(H1, H1.T)
    # identical image -> (numerically) zero residual: back_project/project round-off moves u,v by ~1e-5 px
    e, n, g, H = kf.eval(0, f0[0], ident, 0)
    assert e < 1e-5 and n > 0


def test_pair_alignment_recovers_ground_truth(oracle):
    scene, f0, f1, pose1 = synth.make_pair(seed=1000)
    cfg = _cfg(oracle, scene, nb_levels=5)
    tr = oracle.Tracker(cfg, 0.0, f0[1], 0.0, f0[0])
    st, stats, trace = tr.track(1.0, f1[1], 1.01, f1[0], trace_cap=512)
    assert st == 0 and stats.status == 0
    ts, p = tr.current_frame()
    assert ts == 1.0  # depth timestamp (inverse_compositional.rs:245)
    ang, dist = oracle.pose_error(p.as_array(), np.concatenate(pose1))
    assert ang < 2e-3 and dist < 5e-3, (ang, dist)
    # trace invariants of the LM loop: levels coarse to fine, energy never increases on accepted steps
    levels = [r.level for r in trace]
    assert levels == sorted(levels, reverse=True) and levels[0] == 4 and levels[-1] == 0
    for a, b in zip(trace, trace[1:]):
        if a.level == b.level and b.accepted:
            last_acc = [r for r in trace if r.level == a.level and r.iter < b.iter and r.accepted][-1]
            assert b.energy <= last_acc.energy


def test_fixed_iters_runs_exactly_k_rounds(oracle, small_pair):
    scene, f0, f1, _ = small_pair
    cfg = _cfg(oracle, scene, nb_levels=3, fixed_iters=10, candidate_mode=1)
    tr = oracle.Tracker(cfg, 0.0, f0[1], 0.0, f0[0])
    st, stats, trace = tr.track(1.0, f1[1], 1.0, f1[0], trace_cap=512)
    assert st == 0
    assert list(stats.n_iters)[:3] == [10, 10, 10]
    assert len(trace) == 3 * 11
    kf = tr.keyframe()
    # dense mode: every pixel with depth != 0 is a level-0 candidate
    assert kf.n_points(0) == int(np.count_nonzero(f0[1]))


def test_tracker_keyframe_switch_and_failure_semantics(oracle, small_pair):
    scene, f0, f1, pose1 = small_pair
    cfg = _cfg(oracle, scene, nb_levels=3, keyframe_flow_threshold=1e-3)
    tr = oracle.Tracker(cfg, 0.0, f0[1], 0.0, f0[0])
    st, stats, _ = tr.track(1.0, f1[1], 1.0, f1[0])
    assert st == 0 and stats.keyframe_changed == 1
    _, cur = tr.current_frame()
    assert np.array_equal(tr.keyframe_pose().as_array(), cur.as_array())  # inverse_compositional.rs:238
    # all-unknown depth at init -> no candidates -> NaN energy -> zero Hessian -> Cholesky fails;
    # pose is kept, timestamp still advances (inverse_compositional.rs:191-208)
    tr2 = oracle.Tracker(cfg, 0.0, np.zeros_like(f0[1]), 0.0, f0[0])
    st, stats, _ = tr2.track(2.0, f1[1], 2.0, f1[0])
    assert st == 1 and stats.status == 1
    ts, p = tr2.current_frame()
    assert ts == 2.0 and np.array_equal(p.as_array(), np.array([0, 0, 0, 0, 0, 0, 1], np.float32))
    assert stats.keyframe_changed == 0  # optical flow is NaN, `NaN >= thresh` is false


def test_short_pyramid_is_rejected(oracle):
    scene = synth.make_scene(0, 16, 16)
    cfg = _cfg(oracle, scene, nb_levels=6)
    with pytest.raises(ValueError):
        oracle.Tracker(cfg, 0.0, np.ones((16, 16), np.uint16), 0.0, np.zeros((16, 16), np.uint8))


def test_dso_select_bit_for_bit_against_numpy_restatement(oracle):
    """Row R: the oracle's `candidates::dso::select` against the independent numpy restatement written from
    candidates/dso.rs:98-325 (np_restate.dso_select): masks, candidate counts and the branch taken are identical over random
    heavy-tailed gradient images of awkward sizes (regions and blocks cut by the border, images smaller than a region), targets
    that exercise every branch - accepted as is, block-size recursion (both directions, 0 / 1 / 2 iterations left), seeded
    thinning - and the case where the reference panics (threshold beyond u16)."""
    import np_restate as R

    def oracle_dso(g, target, iters, seed):
        rows, cols = g.shape
        mask = np.zeros(g.size, np.uint8)
        used = C.c_int()
        n = oracle.lib().ref_dso_select(np.ascontiguousarray(g.T).reshape(-1), rows, cols, target, iters, seed, mask, C.byref(used))
        return mask.reshape(cols, rows).T.astype(bool), n, bool(used.value)

    rng = np.random.default_rng(5)
    cover = {"plain": 0, "random": 0, "exhausted": 0, "recursed": 0}
    for trial in range(21):
        rows, cols = [(60, 80), (37, 53), (96, 64), (33, 31), (120, 160), (8, 9), (65, 130)][trial % 7]
        g = np.minimum(rng.lognormal(np.log(4.0), [0.8, 1.2, 1.6][trial % 3], (rows, cols)), 60000).astype(np.uint16)
        if trial % 5 == 0:
            g[: rows // 2] //= 4  # regions with very different medians
        for target in (20, 150, 600, 3000):
            base = None
            for iters in (0, 1, 2):
                want_mask, want_n, want_random = R.dso_select(g, target, iters, 99 + trial)
                got_mask, got_n, got_random = oracle_dso(g, target, iters, 99 + trial)
                assert got_n == want_n and got_random == want_random and np.array_equal(got_mask, want_mask), (trial, target, iters)
                ratio = want_n / target
                cover["random" if want_random else "plain" if 0.8 <= ratio <= 4.0 else "exhausted"] += 1
                if iters == 0:
                    base = want_n
                elif want_n != base:
                    cover["recursed"] += 1  # another block size was tried
    assert all(v > 0 for v in cover.values()), cover
    # the reference's `num_traits::cast(..).expect("woops")` (dso.rs:300): (median + 3)^2 does not fit a u16
    g = np.full((40, 40), 300, np.uint16)
    assert R.dso_select(g, 100, 1, 1) is None and oracle_dso(g, 100, 1, 1)[1] == -1


def test_dso_select_deterministic_branches(oracle):
    rng = np.random.default_rng(11)
    scene = synth.make_scene(3, 240, 320)
    gray, _ = synth.render(scene)
    g2 = np.zeros(gray.size, np.uint16)
    oracle.lib().ref_squared_norm_direct(np.ascontiguousarray(gray.T).reshape(-1), 240, 320, g2)
    mag = np.sqrt(g2.astype(np.float32)).astype(np.uint16)
    mask = np.zeros(gray.size, np.uint8)
    used = C.c_int()
    n = oracle.lib().ref_dso_select(mag, 240, 320, 500, 2, 1, mask, C.byref(used))
    assert n > 0 and 0 < mask.sum() <= n
    mask2 = np.zeros(gray.size, np.uint8)
    oracle.lib().ref_dso_select(mag, 240, 320, 500, 2, 1, mask2, C.byref(used))
    assert np.array_equal(mask, mask2)
    m = mask.reshape(320, 240).T.astype(bool)
    assert m.sum() > 0 and np.all(mag.reshape(320, 240).T[m] > 0)
