"""The oracle against every known answer the reference holds for this path:
the prune_with_thresh doc vectors and the so3/se3 tests restated from the reference's own test modules."""
import ctypes as C

import numpy as np
import pytest

RNG = np.random.default_rng(20240519)


def test_prune_with_thresh_doc_vectors(oracle):
    # src/core/candidates/coarse_to_fine.rs:68-71, thresh = 5
    assert oracle.prune_with_thresh(5, 0, 1, 8, 9) == [False, False, True, True]
    assert oracle.prune_with_thresh(5, 0, 9, 1, 8) == [False, True, False, True]
    assert oracle.prune_with_thresh(5, 1, 0, 9, 0) == [False, False, True, False]


def test_prune_ties_pick_later_index(oracle):
    # insertion-sort stability: among equal maxima the later index gets the "first" slot
    assert oracle.prune_with_thresh(7, 5, 5, 5, 5) == [False, False, False, True]
    # (9,9,1,1), thresh 0: sorted = 1(c),1(d),9(a),9(b): first = b, second = a, 9 > 1 + 0 keeps a too
    assert oracle.prune_with_thresh(0, 9, 9, 1, 1) == [True, True, False, False]
    # (9,9,9,1), thresh 7: second (b) is not > third (a) + 7 -> only c
    assert oracle.prune_with_thresh(7, 9, 9, 9, 1) == [False, False, True, False]
    # u16 wrap of `third + thresh` (release-mode Rust): 65000 + 1000 wraps to 464 < 65010
    assert oracle.prune_with_thresh(1000, 0, 65000, 65010, 65020) == [False, False, True, True]


def test_so3_exp_log_zero(oracle):
    # src/math/so3.rs:115-118
    q = np.zeros(4, np.float32)
    w = np.zeros(3, np.float32)
    oracle.lib().ref_so3_exp(np.zeros(3, np.float32), q)
    oracle.lib().ref_so3_log(q, w)
    assert np.array_equal(w, np.zeros(3, np.float32))


def test_se3_exp_log_zero(oracle):
    # src/math/se3.rs:145-148
    xi = oracle.se3_log(oracle.se3_exp(np.zeros(6)))
    assert np.array_equal(xi, np.zeros(6, np.float32))


@pytest.mark.parametrize("trial", range(50))
def test_hat_vee_roundtrips(oracle, trial):
    # so3.rs:123-126, se3.rs:153-156 (quickcheck over arbitrary floats)
    w = (RNG.standard_normal(3) * 10 ** RNG.uniform(-3, 3)).astype(np.float32)
    m = np.zeros(9, np.float32)
    back = np.zeros(3, np.float32)
    oracle.lib().ref_so3_hat(w, m)
    oracle.lib().ref_so3_vee(m, back)
    assert np.array_equal(back, w)
    xi = (RNG.standard_normal(6) * 10 ** RNG.uniform(-3, 3)).astype(np.float32)
    m4 = np.zeros(16, np.float32)
    back6 = np.zeros(6, np.float32)
    oracle.lib().ref_se3_hat(xi, m4)
    oracle.lib().ref_se3_vee(m4, back6)
    assert np.array_equal(back6, xi)


@pytest.mark.parametrize("trial", range(50))
def test_hat_2_ok(oracle, trial):
    # so3.rs:129-132: hat_2(w) == hat(w) * hat(w), exact in f32 (nalgebra 3x3 product, sequential dot)
    w = RNG.integers(-64, 64, 3).astype(np.float32) / 8  # exactly representable products
    h = np.zeros(9, np.float32)
    h2 = np.zeros(9, np.float32)
    oracle.lib().ref_so3_hat(w, h)
    oracle.lib().ref_so3_hat2(w, h2)
    H = h.reshape(3, 3)
    assert np.array_equal((H @ H).astype(np.float32), h2.reshape(3, 3))


def _relative_eq(a, b, eps):
    # approx::relative_eq!(epsilon = eps, max_relative = f32::EPSILON default)
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    d = np.abs(a - b)
    return bool(np.all((d <= eps) | (d <= np.maximum(np.abs(a), np.abs(b)) * np.finfo(np.float32).eps)))


@pytest.mark.parametrize("trial", range(100))
def test_so3_log_exp_round_trip(oracle, trial):
    # so3.rs:135-142, epsilon 1e-6 on rotations built from Euler angles (so3.rs:146)
    r, p, y = RNG.uniform(-3.0, 3.0, 3).astype(np.float32)
    q = np.zeros(4, np.float32)
    oracle.lib().ref_quat_from_euler(r, p, y, q)
    w = np.zeros(3, np.float32)
    q2 = np.zeros(4, np.float32)
    oracle.lib().ref_so3_log(q, w)
    oracle.lib().ref_so3_exp(w, q2)
    assert _relative_eq(q, q2, 1e-6) or _relative_eq(q, -q2, 1e-6)


@pytest.mark.parametrize("trial", range(100))
def test_se3_log_exp_round_trip(oracle, trial):
    # se3.rs:159-173, epsilon 1e-4 (se3.rs:142)
    r, p, y = RNG.uniform(-3.0, 3.0, 3).astype(np.float32)
    t = RNG.uniform(-5.0, 5.0, 3).astype(np.float32)
    q = np.zeros(4, np.float32)
    oracle.lib().ref_quat_from_euler(r, p, y, q)
    pose = oracle.Pose.from_arrays(t, q)
    back = oracle.se3_exp(oracle.se3_log(pose)).as_array()
    a = pose.as_array()
    ok = _relative_eq(a, back, 1e-4) or _relative_eq(np.concatenate([a[:3], -a[3:]]), back, 1e-4)
    assert ok, (a, back)


def test_se3_exp_matches_float64_formula(oracle):
    from vors_b200 import synth

    for _ in range(50):
        xi = RNG.uniform(-0.5, 0.5, 6)
        if _ % 5 == 0:
            xi[3:] *= 1e-3  # Taylor branch (theta^2 < 1e-4)
        p = oracle.se3_exp(xi).as_array()
        t, q = synth.se3_exp(xi.astype(np.float32).astype(np.float64))
        assert np.allclose(p[:3], t, atol=2e-6)
        assert np.allclose(p[3:], q, atol=2e-6)


def test_pose_algebra(oracle):
    from vors_b200 import synth

    for _ in range(20):
        a = oracle.se3_exp(RNG.uniform(-1, 1, 6))
        b = oracle.se3_exp(RNG.uniform(-1, 1, 6))
        ab = oracle.pose_mul(a, b).as_array()
        t, q = synth.pose_mul((np.array(a.t, np.float64), np.array(a.q, np.float64)),
                              (np.array(b.t, np.float64), np.array(b.q, np.float64)))
        assert np.allclose(ab[:3], t, atol=1e-5) and (np.allclose(ab[3:], q, atol=1e-5) or np.allclose(ab[3:], -q, atol=1e-5))
        ident = oracle.pose_mul(a, oracle.pose_inverse(a)).as_array()
        assert np.allclose(ident, [0, 0, 0, 0, 0, 0, 1], atol=1e-5)


def test_cholesky_solve_matches_numpy(oracle):
    for _ in range(20):
        A = RNG.standard_normal((6, 12))
        H = (A @ A.T).astype(np.float32)
        g = RNG.standard_normal(6).astype(np.float32)
        x = np.zeros(6, np.float32)
        assert oracle.lib().ref_cholesky_solve6(np.ascontiguousarray(H.reshape(-1)), g, x) == 1
        assert np.allclose(x, np.linalg.solve(H.astype(np.float64), g), rtol=2e-3, atol=1e-5)
    # non positive-definite / NaN pivots fail like nalgebra's `diag > 0` test
    x = np.zeros(6, np.float32)
    assert oracle.lib().ref_cholesky_solve6(np.zeros(36, np.float32), np.ones(6, np.float32), x) == 0
    bad = np.eye(6, dtype=np.float32)
    bad[3, 3] = np.nan
    assert oracle.lib().ref_cholesky_solve6(np.ascontiguousarray(bad.reshape(-1)), np.ones(6, np.float32), x) == 0
