"""Independent vectorised numpy restatements of the integer stages (second opinion on the C++ oracle).

Arrays are [row, col].  Signed division truncates toward zero like Rust (numpy // floors)."""
import numpy as np


def tdiv(a, d):
    a = np.asarray(a, np.int32)
    return (np.sign(a) * (np.abs(a) // d)).astype(np.int32)


def mean_pyramid(img, max_levels):
    """src/core/multires.rs:21-31, 38-88."""
    out = [np.asarray(img, np.uint8)]
    while len(out) < max_levels:
        m = out[-1]
        hr, hc = m.shape[0] // 2, m.shape[1] // 2
        if hr == 0 or hc == 0:
            break
        m = m[:2 * hr, :2 * hc].astype(np.uint16)
        s = m[0::2, 0::2] + m[1::2, 0::2] + m[0::2, 1::2] + m[1::2, 1::2]
        out.append((s // 4).astype(np.uint8))
    return out


def centered(img):
    """src/core/gradient.rs:15-33."""
    m = np.asarray(img, np.int32)
    gx = np.zeros(m.shape, np.int16)
    gy = np.zeros(m.shape, np.int16)
    if m.shape[0] > 2 and m.shape[1] > 2:
        gx[1:-1, 1:-1] = tdiv(m[1:-1, 2:] - m[1:-1, :-2], 2)
        gy[1:-1, 1:-1] = tdiv(m[2:, 1:-1] - m[:-2, 1:-1], 2)
    return gx, gy


def bloc_gradients(finer):
    """src/core/gradient.rs:74-93 through multires::halve: a=(2i,2j) b=(2i+1,2j) c=(2i,2j+1) d=(2i+1,2j+1)."""
    m = np.asarray(finer, np.int32)
    hr, hc = m.shape[0] // 2, m.shape[1] // 2
    m = m[:2 * hr, :2 * hc]
    a, b, c, d = m[0::2, 0::2], m[1::2, 0::2], m[0::2, 1::2], m[1::2, 1::2]
    return tdiv(c + d - a - b, 2).astype(np.int16), tdiv(b - a + d - c, 2).astype(np.int16)


def gradients_tracker(pyr):
    """inverse_compositional.rs:112-117."""
    gxs, gys = [], []
    gx, gy = centered(pyr[0])
    gxs.append(gx)
    gys.append(gy)
    for l in range(1, len(pyr)):
        gx, gy = bloc_gradients(pyr[l - 1])
        gxs.append(gx)
        gys.append(gy)
    g2 = [(x.astype(np.int32) ** 2 + y.astype(np.int32) ** 2).astype(np.uint16) for x, y in zip(gxs, gys)]
    return gxs, gys, g2


def c2f_select(thresh, g2_levels):
    """src/core/candidates/coarse_to_fine.rs:15-89; g2 finest first -> masks finest first."""
    L = len(g2_levels)
    masks = [None] * L
    masks[L - 1] = np.ones(g2_levels[L - 1].shape, bool)
    for l in range(L - 2, -1, -1):
        g = np.asarray(g2_levels[l], np.uint16)
        pre = masks[l + 1]
        hr, hc = pre.shape
        assert (hr, hc) == (g.shape[0] // 2, g.shape[1] // 2)
        gg = g[:2 * hr, :2 * hc]
        vals = np.stack([gg[0::2, 0::2], gg[1::2, 0::2], gg[0::2, 1::2], gg[1::2, 1::2]], 0).reshape(4, -1)
        order = np.argsort(vals, axis=0, kind="stable")
        srt = np.take_along_axis(vals, order, 0)
        first, second = order[3], order[2]
        keep2 = srt[2] > (srt[1] + np.uint16(thresh)).astype(np.uint16)
        sel = np.zeros((4, vals.shape[1]), bool)
        cols = np.arange(vals.shape[1])
        sel[first, cols] = True
        sel[second[keep2], cols[keep2]] = True
        sel &= pre.reshape(1, -1)
        sel = sel.reshape(4, hr, hc)
        m = np.zeros(g.shape, bool)
        m[0:2 * hr:2, 0:2 * hc:2] = sel[0]
        m[1:2 * hr:2, 0:2 * hc:2] = sel[1]
        m[0:2 * hr:2, 1:2 * hc:2] = sel[2]
        m[1:2 * hr:2, 1:2 * hc:2] = sel[3]
        masks[l] = m
    return masks


# ======================================================================================================================
# Second, independent restatement of the FLOAT path, written from the reference's Rust sources and
# from nalgebra 0.17's published arithmetic - not from the C++ oracle.  Everything is IEEE f32 with one rounding per
# operation (numpy float32 ufuncs do not contract a*b+c into an FMA, matching rustc's default) and sums run in the
# reference's sequential order (np.cumsum on float32 accumulates left to right in float32).
# Candidates are vectorised (the per-candidate arithmetic is independent); only the reductions are order-sensitive.
import math

F = np.float32


def _f(x):
    return np.asarray(x, dtype=F)


def quat_rotate(q, p):
    """nalgebra 0.17 `UnitQuaternion * Vector3`: t = 2 (v x p); t * w + v x t + p.  q = (i, j, k, w); p = (x, y, z) arrays."""
    qi, qj, qk, qw = [F(v) for v in q]
    px, py, pz = p
    two = F(2.0)
    tx, ty, tz = (qj * pz - qk * py) * two, (qk * px - qi * pz) * two, (qi * py - qj * px) * two
    cx, cy, cz = qj * tz - qk * ty, qk * tx - qi * tz, qi * ty - qj * tx
    return tx * qw + cx + px, ty * qw + cy + py, tz * qw + cz + pz


def quat_mul(a, b):
    """nalgebra 0.17 `Quaternion * Quaternion` (Hamilton product), coordinates (i, j, k, w)."""
    ai, aj, ak, aw = [F(v) for v in a]
    bi, bj, bk, bw = [F(v) for v in b]
    return np.array([aw * bi + ai * bw + aj * bk - ak * bj,
                     aw * bj - ai * bk + aj * bw + ak * bi,
                     aw * bk + ai * bj - aj * bi + ak * bw,
                     aw * bw - ai * bi - aj * bj - ak * bk], F)


def quat_norm2(q):
    """nalgebra 0.17 dot() on a static 4-vector: a = x0 y0, b = x1 y1, a += x2 y2, b += x3 y3, a + b."""
    q = _f(q)
    a, b = q[0] * q[0], q[1] * q[1]
    a = a + q[2] * q[2]
    b = b + q[3] * q[3]
    return a + b


def iso_mul(a, b):
    """`Isometry3 * Isometry3`: t = t_a + q_a * t_b, q = q_a q_b.  An isometry is (t[3], q[4]) of float32."""
    r = quat_rotate(a[1], (b[0][0], b[0][1], b[0][2]))
    return np.array([a[0][0] + r[0], a[0][1] + r[1], a[0][2] + r[2]], F), quat_mul(a[1], b[1])


def iso_inverse(a):
    """`Isometry3::inverse`: q^-1 = conj(q), t' = q^-1 * (-t)."""
    qi = np.array([-a[1][0], -a[1][1], -a[1][2], a[1][3]], F)
    r = quat_rotate(qi, (-a[0][0], -a[0][1], -a[0][2]))
    return np.array(r, F), qi


def back_project(intr, x, y, depth):
    """src/core/camera.rs:135-140.  intr = (fx, fy, cx, cy, s)."""
    fx, fy, cx, cy, s = [F(v) for v in intr]
    z = depth
    Y = (y - cy) * z / fy
    X = ((x - cx) * z - s * Y) / fx
    return X, Y, z


def project(intr, p):
    """src/core/camera.rs:126-132."""
    fx, fy, cx, cy, s = [F(v) for v in intr]
    return fx * p[0] + s * p[1] + cx * p[2], fy * p[1] + cy * p[2], p[2]


def warp(model, x, y, idepth, intr):
    """src/core/track/lm_optimizer.rs:213-219: back_project at depth 1 / _z, `model * point`, project, divide."""
    x1 = back_project(intr, _f(x), _f(y), F(1.0) / _f(idepth))
    r = quat_rotate(model[1], x1)
    x2 = (r[0] + model[0][0], r[1] + model[0][1], r[2] + model[0][2])  # Translation * (Rotation * point)
    uvz = project(intr, x2)
    return uvz[0] / uvz[2], uvz[1] / uvz[2]


def interpolate(x, y, image):
    """src/core/track/lm_optimizer.rs:227-251 on arrays: (inside mask, value); image is [row, col] u8."""
    height, width = image.shape
    with np.errstate(invalid="ignore"):
        u, v = np.floor(x), np.floor(y)
        inside = (u >= 0.0) & (u < F(width - 2)) & (v >= 0.0) & (v < F(height - 2))
    u0 = np.where(inside, u, 0).astype(np.int64)
    v0 = np.where(inside, v, 0).astype(np.int64)
    img = image.astype(F)
    vu00, vu10, vu01, vu11 = img[v0, u0], img[v0 + 1, u0], img[v0, u0 + 1], img[v0 + 1, u0 + 1]
    a, b = x - u, y - v
    one = F(1.0)
    val = (one - b) * (one - a) * vu00 + b * (one - a) * vu10 + (one - b) * a * vu01 + b * a * vu11
    return inside, val


def warp_jacobians(intr, xy, idepth, gx, gy):
    """src/core/track/inverse_compositional.rs:284-341 (`warp_jacobian_at`), one row per candidate."""
    fu, fv, cu, cv, s = [F(v) for v in intr]
    u, v, _z = _f(xy[:, 0]), _f(xy[:, 1]), _f(idepth)
    gu, gv = _f(gx), _f(gy)
    a, b = u - cu, v - cv
    c = a * fv - s * b
    _fv = F(1.0) / fv
    _fuv = F(1.0) / (fu * fv)
    return np.stack([gu * _z * fu,
                     _z * (gu * s + gv * fv),
                     -_z * (gu * a + gv * b),
                     gu * (-a * b * _fv - s) + gv * (-b * b * _fv - fv),
                     gu * (a * c * _fuv + fu) + gv * (b * c * _fuv),
                     gu * (-fu * fu * b + s * c) * _fuv + gv * (c / fu)], 1).astype(F)


def _seq_sum(a):
    """Sequential float32 sum down axis 0 starting from 0 (x += a[0]; x += a[1]; ...)."""
    a = np.asarray(a, F)
    if a.shape[0] == 0:
        return np.zeros(a.shape[1:], F)
    return np.cumsum(a, axis=0, dtype=F)[-1]


def eval_energy(obs, model):
    """lm_optimizer.rs:68-87 -> (energy, inside indices, residuals)."""
    u, v = warp(model, obs["xy"][:, 0], obs["xy"][:, 1], obs["idepth"], obs["intr"])
    inside, val = interpolate(u, v, obs["image"])
    idx = np.nonzero(inside)[0]
    tmpl = obs["template"][obs["xy"][idx, 1], obs["xy"][idx, 0]].astype(F)
    r = (val[idx] - tmpl).astype(F)
    with np.errstate(invalid="ignore", divide="ignore"):
        energy = _seq_sum(r * r) / F(len(idx))
    return F(energy), idx, r


def compute_eval_data(obs, model, pre):
    """lm_optimizer.rs:90-107: gradient += jac * r; hessian += jac jac^T over the inside set, in order."""
    energy, idx, r = pre
    jac = obs["jac"][idx]
    g = _seq_sum(jac * r[:, None])
    hes = (jac[:, :, None] * jac[:, None, :]).reshape(-1, 36)  # `j * j.transpose()`: one product per entry
    H = _seq_sum(hes).reshape(6, 6)
    return dict(hessian=H, gradient=g, energy=energy, model=model)


def cholesky_solve6(H, g):
    """nalgebra 0.17 `Matrix6::cholesky()` (left-looking, column axpy `a * x + 1 * y`, fails on a pivot that is not > 0)
    + `Cholesky::solve` (forward substitution with axpy, then transposed back substitution with a sequential dot)."""
    A = np.array(H, F)
    n = 6
    for j in range(n):
        for k in range(j):
            factor = -A[j, k]
            A[j:, j] = factor * A[j:, k] + A[j:, j]
        diag = A[j, j]
        if not diag > 0:
            return None
        denom = np.sqrt(diag)
        A[j, j] = denom
        A[j + 1:, j] = A[j + 1:, j] / denom
    b = np.array(g, F)
    for i in range(n):
        coeff = b[i] / A[i, i]
        b[i] = coeff
        b[i + 1:] = (-coeff) * A[i + 1:, i] + b[i + 1:]
    for i in range(n - 1, -1, -1):
        dot = F(0.0)
        for r_ in range(i + 1, n):
            dot = dot + A[r_, i] * b[r_]
        b[i] = (b[i] - dot) / A[i, i]
    return b


# f32::cos / f32::sin in Rust lower to the platform libm's cosf / sinf (glibc here, as for the C++ oracle); they are
# faithfully but not always correctly rounded, so a double-precision evaluation rounded to f32 differs in rare cases.
import ctypes as _C
import ctypes.util as _Cu

_libm = _C.CDLL(_Cu.find_library("m") or "libm.so.6")
_libm.cosf.restype = _libm.sinf.restype = _C.c_float
_libm.cosf.argtypes = _libm.sinf.argtypes = [_C.c_float]


def _cosf(x):
    return F(_libm.cosf(float(x)))


def _sinf(x):
    return F(_libm.sinf(float(x)))


def se3_exp(xi):
    """src/math/se3.rs:65-95 with so3::hat / hat_2 (so3.rs:27-50); Taylor branch below theta^2 = 1e-4."""
    xi = _f(xi)
    v, w = xi[:3], xi[3:]
    theta_2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2]
    if theta_2 < F(1e-2) * F(1e-2):
        real = F(1.0) - F(0.125) * theta_2
        imag = F(0.5) - (F(1.0) / F(48.0)) * theta_2
        c1 = F(0.5) - (F(1.0) / F(24.0)) * theta_2
        c2 = (F(1.0) / F(6.0)) - (F(1.0) / F(120.0)) * theta_2
    else:
        theta = np.sqrt(theta_2)
        half = F(0.5) * theta
        real = _cosf(half)
        imag = _sinf(half) / theta
        c1 = (F(1.0) - _cosf(theta)) / theta_2
        c2 = (theta - _sinf(theta)) / (theta * theta_2)
    z = F(0.0)
    O = np.array([[z, -w[2], w[1]], [w[2], z, -w[0]], [-w[1], w[0], z]], F)
    w11, w12, w13, w22, w23, w33 = w[0] * w[0], w[0] * w[1], w[0] * w[2], w[1] * w[1], w[1] * w[2], w[2] * w[2]
    O2 = np.array([[-w22 - w33, w12, w13], [w12, -w11 - w33, w23], [w13, w23, -w11 - w22]], F)
    V = (np.eye(3, dtype=F) + c1 * O) + c2 * O2
    # Mat3 * Vec3 (gemv): y = V[:,0] v0; y = V[:,1] v1 + y; y = V[:,2] v2 + y
    t = V[:, 0] * v[0]
    t = V[:, 1] * v[1] + t
    t = V[:, 2] * v[2] + t
    q = np.array([imag * w[0], imag * w[1], imag * w[2], real], F)
    nrm = np.sqrt(quat_norm2(q))  # UnitQuaternion::from_quaternion normalises
    return t.astype(F), (q / nrm).astype(F)


def renormalize(m):
    """lm_optimizer.rs:198-209: q <- 0.5 (3 - |q|^2) q."""
    f = F(0.5) * (F(3.0) - quat_norm2(m[1]))
    return m[0], (f * m[1]).astype(F)


def lm_step(state):
    """lm_optimizer.rs:123-136; None = Cholesky failure."""
    H = np.array(state["eval"]["hessian"], F)
    for i in range(6):
        H[i, i] = H[i, i] * (F(1.0) + state["lm_coef"])
    delta = cholesky_solve6(H, state["eval"]["gradient"])
    if delta is None:
        return None
    return renormalize(iso_mul(state["eval"]["model"], iso_inverse(se3_exp(delta))))


def iterative_solve(obs, model, fixed_iters=0):
    """src/math/optimizer.rs:57-70 driving lm_optimizer.rs:113-192.  Returns (status, final model, nb_iter, trace) where a
    trace record is (iter, energy, n_inside, lm_coef used, accepted) like the oracle's / the GPU's."""
    pre = eval_energy(obs, model)
    state = dict(lm_coef=F(0.1), eval=compute_eval_data(obs, model, pre))
    trace = [(0, state["eval"]["energy"], len(pre[1]), F(0.1), 1)]
    nb_iter = 0
    while True:
        nb_iter += 1
        new_model = lm_step(state)
        if new_model is None:
            return 1, state["eval"]["model"], nb_iter, trace
        pre = eval_energy(obs, new_model)
        energy, old = pre[0], state["eval"]["energy"]
        rejected = bool(energy > old)  # NaN compares false -> accepted, like the reference
        trace.append((nb_iter, energy, len(pre[1]), state["lm_coef"], 0 if rejected else 1))
        too_many = nb_iter >= fixed_iters if fixed_iters else nb_iter > 20
        if rejected:
            if too_many:
                break
            state["lm_coef"] = state["lm_coef"] * F(10.0)
            continue
        ev = compute_eval_data(obs, new_model, pre)
        if too_many:
            state = dict(lm_coef=state["lm_coef"], eval=ev)
            break
        d_energy = old - energy
        go_on = True if fixed_iters else bool(d_energy > F(1.0))
        state = dict(lm_coef=F(0.1) * state["lm_coef"], eval=ev)
        if not go_on:
            break
    return 0, state["eval"]["model"], nb_iter, trace


def tracker_track(state, levels_obs, coarsest, fixed_iters=0):
    """src/core/track/inverse_compositional.rs:170-240 around `iterative_solve`.

    state: dict(keyframe_pose, current_frame_pose); levels_obs: per-level obs dicts (finest first) for the new frame;
    coarsest: dict(xy, idepth, intr) of the last level (the optical-flow test uses `.last()`).
    Returns (went_well, optical_flow); mutates state["current_frame_pose"]."""
    lm_model = iso_mul(iso_inverse(state["current_frame_pose"]), state["keyframe_pose"])
    went_well = True
    for lvl in range(len(levels_obs) - 1, -1, -1):
        st, model, _, _ = iterative_solve(levels_obs[lvl], lm_model, fixed_iters)
        if st != 0:
            went_well = False
            break
        lm_model = model
    if went_well:
        state["current_frame_pose"] = iso_mul(state["keyframe_pose"], iso_inverse(lm_model))
    x, y = _f(coarsest["xy"][:, 0]), _f(coarsest["xy"][:, 1])
    u, v = warp(lm_model, x, y, coarsest["idepth"], coarsest["intr"])
    flow_sum = _seq_sum(np.abs(x - u) + np.abs(y - v))
    with np.errstate(invalid="ignore", divide="ignore"):
        optical_flow = flow_sum / F(len(x))
    return went_well, F(optical_flow)


# ---- inverse-depth pyramid (rows F, G) and extract_z (row H), f32, the reference's operation order --------------------------
def idepth_level0(depth_u16, mask, scale, variance):
    """helper::zip_mask_map + inverse_depth::from_depth (helper.rs:40-47, inverse_depth.rs:24-29): scale / depth where the
    mask is set and depth != 0; Unknown -> (NaN, 0)."""
    known = mask.astype(bool) & (depth_u16 != 0)
    with np.errstate(divide="ignore"):
        d = np.where(known, np.float32(scale) / depth_u16.astype(np.float32), np.float32(np.nan)).astype(np.float32)
    return d, np.where(known, np.float32(variance), np.float32(0)).astype(np.float32)


def idepth_halve_dso_mean(d, v):
    """multires::halve (multires.rs:67-88) with inverse_depth::fuse + strategy_dso_mean (inverse_depth.rs:49-98): the known
    children of a 2x2 bloc in (a, b, c, d) = ((2i,2j), (2i+1,2j), (2i,2j+1), (2i+1,2j+1)) order; one child is copied, more are
    averaged with weights, sums and products rounded one by one, left to right, in f32."""
    f = np.float32
    hr, hc = d.shape[0] // 2, d.shape[1] // 2
    kids = [(d[0:2 * hr:2, 0:2 * hc:2], v[0:2 * hr:2, 0:2 * hc:2]), (d[1:2 * hr:2, 0:2 * hc:2], v[1:2 * hr:2, 0:2 * hc:2]),
            (d[0:2 * hr:2, 1:2 * hc:2], v[0:2 * hr:2, 1:2 * hc:2]), (d[1:2 * hr:2, 1:2 * hc:2], v[1:2 * hr:2, 1:2 * hc:2])]
    count = np.zeros((hr, hc), np.int32)
    first_d = np.full((hr, hc), np.nan, f)
    first_v = np.zeros((hr, hc), f)
    num = np.zeros((hr, hc), f)
    den = np.zeros((hr, hc), f)
    for kd, kv in kids:
        known = ~np.isnan(kd)
        prod = (np.where(known, kd, f(0)) * kv).astype(f)           # d_k * v_k, rounded
        is_first = known & (count == 0)
        first_d = np.where(is_first, kd, first_d)
        first_v = np.where(is_first, kv, first_v)
        num = np.where(is_first, prod, np.where(known, (num + prod).astype(f), num))   # ((d1 v1 + d2 v2) + d3 v3) + d4 v4
        den = np.where(is_first, kv, np.where(known, (den + kv).astype(f), den))        # ((v1 + v2) + v3) + v4
        count += known
    with np.errstate(invalid="ignore", divide="ignore"):
        mean = (num / den).astype(f)
    out_d = np.where(count == 0, f(np.nan), np.where(count == 1, first_d, mean)).astype(f)
    out_v = np.where(count == 0, f(0), np.where(count == 1, first_v, den)).astype(f)
    return out_d, out_v


def idepth_pyramid(depth_u16, mask, scale, variance, nb_levels):
    """multires::limited_sequence(nb_levels, level 0, halve) (inverse_compositional.rs:127-138)."""
    levels = [idepth_level0(depth_u16, mask, scale, variance)]
    while len(levels) < nb_levels and min(levels[-1][0].shape) >= 2:
        levels.append(idepth_halve_dso_mean(*levels[-1]))
    return levels


def extract_z(d):
    """extract_z (inverse_compositional.rs:260-279): known inverse depths in the matrix' (column-major) iteration order with
    their (u = column, v = row) coordinates."""
    cols, rows = np.nonzero(~np.isnan(d.T))
    return np.stack([cols, rows], 1), d[rows, cols]


# ---- candidates::dso::select (row R), restated from candidates/dso.rs:98-325 with the DEFAULT_* configurations (:72-90) ------
def _splitmix64(state):
    """The seeded generator that stands in for `rand::thread_rng()` (dso.rs:142) in the oracle and in the CUDA path."""
    M = (1 << 64) - 1
    state = (state + 0x9E3779B97F4A7C15) & M
    z = state
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
    return state, z ^ (z >> 31)


def _ceil_div(n, d):
    q, r = divmod(n, d)
    return q if r == 0 else q + 1


def dso_region_medians(g, size=32):
    """region_median_gradients (dso.rs:307-325): sorted region values, element len / 2."""
    rows, cols = g.shape
    out = np.zeros((_ceil_div(rows, size), _ceil_div(cols, size)), np.uint16)
    for i in range(out.shape[0]):
        for j in range(out.shape[1]):
            v = np.sort(g[i * size:min(rows, (i + 1) * size), j * size:min(cols, (j + 1) * size)].reshape(-1))
            out[i, j] = v[len(v) // 2]
    return out


def dso_region_thresholds(med, a=1.0, b=3):
    """region_thresholds (dso.rs:275-303): 3x3 (clipped) sum IN u16 ARITHMETIC (release builds wrap), mean + b in f32, squared,
    cast back to u16 (truncating; None -> the reference panics: returned as None here)."""
    f = np.float32
    rows, cols = med.shape
    out = np.zeros((rows, cols), np.uint16)
    for i in range(rows):
        for j in range(cols):
            blk = med[max(0, i - 1):min(rows, i + 2), max(0, j - 1):min(cols, j + 2)]
            s = int(blk.astype(np.int64).sum()) & 0xFFFF
            t = f(f(s) / f(blk.size)) + f(b)
            sq = f(f(f(a) * t) * t)
            if not (sq < f(65536.0)):
                return None
            out[i, j] = int(sq)
    return out


def dso_pick_all(g, thresholds, base_size, nb_levels=3, threshold_factor=0.5, region_size=32):
    """pick_all_block_candidates (dso.rs:154-187) with init_max_gradients (:190-216), max_of_four_gradients (:219-233) and
    pick_level_block_candidates (:240-270).  Returns (number picked, picked levels u8 [rows, cols])."""
    f = np.float32
    rows, cols = g.shape
    br, bc = _ceil_div(rows, base_size), _ceil_div(cols, base_size)
    # init_max_gradients: first strict maximum scanning the block column by column
    mg = np.zeros((br, bc), np.int64)   # value
    mi = np.zeros((br, bc), np.int64)
    mj = np.zeros((br, bc), np.int64)
    for bi in range(br):
        for bj in range(bc):
            blk = g[bi * base_size:min(rows, (bi + 1) * base_size), bj * base_size:min(cols, (bj + 1) * base_size)]
            k = int(np.argmax(blk.T.reshape(-1)))  # column-major scan; argmax returns the FIRST maximum = strict `>` update
            mg[bi, bj] = blk.T.reshape(-1)[k]
            mj[bi, bj] = bj * base_size + k // blk.shape[0]
            mi[bi, bj] = bi * base_size + k % blk.shape[0]
    levels = [(mg, mi, mj)]
    while len(levels) < nb_levels and min(levels[-1][0].shape) >= 2:  # multires::limited_sequence + halve
        pg, pi, pj = levels[-1]
        hr, hc = pg.shape[0] // 2, pg.shape[1] // 2
        ng, ni, nj = np.zeros((hr, hc), np.int64), np.zeros((hr, hc), np.int64), np.zeros((hr, hc), np.int64)
        for i in range(hr):
            for j in range(hc):
                best = None
                # g_max(g1, g_max(g2, g_max(g3, g4))), `if m1 < m2 { m2 } else { m1 }`: fold from the right, ties keep the left
                for (a, b_) in ((2 * i + 1, 2 * j + 1), (2 * i, 2 * j + 1), (2 * i + 1, 2 * j), (2 * i, 2 * j)):
                    cand = (pg[a, b_], pi[a, b_], pj[a, b_])
                    best = cand if best is None or not (cand[0] < best[0]) else best
                ng[i, j], ni[i, j], nj[i, j] = best
        levels.append((ng, ni, nj))
    picked = np.zeros((rows, cols), np.uint8)
    mask = np.ones((br, bc), bool)
    coef = f(1.0)
    total = 0
    for level, (lg, li, lj) in enumerate(levels):
        mh, mw = mask.shape
        nxt = np.ones((mh // 2, mw // 2), bool)
        for j in range(mw // 2 * 2):
            for i in range(mh // 2 * 2):
                if mask[i, j]:
                    th = thresholds[li[i, j] // region_size, lj[i, j] // region_size]
                    if f(lg[i, j]) >= f(coef * f(th)):
                        nxt[i // 2, j // 2] = False
                        picked[li[i, j], lj[i, j]] = level + 1
                        total += 1
                else:
                    nxt[i // 2, j // 2] = False
        mask = nxt
        coef = f(coef * f(threshold_factor))
    return total, picked


def dso_select(g, nb_target, nb_iterations_left, seed, base_size=4):
    """select (dso.rs:98-147).  Returns (mask, number of block candidates of the last iteration, used the random branch) or
    None where the reference would panic (threshold out of u16)."""
    f = np.float32
    th = dso_region_thresholds(dso_region_medians(g))
    if th is None:
        return None
    while True:
        nb, picked = dso_pick_all(g, th, base_size)
        ratio = f(f(nb) / f(nb_target))
        target_f = f(f(np.sqrt(ratio)) * f(f(base_size) + f(1.0)) - f(1.0))
        rounded = int(np.floor(abs(float(target_f)) + 0.5)) * (1 if target_f >= 0 else -1)  # f32::round: half away from zero
        target_size = max(1, rounded)
        if ratio < f(0.8) or ratio > f(4.0):
            if target_size != base_size and nb_iterations_left > 0:
                base_size, nb_iterations_left = target_size, nb_iterations_left - 1
                continue
            return picked > 0, nb, False
        if ratio > f(1.1):
            lim = int(f(255.0) / ratio) & 0xFF
            mask = np.zeros(picked.shape, bool)
            state = seed
            cols_, rows_ = np.nonzero(picked.T > 0)  # `picked.map` walks column-major; one draw per picked pixel
            for c, r in zip(cols_, rows_):
                state, z = _splitmix64(state)
                mask[r, c] = (z & 0xFF) <= lim
            return mask, nb, True
        return picked > 0, nb, False
