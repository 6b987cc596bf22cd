// lie.cuh — f32 rigid-motion algebra shared by the device LM step and the host tracker state machine.
//
// Restates what the reference gets from nalgebra 0.17 (Isometry3 / UnitQuaternion products, 6x6
// Cholesky) and its own src/math/se3.rs + src/math/so3.rs, in the same operation order so the
// device decisions follow the reference's.
#pragma once

#include <cuda_runtime.h>
#include <math.h>

#define VORS_HD __host__ __device__ __forceinline__

namespace vors {

struct Vec3 {
    float x, y, z;
};
struct Quat {  // nalgebra coords order (i, j, k, w)
    float i, j, k, w;
};
struct Pose {  // Isometry3<f32>: x -> q*x + t
    Vec3 t;
    Quat q;
};

VORS_HD Pose pose_identity() { return Pose{{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 1.f}}; }

VORS_HD Vec3 cross3(Vec3 a, Vec3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

// nalgebra `&UnitQuaternion * &Vector3`: t = 2 (v x p); t*w + v x t + p.
VORS_HD Vec3 quat_rotate(Quat q, Vec3 p) {
    const Vec3 v{q.i, q.j, q.k};
    Vec3 t = cross3(v, p);
    t = {t.x * 2.0f, t.y * 2.0f, t.z * 2.0f};
    const Vec3 c = cross3(v, t);
    return {t.x * q.w + c.x + p.x, t.y * q.w + c.y + p.y, t.z * q.w + c.z + p.z};
}

// nalgebra `&Quaternion * &Quaternion` (Hamilton product).
VORS_HD Quat quat_mul(Quat a, Quat b) {
    return {a.w * b.i + a.i * b.w + a.j * b.k - a.k * b.j, a.w * b.j - a.i * b.k + a.j * b.w + a.k * b.i,
            a.w * b.k + a.i * b.j - a.j * b.i + a.k * b.w, a.w * b.w - a.i * b.i - a.j * b.j - a.k * b.k};
}

// nalgebra dot() on a static 4-vector: (x0 y0 + x2 y2) + (x1 y1 + x3 y3).
VORS_HD float quat_norm2(Quat q) {
    float a = q.i * q.i, b = q.j * q.j;
    a += q.k * q.k;
    b += q.w * q.w;
    return a + b;
}

// Isometry3 * Isometry3 (inverse_compositional.rs:177, :207; lm_optimizer.rs:135).
VORS_HD Pose pose_mul(const Pose& a, const Pose& b) {
    const Vec3 s = quat_rotate(a.q, b.t);
    return Pose{{a.t.x + s.x, a.t.y + s.y, a.t.z + s.z}, quat_mul(a.q, b.q)};
}

// Isometry3::inverse: conj(q), q^-1 * (-t).
VORS_HD Pose pose_inverse(const Pose& a) {
    const Quat qi{-a.q.i, -a.q.j, -a.q.k, a.q.w};
    return Pose{quat_rotate(qi, Vec3{-a.t.x, -a.t.y, -a.t.z}), qi};
}

// lm_optimizer.rs:198-209 `renormalize`: first-order re-normalisation of the rotation.
VORS_HD Pose pose_renormalize(Pose m) {
    const float f = 0.5f * (3.0f - quat_norm2(m.q));
    m.q = {f * m.q.i, f * m.q.j, f * m.q.k, f * m.q.w};
    return m;
}

// src/math/se3.rs:65-95 `exp` with src/math/so3.rs:27-50 `hat`, `hat_2`; Taylor branch for theta^2 < 1e-4.
VORS_HD Pose se3_exp(const float xi[6]) {
    const float v0 = xi[0], v1 = xi[1], v2 = xi[2];
    const float wx = xi[3], wy = xi[4], wz = xi[5];
    const float theta_2 = wx * wx + wy * wy + wz * wz;
    float real_factor, imag_factor, c1, c2;
    if (theta_2 < 1e-2f * 1e-2f) {
        real_factor = 1.0f - 0.125f * theta_2;
        imag_factor = 0.5f - (1.0f / 48.0f) * theta_2;
        c1 = 0.5f - (1.0f / 24.0f) * theta_2;
        c2 = (1.0f / 6.0f) - (1.0f / 120.0f) * theta_2;
    } else {
        const float theta = sqrtf(theta_2);
        const float half_theta = 0.5f * theta;
        real_factor = cosf(half_theta);
        imag_factor = sinf(half_theta) / theta;
        c1 = (1.0f - cosf(theta)) / theta_2;
        c2 = (theta - sinf(theta)) / (theta * theta_2);
    }
    // Omega = hat(w), Omega^2 = hat_2(w)
    const float w11 = wx * wx, w12 = wx * wy, w13 = wx * wz, w22 = wy * wy, w23 = wy * wz, w33 = wz * wz;
    const float O[3][3] = {{0.f, -wz, wy}, {wz, 0.f, -wx}, {-wy, wx, 0.f}};
    const float O2[3][3] = {{-w22 - w33, w12, w13}, {w12, -w11 - w33, w23}, {w13, w23, -w11 - w22}};
    float V[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) V[r][c] = ((r == c ? 1.0f : 0.0f) + c1 * O[r][c]) + c2 * O2[r][c];
    Pose out;
    out.t = {(V[0][0] * v0 + V[0][1] * v1) + V[0][2] * v2, (V[1][0] * v0 + V[1][1] * v1) + V[1][2] * v2,
             (V[2][0] * v0 + V[2][1] * v1) + V[2][2] * v2};
    Quat q{imag_factor * wx, imag_factor * wy, imag_factor * wz, real_factor};
    const float n = sqrtf(quat_norm2(q));  // UnitQuaternion::from_quaternion normalises
    out.q = {q.i / n, q.j / n, q.k / n, q.w / n};
    return out;
}

// src/math/so3.rs:61-77 `exp`.
VORS_HD Quat so3_exp(Vec3 w) {
    const float theta_2 = w.x * w.x + w.y * w.y + w.z * w.z;
    float real_factor, imag_factor;
    if (theta_2 < 1e-2f * 1e-2f) {
        real_factor = 1.0f - 0.125f * theta_2;
        imag_factor = 0.5f - (1.0f / 48.0f) * theta_2;
    } else {
        const float theta = sqrtf(theta_2);
        const float half_theta = 0.5f * theta;
        real_factor = cosf(half_theta);
        imag_factor = sinf(half_theta) / theta;
    }
    const Quat q{imag_factor * w.x, imag_factor * w.y, imag_factor * w.z, real_factor};
    const float n = sqrtf(quat_norm2(q));
    return {q.i / n, q.j / n, q.k / n, q.w / n};
}

// src/math/so3.rs:81-99 `log`.
VORS_HD Vec3 so3_log(Quat q) {
    const float imag_norm_2 = q.i * q.i + q.j * q.j + q.k * q.k;
    const float real_factor = q.w;
    float f;
    if (imag_norm_2 < 1e-2f * 1e-2f) {
        f = 2.0f / real_factor;
    } else if (fabsf(real_factor) < 1e-2f) {
        const float imag_norm = sqrtf(imag_norm_2);
        const float alpha = fabsf(real_factor) / imag_norm;
        const float theta = copysignf(1.0f, real_factor) * (3.14159265358979323846f - 2.0f * alpha);
        f = theta / imag_norm;
    } else {
        const float imag_norm = sqrtf(imag_norm_2);
        f = 2.0f * atanf(imag_norm / real_factor) / imag_norm;
    }
    return {f * q.i, f * q.j, f * q.k};
}

// src/math/se3.rs:99-130 `log` (trajectory error metrics; not on the tracking path).
VORS_HD void se3_log(const Pose& iso, float xi[6]) {
    const float imag_norm_2 = iso.q.i * iso.q.i + iso.q.j * iso.q.j + iso.q.k * iso.q.k;
    const float real_factor = iso.q.w;
    float wx, wy, wz, coef_omega_2;
    if (imag_norm_2 < 1e-2f * 1e-2f) {
        const float t = 2.0f / real_factor;
        wx = t * iso.q.i; wy = t * iso.q.j; wz = t * iso.q.k;
        const float x_2 = imag_norm_2 / (real_factor * real_factor);
        coef_omega_2 = (1.0f / 12.0f) * (1.0f + (1.0f / 15.0f) * x_2);
    } else {
        const float imag_norm = sqrtf(imag_norm_2);
        float theta;
        if (fabsf(real_factor) < 1e-2f) {
            const float alpha = fabsf(real_factor) / imag_norm;
            theta = copysignf(1.0f, real_factor) * (3.14159265358979323846f - 2.0f * alpha);
        } else {
            theta = 2.0f * atanf(imag_norm / real_factor);
        }
        const float theta_2 = theta * theta;
        const float t = theta / imag_norm;
        wx = t * iso.q.i; wy = t * iso.q.j; wz = t * iso.q.k;
        coef_omega_2 = (1.0f - 0.5f * theta * real_factor / imag_norm) / theta_2;
    }
    const float w11 = wx * wx, w12 = wx * wy, w13 = wx * wz, w22 = wy * wy, w23 = wy * wz, w33 = wz * wz;
    const float O[3][3] = {{0.f, -wz, wy}, {wz, 0.f, -wx}, {-wy, wx, 0.f}};
    const float O2[3][3] = {{-w22 - w33, w12, w13}, {w12, -w11 - w33, w23}, {w13, w23, -w11 - w22}};
    float V[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) V[r][c] = ((r == c ? 1.0f : 0.0f) - 0.5f * O[r][c]) + coef_omega_2 * O2[r][c];
    xi[0] = (V[0][0] * iso.t.x + V[0][1] * iso.t.y) + V[0][2] * iso.t.z;
    xi[1] = (V[1][0] * iso.t.x + V[1][1] * iso.t.y) + V[1][2] * iso.t.z;
    xi[2] = (V[2][0] * iso.t.x + V[2][1] * iso.t.y) + V[2][2] * iso.t.z;
    xi[3] = wx; xi[4] = wy; xi[5] = wz;
}

// nalgebra 0.17 `Matrix6::cholesky()` + `Cholesky::solve` (lm_optimizer.rs:131-134): left-looking
// LL^T on the lower triangle, fails on a pivot that is not > 0 (zero and NaN), then forward and
// transposed-back substitution.  A is row-major 6x6 (only the lower triangle is read), b is
// overwritten with the solution.  Returns false when the decomposition fails.
VORS_HD bool cholesky6_solve(float A[36], float b[6]) {
    // fully unrolled: on the device A and b then live in registers (the LM step is a latency-bound single-thread chain)
#pragma unroll
    for (int j = 0; j < 6; ++j) {
#pragma unroll
        for (int k = 0; k < j; ++k) {
            const float factor = -A[j * 6 + k];
#pragma unroll
            for (int i = j; i < 6; ++i) A[i * 6 + j] = factor * A[i * 6 + k] + A[i * 6 + j];
        }
        const float diag = A[j * 6 + j];
        if (!(diag > 0.0f)) return false;
        const float denom = sqrtf(diag);
        A[j * 6 + j] = denom;
#pragma unroll
        for (int i = j + 1; i < 6; ++i) A[i * 6 + j] /= denom;
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const float coeff = b[i] / A[i * 6 + i];
        b[i] = coeff;
#pragma unroll
        for (int r = i + 1; r < 6; ++r) b[r] = (-coeff) * A[r * 6 + i] + b[r];
    }
#pragma unroll
    for (int i = 5; i >= 0; --i) {
        float dot = 0.0f;
#pragma unroll
        for (int r = i + 1; r < 6; ++r) dot += A[r * 6 + i] * b[r];
        b[i] = (b[i] - dot) / A[i * 6 + i];
    }
    return true;
}

// Per-level pinhole intrinsics (src/core/camera.rs:84-123).
struct Intrinsics {
    float fx, fy, cx, cy, s;
};
// camera.rs:115-123 `half_res`.
VORS_HD Intrinsics half_res(const Intrinsics& k) {
    return {0.5f * k.fx, 0.5f * k.fy, (k.cx + 0.5f) / 2.0f - 0.5f, (k.cy + 0.5f) / 2.0f - 0.5f, k.s};
}

// The warp of lm_optimizer.rs:213-219 (back_project, rigid motion, project, perspective divide)
// folded into one 3x4 matrix acting on (x, y, 1, idepth):
//   [U V W]^T = K (R K^-1 [x y 1]^T + idepth * t),   u = U / W,  v = V / W.
// Built in f64 from the f32 model so the folding itself adds no error, stored as f32.
// `centred` = true builds the matrix that maps (x - cx, y - cy, 1, idepth) to (U - cx W, V - cy W, W) instead, so that
// u = cx + Uc / W: the align kernel feeds it the rounded differences the reference itself starts from
// (camera.rs:135-140 `back_project`) and adds the principal point last (camera.rs:126-132 `project`).  No entry of
// that matrix carries the ~cx-sized constant whose f32 rounding would shift every warped pixel by the same 1e-5 px.
VORS_HD void warp_matrix(const Pose& m, const Intrinsics& k, float M[12], bool centred = false) {
    const double qi = m.q.i, qj = m.q.j, qk = m.q.k, qw = m.q.w;
    // rotation matrix of the (possibly slightly non-unit) quaternion exactly as quat_rotate applies it:
    // p + 2 w (v x p) + 2 v x (v x p)
    const double R[3][3] = {{1.0 - 2.0 * (qj * qj + qk * qk), 2.0 * (qi * qj - qk * qw), 2.0 * (qi * qk + qj * qw)},
                            {2.0 * (qi * qj + qk * qw), 1.0 - 2.0 * (qi * qi + qk * qk), 2.0 * (qj * qk - qi * qw)},
                            {2.0 * (qi * qk - qj * qw), 2.0 * (qj * qk + qi * qw), 1.0 - 2.0 * (qi * qi + qj * qj)}};
    const double fx = k.fx, fy = k.fy, cx = k.cx, cy = k.cy, s = k.s;
    // K^-1 columns: ray(x, y) = ((x - cx - s (y - cy) / fy) / fx, (y - cy) / fy, 1)
    const double Ki[3][3] = {{1.0 / fx, -s / (fx * fy), centred ? 0.0 : (s * cy / fy - cx) / fx},
                             {0.0, 1.0 / fy, centred ? 0.0 : -cy / fy},
                             {0.0, 0.0, 1.0}};
    double RK[3][3];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) RK[r][c] = R[r][0] * Ki[0][c] + R[r][1] * Ki[1][c] + R[r][2] * Ki[2][c];
    const double t[3] = {m.t.x, m.t.y, m.t.z};
    for (int c = 0; c < 4; ++c) {
        const double a0 = c < 3 ? RK[0][c] : t[0], a1 = c < 3 ? RK[1][c] : t[1], a2 = c < 3 ? RK[2][c] : t[2];
        M[0 * 4 + c] = float(fx * a0 + s * a1 + (centred ? 0.0 : cx * a2));
        M[1 * 4 + c] = float(fy * a1 + (centred ? 0.0 : cy * a2));
        M[2 * 4 + c] = float(a2);
    }
}

}  // namespace vors
