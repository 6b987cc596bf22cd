// vors_oracle.cpp — CPU ORACLE (test infrastructure, NOT product code; see vors_oracle.h).
//
// Behavioural restatement of the reference's hot path, operation order preserved:
// sequential f32 sums in column-major candidate order, truncating integer division,
// stable tie-break of the 4-element sort, nalgebra 0.17 arithmetic as published.
// PARITY UNPINNED for everything but `prune_with_thresh` (doc vectors) and the
// so3/se3 round-trip properties: the reference holds no tests for the rest and
// cannot be built in this image (no Rust toolchain, nalgebra not vendored).
//
// Build: see oracle/Makefile (parity build -O2 -ffp-contract=off; timing build -O3 -march=native).
// All citations are path:line under /root/reference.

#include "vors_oracle.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <utility>
#include <vector>

namespace {

using Float = float;  // src/misc/type_aliases.rs:10

// Column-major dense matrix, the layout of nalgebra::DMatrix.
template <typename T>
struct Mat {
    int rows = 0, cols = 0;
    std::vector<T> d;
    Mat() = default;
    Mat(int r, int c, T v = T()) : rows(r), cols(c), d(size_t(r) * size_t(c), v) {}
    T& operator()(int r, int c) { return d[size_t(c) * rows + r]; }
    const T& operator()(int r, int c) const { return d[size_t(c) * rows + r]; }
    size_t size() const { return d.size(); }
};

template <typename T>
Mat<T> mat_from(const T* p, int rows, int cols) {
    Mat<T> m(rows, cols);
    std::memcpy(m.d.data(), p, sizeof(T) * m.size());
    return m;
}

// ---------------------------------------------------------------------------------------
// multires.rs:67-88 `halve`: f(a,b,c,d) over 2x2 blocks, a=(2i,2j) b=(2i+1,2j) c=(2i,2j+1)
// d=(2i+1,2j+1); odd last row/col dropped; "None" (ok=false) when a half size is 0.
template <typename T, typename U, typename F>
bool halve(const Mat<T>& m, F f, Mat<U>& out) {
    const int hr = m.rows / 2, hc = m.cols / 2;
    if (hr == 0 || hc == 0) return false;
    out = Mat<U>(hr, hc);
    for (int j = 0; j < hc; ++j)
        for (int i = 0; i < hr; ++i)
            out(i, j) = f(m(2 * i, 2 * j), m(2 * i + 1, 2 * j), m(2 * i, 2 * j + 1), m(2 * i + 1, 2 * j + 1));
    return true;
}

// multires.rs:38-60 `limited_sequence` / `sequence`: at most max_length elements
// (max_length 0 behaves like 1), stops early when f yields None.
template <typename T, typename F>
std::vector<T> limited_sequence(int max_length, T data, F f) {
    std::vector<T> s;
    s.push_back(std::move(data));
    int length = 1;
    for (;;) {
        if (!(length < max_length)) break;
        ++length;
        T next;
        if (!f(s.back(), next)) break;
        s.push_back(std::move(next));
    }
    return s;
}

// multires.rs:21-31 `mean_pyramid`: ((a+b+c+d)/4) as u8 in u16 arithmetic.
std::vector<Mat<uint8_t>> mean_pyramid(int max_levels, Mat<uint8_t> img) {
    return limited_sequence(max_levels, std::move(img), [](const Mat<uint8_t>& m, Mat<uint8_t>& o) {
        return halve<uint8_t, uint8_t>(
            m,
            [](uint8_t a, uint8_t b, uint8_t c, uint8_t d) {
                const uint16_t s = uint16_t(uint16_t(a) + uint16_t(b) + uint16_t(c) + uint16_t(d));
                return uint8_t(s / 4);
            },
            o);
    });
}

// gradient.rs:15-33 `centered`: (right-left)/2 and (bottom-top)/2 in i16, Rust `/` truncates
// toward zero (so does C++); 1-pixel border stays 0.
void gradient_centered(const Mat<uint8_t>& img, Mat<int16_t>& gx, Mat<int16_t>& gy) {
    const int R = img.rows, C = img.cols;
    gx = Mat<int16_t>(R, C, 0);
    gy = Mat<int16_t>(R, C, 0);
    for (int j = 0; j + 2 < C; ++j)
        for (int i = 0; i + 2 < R; ++i) {
            const int16_t right = img(i + 1, j + 2), left = img(i + 1, j);
            const int16_t bottom = img(i + 2, j + 1), top = img(i, j + 1);
            gx(i + 1, j + 1) = int16_t((right - left) / 2);
            gy(i + 1, j + 1) = int16_t((bottom - top) / 2);
        }
}

// gradient.rs:74-80 `bloc_x`, :87-93 `bloc_y` (block a c / b d).
inline int16_t bloc_x(uint8_t a, uint8_t b, uint8_t c, uint8_t d) {
    return int16_t((int16_t(c) + int16_t(d) - int16_t(a) - int16_t(b)) / 2);
}
inline int16_t bloc_y(uint8_t a, uint8_t b, uint8_t c, uint8_t d) {
    return int16_t((int16_t(b) - int16_t(a) + int16_t(d) - int16_t(c)) / 2);
}
// gradient.rs:102-111 `bloc_squared_norm`.
inline uint16_t bloc_squared_norm(uint8_t a, uint8_t b, uint8_t c, uint8_t d) {
    const int32_t dx = int32_t(c) + int32_t(d) - int32_t(a) - int32_t(b);
    const int32_t dy = int32_t(b) - int32_t(a) + int32_t(d) - int32_t(c);
    return uint16_t((dx * dx + dy * dy) / 4);
}

// gradient.rs:38-44 `squared_norm`: (gx*gx + gy*gy) as u16 from the truncated i16 gradients.
Mat<uint16_t> squared_norm(const Mat<int16_t>& gx, const Mat<int16_t>& gy) {
    Mat<uint16_t> o(gx.rows, gx.cols);
    for (size_t k = 0; k < o.size(); ++k) {
        const int32_t x = gx.d[k], y = gy.d[k];
        o.d[k] = uint16_t(x * x + y * y);
    }
    return o;
}

// gradient.rs:49-65 `squared_norm_direct`: ((dx^2+dy^2)/4) as u16 without intermediate truncation.
Mat<uint16_t> squared_norm_direct(const Mat<uint8_t>& im) {
    const int R = im.rows, C = im.cols;
    Mat<uint16_t> o(R, C, 0);
    for (int j = 0; j + 2 < C; ++j)
        for (int i = 0; i + 2 < R; ++i) {
            const int32_t gx = int32_t(im(i + 1, j + 2)) - int32_t(im(i + 1, j));
            const int32_t gy = int32_t(im(i + 2, j + 1)) - int32_t(im(i, j + 1));
            o(i + 1, j + 1) = uint16_t((gx * gx + gy * gy) / 4);
        }
    return o;
}

// multires.rs:112-126 `gradients_xy` + inverse_compositional.rs:112-117: G[0]=centered(I[0]),
// G[l]=halve(I[l-1], bloc_x / bloc_y) for l>=1, g2[l]=squared_norm(G[l]).
// Extension (not in the reference; the north_star's "Scharr"): 3x3 Scharr operator on a level's own image, normalised by 32 so
// that it has the scale of the centred difference, i16 division truncating toward zero, 1-px border 0.
void gradient_scharr(const Mat<uint8_t>& im, Mat<int16_t>& gx, Mat<int16_t>& gy) {
    gx = Mat<int16_t>(im.rows, im.cols);
    gy = Mat<int16_t>(im.rows, im.cols);
    for (int c = 1; c + 1 < im.cols; ++c)
        for (int r = 1; r + 1 < im.rows; ++r) {
            auto I = [&](int dr, int dc) { return int(im(r + dr, c + dc)); };
            const int sx = 3 * (I(-1, 1) - I(-1, -1)) + 10 * (I(0, 1) - I(0, -1)) + 3 * (I(1, 1) - I(1, -1));
            const int sy = 3 * (I(1, -1) - I(-1, -1)) + 10 * (I(1, 0) - I(-1, 0)) + 3 * (I(1, 1) - I(-1, 1));
            gx(r, c) = int16_t(sx / 32);
            gy(r, c) = int16_t(sy / 32);
        }
}

void gradients_tracker(const std::vector<Mat<uint8_t>>& pyr, std::vector<Mat<int16_t>>& gxs,
                       std::vector<Mat<int16_t>>& gys, std::vector<Mat<uint16_t>>& g2s, bool scharr = false) {
    const size_t L = pyr.size();
    gxs.assign(L, {});
    gys.assign(L, {});
    if (scharr) {
        for (size_t l = 0; l < L; ++l) gradient_scharr(pyr[l], gxs[l], gys[l]);
    } else {
        gradient_centered(pyr[0], gxs[0], gys[0]);
        for (size_t l = 1; l < L; ++l) {
            halve<uint8_t, int16_t>(pyr[l - 1], bloc_x, gxs[l]);
            halve<uint8_t, int16_t>(pyr[l - 1], bloc_y, gys[l]);
        }
    }
    g2s.clear();
    for (size_t l = 0; l < L; ++l) g2s.push_back(squared_norm(gxs[l], gys[l]));
}

// ---------------------------------------------------------------------------------------
// coarse_to_fine.rs:73-89 `prune_with_thresh`.  `sort_unstable_by` on 4 elements is an
// insertion sort in Rust's std (len <= 20) and therefore behaves stably: among equal values
// the element with the larger original index ends up later.  The comparison is in T = u16,
// so `y + thresh` wraps like release-mode Rust.
void prune_with_thresh(uint16_t thresh, uint16_t a, uint16_t b, uint16_t c, uint16_t d, bool out[4]) {
    std::pair<uint16_t, int> temp[4] = {{a, 0}, {b, 1}, {c, 2}, {d, 3}};
    for (int i = 1; i < 4; ++i) {  // insertion sort, strict "less" shifts only
        auto key = temp[i];
        int j = i - 1;
        while (j >= 0 && key.first < temp[j].first) {
            temp[j + 1] = temp[j];
            --j;
        }
        temp[j + 1] = key;
    }
    const int first = temp[3].second;
    const uint16_t x = temp[2].first;
    const int second = temp[2].second;
    const uint16_t y = temp[1].first;
    out[0] = out[1] = out[2] = out[3] = false;
    out[first] = true;
    if (x > uint16_t(y + thresh)) out[second] = true;
}

// coarse_to_fine.rs:37-62 `select_2x2_bloc`.
Mat<uint8_t> select_2x2_bloc(uint16_t thresh, const Mat<uint8_t>& pre_mask, const Mat<uint16_t>& mat) {
    Mat<uint8_t> mask(mat.rows, mat.cols, 0);
    for (int j = 0; j < pre_mask.cols; ++j)
        for (int i = 0; i < pre_mask.rows; ++i)
            if (pre_mask(i, j)) {
                bool ok[4];
                prune_with_thresh(thresh, mat(2 * i, 2 * j), mat(2 * i + 1, 2 * j), mat(2 * i, 2 * j + 1),
                                  mat(2 * i + 1, 2 * j + 1), ok);
                mask(2 * i, 2 * j) = ok[0];
                mask(2 * i + 1, 2 * j) = ok[1];
                mask(2 * i, 2 * j + 1) = ok[2];
                mask(2 * i + 1, 2 * j + 1) = ok[3];
            }
    return mask;
}

// coarse_to_fine.rs:15-32 `select`: input finest-first; all-true at the coarsest level; returns
// masks coarsest-first (the Tracker pops the last = finest).
std::vector<Mat<uint8_t>> c2f_select(uint16_t thresh, const std::vector<Mat<uint16_t>>& g2) {
    std::vector<Mat<uint8_t>> masks;
    masks.emplace_back(g2.back().rows, g2.back().cols, uint8_t(1));
    for (int l = int(g2.size()) - 2; l >= 0; --l) masks.push_back(select_2x2_bloc(thresh, masks.back(), g2[l]));
    return masks;
}

// ---------------------------------------------------------------------------------------
// inverse_depth.rs:12-20 `InverseDepth`: Unknown / Discarded / WithVariance(idepth, weight).
struct IDepth {
    uint8_t kind = 0;  // 0 Unknown, 1 Discarded, 2 WithVariance
    Float d = 0, v = 0;
};

// inverse_depth.rs:24-29 `from_depth`.
inline IDepth from_depth(Float scale, uint16_t depth, Float variance) {
    IDepth r;
    if (depth != 0) {
        r.kind = 2;
        r.d = scale / Float(depth);
        r.v = variance;
    }
    return r;
}

// inverse_depth.rs:49-66 `fuse` + :81-98 `strategy_dso_mean`: known children in (a,b,c,d) order,
// weighted mean evaluated left to right in f32.
inline IDepth fuse_dso_mean(IDepth a, IDepth b, IDepth c, IDepth d) {
    Float ds[4], vs[4];
    int n = 0;
    for (const IDepth* p : {&a, &b, &c, &d})
        if (p->kind == 2) {
            ds[n] = p->d;
            vs[n] = p->v;
            ++n;
        }
    IDepth r;
    switch (n) {
        case 1: r.kind = 2; r.d = ds[0]; r.v = vs[0]; break;
        case 2: { const Float sum = vs[0] + vs[1];
                  r.kind = 2; r.d = (ds[0] * vs[0] + ds[1] * vs[1]) / sum; r.v = sum; break; }
        case 3: { const Float sum = vs[0] + vs[1] + vs[2];
                  r.kind = 2; r.d = (ds[0] * vs[0] + ds[1] * vs[1] + ds[2] * vs[2]) / sum; r.v = sum; break; }
        case 4: { const Float sum = vs[0] + vs[1] + vs[2] + vs[3];
                  r.kind = 2; r.d = (ds[0] * vs[0] + ds[1] * vs[1] + ds[2] * vs[2] + ds[3] * vs[3]) / sum;
                  r.v = sum; break; }
        default: break;
    }
    return r;
}

// inverse_depth.rs:105-152 `strategy_statistically_similar` behind `fuse`: the known children merge only if every one of
// them lies within one (fused) standard deviation of the fused value; otherwise the bloc is Discarded (kind 1), which
// every consumer treats like Unknown (`with_variance`, :68-74; `extract_z`, inverse_compositional.rs:270-276).
inline IDepth fuse_statistically_similar(IDepth a, IDepth b, IDepth c, IDepth d) {
    Float ds[4], vs[4];
    int n = 0;
    for (const IDepth* p : {&a, &b, &c, &d})
        if (p->kind == 2) {
            ds[n] = p->d;
            vs[n] = p->v;
            ++n;
        }
    IDepth r;
    Float new_d = 0, new_v = 0;
    switch (n) {
        case 1: r.kind = 2; r.d = ds[0]; r.v = 2.0f * vs[0]; return r;
        case 2: new_d = (ds[0] * vs[1] + ds[1] * vs[0]) / (vs[0] + vs[1]);
                new_v = (vs[0] + vs[1]) / 2.0f;
                break;
        case 3: { const Float v12 = vs[0] * vs[1], v13 = vs[0] * vs[2], v23 = vs[1] * vs[2];
                  new_d = (ds[0] * v23 + ds[1] * v13 + ds[2] * v12) / (v12 + v13 + v23);
                  new_v = 2.0f * (vs[0] + vs[1] + vs[2]) / 9.0f;
                  break; }
        case 4: { const Float v123 = vs[0] * vs[1] * vs[2], v234 = vs[1] * vs[2] * vs[3], v341 = vs[2] * vs[3] * vs[0],
                              v412 = vs[3] * vs[0] * vs[1];
                  const Float sum = v123 + v234 + v341 + v412;
                  new_d = (ds[0] * v234 + ds[1] * v341 + ds[2] * v412 + ds[3] * v123) / sum;
                  new_v = (vs[0] + vs[1] + vs[2] + vs[3]) / 8.0f;
                  break; }
        default: return r;
    }
    bool similar = true;
    for (int k = 0; k < n; ++k) {
        const Float e = ds[k] - new_d;
        similar = similar && (e * e < new_v);
    }
    if (similar) {
        r.kind = 2;
        r.d = new_d;
        r.v = new_v;
    } else {
        r.kind = 1;
    }
    return r;
}

// ---------------------------------------------------------------------------------------
// camera.rs:84-140 `Intrinsics`.
struct Intrinsics {
    Float cx, cy, fx, fy, s;
    // camera.rs:115-123 `half_res`.
    Intrinsics half_res() const {
        return {(cx + 0.5f) / 2.0f - 0.5f, (cy + 0.5f) / 2.0f - 0.5f, 0.5f * fx, 0.5f * fy, s};
    }
};
// camera.rs:106-108 `multi_res` = limited_sequence(n, self, half_res) (never yields None).
std::vector<Intrinsics> intrinsics_multi_res(Intrinsics k, int n) {
    std::vector<Intrinsics> v{k};
    for (int len = 1; len < n; ++len) v.push_back(v.back().half_res());
    return v;
}

struct V3 { Float x, y, z; };
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator*(V3 a, Float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 cross(V3 a, V3 b) {  // nalgebra Matrix::cross for 3x1
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

// Quaternion stored like nalgebra's coords (i, j, k, w).
struct Quat { Float i, j, k, w; };

// nalgebra 0.17 quaternion_ops.rs, `&Quaternion * &Quaternion` (Hamilton product).
inline Quat qmul(Quat a, Quat b) {
    return {a.w * b.i + a.i * b.w + a.j * b.k - a.k * b.j,
            a.w * b.j - a.i * b.k + a.j * b.w + a.k * b.i,
            a.w * b.k + a.i * b.j - a.j * b.i + a.k * b.w,
            a.w * b.w - a.i * b.i - a.j * b.j - a.k * b.k};
}
// nalgebra 0.17 `&UnitQuaternion * &Vector3`: t = 2 (v x p); t*w + v x t + p.
inline V3 qrot(Quat q, V3 p) {
    const V3 v{q.i, q.j, q.k};
    const V3 t = cross(v, p) * 2.0f;
    const V3 c = cross(v, t);
    return t * q.w + c + p;
}
// nalgebra dot() special case for static 4-vectors: (x0y0 + x2y2) + (x1y1 + x3y3).
inline Float qnorm2(Quat q) {
    Float a = q.i * q.i, b = q.j * q.j;
    const Float c = q.k * q.k, d = q.w * q.w;
    a += c;
    b += d;
    return a + b;
}
// `UnitQuaternion::from_quaternion` = Unit::new_normalize: every coordinate / norm.
inline Quat qnormalize(Quat q) {
    const Float n = std::sqrt(qnorm2(q));
    return {q.i / n, q.j / n, q.k / n, q.w / n};
}

// Isometry3<f32>: translation then rotation.
struct Iso {
    V3 t{0, 0, 0};
    Quat q{0, 0, 0, 1};
};
// nalgebra isometry_ops.rs `Isometry * Isometry`: t = t1 + q1*t2, q = q1 q2.
inline Iso iso_mul(const Iso& a, const Iso& b) {
    const V3 shift = qrot(a.q, b.t);
    return {a.t + shift, qmul(a.q, b.q)};
}
// `Isometry::inverse`: q^-1 = conj(q), t' = q^-1 * (-t).
inline Iso iso_inv(const Iso& a) {
    const Quat qi{-a.q.i, -a.q.j, -a.q.k, a.q.w};
    return {qrot(qi, V3{-a.t.x, -a.t.y, -a.t.z}), qi};
}
// `Isometry * Point3`: translation * (rotation * p) = (q*p) + t.
inline V3 iso_apply(const Iso& a, V3 p) { return qrot(a.q, p) + a.t; }

inline Iso to_iso(const ref_pose& p) { return {{p.t[0], p.t[1], p.t[2]}, {p.q[0], p.q[1], p.q[2], p.q[3]}}; }
inline ref_pose from_iso(const Iso& a) {
    ref_pose p;
    p.t[0] = a.t.x; p.t[1] = a.t.y; p.t[2] = a.t.z;
    p.q[0] = a.q.i; p.q[1] = a.q.j; p.q[2] = a.q.k; p.q[3] = a.q.w;
    return p;
}

// 3x3 matrix helper for se3::exp (row-major storage here; arithmetic is element-wise).
struct M3 { Float m[3][3]; };
// so3.rs:27-33 `hat`.
inline M3 so3_hat(V3 w) { return {{{0.0f, -w.z, w.y}, {w.z, 0.0f, -w.x}, {-w.y, w.x, 0.0f}}}; }
// so3.rs:38-50 `hat_2`.
inline M3 so3_hat2(V3 w) {
    const Float w11 = w.x * w.x, w12 = w.x * w.y, w13 = w.x * w.z;
    const Float w22 = w.y * w.y, w23 = w.y * w.z, w33 = w.z * w.z;
    return {{{-w22 - w33, w12, w13}, {w12, -w11 - w33, w23}, {w13, w23, -w11 - w22}}};
}

constexpr Float EPS_TAYLOR = 1e-2f;                    // se3.rs:19
constexpr Float EPS_TAYLOR_2 = EPS_TAYLOR * EPS_TAYLOR;  // se3.rs:20
constexpr Float PI_F = 3.14159265358979323846f;

// se3.rs:65-95 `exp`.  V = I + c1*Omega + c2*Omega^2 (element-wise sums left to right),
// t = V * v with nalgebra's gemv order ((V_i0 v0) + V_i1 v1) + V_i2 v2,
// q = normalize(real, imag * w).
Iso se3_exp(const Float xi[6]) {
    const V3 v{xi[0], xi[1], xi[2]}, w{xi[3], xi[4], xi[5]};
    const Float theta_2 = w.x * w.x + w.y * w.y + w.z * w.z;  // norm_squared, static 3-vector: a+b+c
    const M3 om = so3_hat(w), om2 = so3_hat2(w);
    Float real_factor, imag_factor, c1, c2;
    if (theta_2 < EPS_TAYLOR_2) {
        real_factor = 1.0f - 0.125f * theta_2;
        imag_factor = 0.5f - (1.0f / 48.0f) * theta_2;
        c1 = 0.5f - (1.0f / 24.0f) * theta_2;
        c2 = (1.0f / 6.0f) - (1.0f / 120.0f) * theta_2;
    } else {
        const Float theta = std::sqrt(theta_2);
        const Float half_theta = 0.5f * theta;
        real_factor = std::cos(half_theta);
        imag_factor = std::sin(half_theta) / theta;
        c1 = (1.0f - std::cos(theta)) / theta_2;
        c2 = (theta - std::sin(theta)) / (theta * theta_2);
    }
    Float V[3][3];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) V[r][c] = ((r == c ? 1.0f : 0.0f) + c1 * om.m[r][c]) + c2 * om2.m[r][c];
    Iso out;
    out.t = {(V[0][0] * v.x + V[0][1] * v.y) + V[0][2] * v.z, (V[1][0] * v.x + V[1][1] * v.y) + V[1][2] * v.z,
             (V[2][0] * v.x + V[2][1] * v.y) + V[2][2] * v.z};
    out.q = qnormalize(Quat{imag_factor * w.x, imag_factor * w.y, imag_factor * w.z, real_factor});
    return out;
}

inline V3 m3_mul(const Float V[3][3], V3 v) {
    return {(V[0][0] * v.x + V[0][1] * v.y) + V[0][2] * v.z, (V[1][0] * v.x + V[1][1] * v.y) + V[1][2] * v.z,
            (V[2][0] * v.x + V[2][1] * v.y) + V[2][2] * v.z};
}

// se3.rs:99-130 `log` (utility; used by tests and error metrics only).
void se3_log(const Iso& iso, Float xi[6]) {
    const V3 im{iso.q.i, iso.q.j, iso.q.k};
    const Float imag_norm_2 = im.x * im.x + im.y * im.y + im.z * im.z;
    const Float real_factor = iso.q.w;
    V3 w;
    Float coef_omega_2;
    if (imag_norm_2 < EPS_TAYLOR_2) {
        const Float theta_by_imag_norm = 2.0f / real_factor;
        w = im * theta_by_imag_norm;
        const Float x_2 = imag_norm_2 / (real_factor * real_factor);
        coef_omega_2 = (1.0f / 12.0f) * (1.0f + (1.0f / 15.0f) * x_2);
    } else {
        const Float imag_norm = std::sqrt(imag_norm_2);
        Float theta;
        if (std::fabs(real_factor) < EPS_TAYLOR) {
            const Float alpha = std::fabs(real_factor) / imag_norm;
            const Float sgn = std::signbit(real_factor) ? -1.0f : 1.0f;  // f32::signum
            theta = sgn * (PI_F - 2.0f * alpha);
        } else {
            theta = 2.0f * std::atan(imag_norm / real_factor);
        }
        const Float theta_2 = theta * theta;
        w = im * (theta / imag_norm);
        coef_omega_2 = (1.0f - 0.5f * theta * real_factor / imag_norm) / theta_2;
    }
    const M3 om = so3_hat(w), om2 = so3_hat2(w);
    Float Vi[3][3];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) Vi[r][c] = ((r == c ? 1.0f : 0.0f) - 0.5f * om.m[r][c]) + coef_omega_2 * om2.m[r][c];
    const V3 v = m3_mul(Vi, iso.t);
    xi[0] = v.x; xi[1] = v.y; xi[2] = v.z; xi[3] = w.x; xi[4] = w.y; xi[5] = w.z;
}

// so3.rs:61-77 `exp`.
Quat so3_exp(V3 w) {
    const Float theta_2 = w.x * w.x + w.y * w.y + w.z * w.z;
    Float real_factor, imag_factor;
    if (theta_2 < EPS_TAYLOR_2) {
        real_factor = 1.0f - 0.125f * theta_2;
        imag_factor = 0.5f - (1.0f / 48.0f) * theta_2;
    } else {
        const Float theta = std::sqrt(theta_2);
        const Float half_theta = 0.5f * theta;
        real_factor = std::cos(half_theta);
        imag_factor = std::sin(half_theta) / theta;
    }
    return qnormalize(Quat{imag_factor * w.x, imag_factor * w.y, imag_factor * w.z, real_factor});
}

// so3.rs:81-99 `log`.
V3 so3_log(Quat q) {
    const V3 im{q.i, q.j, q.k};
    const Float imag_norm_2 = im.x * im.x + im.y * im.y + im.z * im.z;
    const Float real_factor = q.w;
    if (imag_norm_2 < EPS_TAYLOR_2) {
        return im * (2.0f / real_factor);
    } else if (std::fabs(real_factor) < EPS_TAYLOR) {
        const Float imag_norm = std::sqrt(imag_norm_2);
        const Float alpha = std::fabs(real_factor) / imag_norm;
        const Float sgn = std::signbit(real_factor) ? -1.0f : 1.0f;
        const Float theta = sgn * (PI_F - 2.0f * alpha);
        return im * (theta / imag_norm);
    } else {
        const Float imag_norm = std::sqrt(imag_norm_2);
        const Float theta = 2.0f * std::atan(imag_norm / real_factor);
        return im * (theta / imag_norm);
    }
}

// ---------------------------------------------------------------------------------------
// camera.rs:135-140 `back_project`, :126-132 `project`; lm_optimizer.rs:213-219 `warp`.
inline void warp(const Iso& model, Float x, Float y, Float _z, const Intrinsics& k, Float& u, Float& v) {
    const Float z = 1.0f / _z;
    const Float Y = (y - k.cy) * z / k.fy;
    const Float X = ((x - k.cx) * z - k.s * Y) / k.fx;
    const V3 p2 = iso_apply(model, V3{X, Y, z});
    const Float px = k.fx * p2.x + k.s * p2.y + k.cx * p2.z;
    const Float py = k.fy * p2.y + k.cy * p2.z;
    u = px / p2.z;
    v = py / p2.z;
}

// lm_optimizer.rs:227-251 `interpolate`; inside iff 0 <= floor(x) < W-2 and 0 <= floor(y) < H-2.
inline bool interpolate(Float x, Float y, const uint8_t* image, int rows, int cols, Float& out) {
    const Float u = std::floor(x), v = std::floor(y);
    if (u >= 0.0f && u < Float(cols - 2) && v >= 0.0f && v < Float(rows - 2)) {
        const size_t u0 = size_t(u), v0 = size_t(v);
        const size_t u1 = u0 + 1, v1 = v0 + 1;
        const Float vu00 = Float(image[u0 * rows + v0]);
        const Float vu10 = Float(image[u0 * rows + v1]);
        const Float vu01 = Float(image[u1 * rows + v0]);
        const Float vu11 = Float(image[u1 * rows + v1]);
        const Float a = x - u, b = y - v;
        out = (1.0f - b) * (1.0f - a) * vu00 + b * (1.0f - a) * vu10 + (1.0f - b) * a * vu01 + b * a * vu11;
        return true;
    }
    return false;
}

// inverse_compositional.rs:313-341 `warp_jacobian_at`.
inline void warp_jacobian_at(Float gu, Float gv, Float u, Float v, Float _z, const Intrinsics& k, Float J[6]) {
    const Float cu = k.cx, cv = k.cy, fu = k.fx, fv = k.fy, s = k.s;
    const Float a = u - cu;
    const Float b = v - cv;
    const Float c = a * fv - s * b;
    const Float _fv = 1.0f / fv;
    const Float _fuv = 1.0f / (fu * fv);
    J[0] = gu * _z * fu;
    J[1] = _z * (gu * s + gv * fv);
    J[2] = -_z * (gu * a + gv * b);
    J[3] = gu * (-a * b * _fv - s) + gv * (-b * b * _fv - fv);
    J[4] = gu * (a * c * _fuv + fu) + gv * (b * c * _fuv);
    J[5] = gu * (-fu * fu * b + s * c) * _fuv + gv * (c / fu);
}

// nalgebra 0.17 `Matrix6::cholesky()` (left-looking LL^T on the lower triangle, returns None on
// a pivot that is not > 0 — zero and NaN both fail) and `Cholesky::solve` (forward substitution
// column by column, then transposed back substitution with a sequential dot).
struct Mat6 { Float m[6][6]; };
bool cholesky6(Mat6& A) {
    for (int j = 0; j < 6; ++j) {
        for (int k = 0; k < j; ++k) {
            const Float factor = -A.m[j][k];
            for (int i = j; i < 6; ++i) A.m[i][j] = factor * A.m[i][k] + A.m[i][j];  // axpy(factor, col_k, 1)
        }
        const Float diag = A.m[j][j];
        if (diag > 0.0f) {
            const Float denom = std::sqrt(diag);
            A.m[j][j] = denom;
            for (int i = j + 1; i < 6; ++i) A.m[i][j] /= denom;
            continue;
        }
        return false;
    }
    return true;
}
void cholesky6_solve(const Mat6& L, Float b[6]) {
    for (int i = 0; i < 6; ++i) {  // solve_lower_triangular_mut
        const Float coeff = b[i] / L.m[i][i];
        b[i] = coeff;
        for (int r = i + 1; r < 6; ++r) b[r] = (-coeff) * L.m[r][i] + b[r];
    }
    for (int i = 5; i >= 0; --i) {  // tr_solve_lower_triangular_mut
        Float dot = 0.0f;
        for (int r = i + 1; r < 6; ++r) dot += L.m[r][i] * b[r];
        b[i] = (b[i] - dot) / L.m[i][i];
    }
}

// lm_optimizer.rs:198-209 `renormalize`: q <- 0.5 (3 - |q|^2) q, translation untouched.
inline Iso renormalize(Iso m) {
    const Float sq = qnorm2(m.q);
    const Float f = 0.5f * (3.0f - sq);
    m.q = {f * m.q.i, f * m.q.j, f * m.q.k, f * m.q.w};
    return m;
}

// ---------------------------------------------------------------------------------------
// inverse_compositional.rs:64-70 `MultiresData` (+ the intermediate maps kept for tests).
struct Keyframe {
    std::vector<Intrinsics> intrinsics;
    std::vector<Mat<uint8_t>> img;
    std::vector<std::vector<std::pair<uint32_t, uint32_t>>> coords;  // (x = col, y = row)
    std::vector<std::vector<Float>> idepth;
    std::vector<std::vector<std::array<Float, 6>>> jac;
    std::vector<std::vector<std::array<Float, 36>>> hes;  // full 6x6 per point, like the reference
    std::vector<Mat<IDepth>> idepth_maps;
    Mat<uint8_t> mask0;
    Float huber_delta = 0.0f;  // extension (not in the reference): > 0 switches eval_energy / compute_eval_data to Huber weights
};

// Huber loss rho(r) (r^2 inside |r| <= delta, delta (2|r| - delta) outside) and its IRLS weight min(1, delta / |r|).
inline Float huber_rho(Float r, Float delta) {
    const Float ar = std::fabs(r);
    return ar <= delta ? r * r : delta * (2.0f * ar - delta);
}
inline Float huber_weight(Float r, Float delta) {
    const Float ar = std::fabs(r);
    return ar <= delta ? 1.0f : delta / ar;
}

int dso_select_impl(const Mat<uint16_t>& gradients, int nb_target, int nb_iterations_left, uint64_t seed,
                    Mat<uint8_t>& mask, int* used_random);

// inverse_compositional.rs:105-161 `precompute_multires_data`.
std::unique_ptr<Keyframe> precompute_multires_data(const ref_config& cfg, const Mat<uint16_t>& depth,
                                                   std::vector<Intrinsics> intrinsics,
                                                   std::vector<Mat<uint8_t>> img_multires) {
    auto kf = std::make_unique<Keyframe>();
    kf->huber_delta = cfg.huber_delta;
    std::vector<Mat<int16_t>> gxs, gys;
    std::vector<Mat<uint16_t>> g2s;
    gradients_tracker(img_multires, gxs, gys, g2s, cfg.gradient_operator == 1);

    // :120-125 candidates::select(...).pop()  (extensions: dense = all pixels; dso at level 0)
    Mat<uint8_t> mask0;
    if (cfg.candidate_mode == 1) {
        mask0 = Mat<uint8_t>(img_multires[0].rows, img_multires[0].cols, uint8_t(1));
    } else if (cfg.candidate_mode == 2) {
        // examples/candidates_dso.rs:42: gradient magnitude = sqrt(squared_norm_direct) as u16
        Mat<uint16_t> g2 = squared_norm_direct(img_multires[0]);
        Mat<uint16_t> mag(g2.rows, g2.cols);
        for (size_t k = 0; k < g2.size(); ++k) mag.d[k] = uint16_t(std::sqrt(Float(g2.d[k])));
        int used_random = 0;
        // nb_iterations_left = 2 as in examples/candidates_dso.rs:46
        dso_select_impl(mag, int(cfg.dso_nb_target ? cfg.dso_nb_target : 2000), 2, 0x9E3779B97F4A7C15ull, mask0,
                        &used_random);
    } else {
        mask0 = c2f_select(uint16_t(cfg.candidates_diff_threshold), g2s).back();
    }

    // :128-134 helper::zip_mask_map(depth, mask, Unknown, from_depth)  (helper.rs:40-47)
    Mat<IDepth> id0(depth.rows, depth.cols);
    for (size_t k = 0; k < id0.size(); ++k)
        id0.d[k] = mask0.d[k] ? from_depth(cfg.depth_scale, depth.d[k], cfg.idepth_variance) : IDepth{};
    // :135-138 idepth pyramid with fuse(strategy_dso_mean); the other strategy the reference ships is selectable
    const bool similar = cfg.idepth_fusion == 1;
    auto idepth_multires = limited_sequence(int(cfg.nb_levels), std::move(id0), [similar](const Mat<IDepth>& m, Mat<IDepth>& o) {
        return similar ? halve<IDepth, IDepth>(m, fuse_statistically_similar, o) : halve<IDepth, IDepth>(m, fuse_dso_mean, o);
    });

    const size_t L = idepth_multires.size();
    kf->coords.resize(L);
    kf->idepth.resize(L);
    kf->jac.resize(L);
    kf->hes.resize(L);
    for (size_t l = 0; l < L; ++l) {
        // :260-279 `extract_z`: column-major scan, coordinates (u = col, v = row)
        const Mat<IDepth>& m = idepth_multires[l];
        for (int c = 0; c < m.cols; ++c)
            for (int r = 0; r < m.rows; ++r)
                if (m(r, c).kind == 2) {
                    kf->coords[l].emplace_back(uint32_t(c), uint32_t(r));
                    kf->idepth[l].push_back(m(r, c).d);
                }
        // :284-306 `warp_jacobians`, :345-348 `hessians_vec` (j * j^T)
        const size_t n = kf->coords[l].size();
        kf->jac[l].resize(n);
        kf->hes[l].resize(n);
        for (size_t p = 0; p < n; ++p) {
            const uint32_t u = kf->coords[l][p].first, v = kf->coords[l][p].second;
            const Float gu = Float(gxs[l](int(v), int(u))), gv = Float(gys[l](int(v), int(u)));
            Float J[6];
            warp_jacobian_at(gu, gv, Float(u), Float(v), kf->idepth[l][p], intrinsics[l], J);
            for (int a = 0; a < 6; ++a) kf->jac[l][p][a] = J[a];
            for (int a = 0; a < 6; ++a)
                for (int b = 0; b < 6; ++b) kf->hes[l][p][a * 6 + b] = J[a] * J[b];
        }
    }
    kf->intrinsics = std::move(intrinsics);
    kf->img = std::move(img_multires);
    kf->idepth_maps = std::move(idepth_multires);
    kf->mask0 = std::move(mask0);
    return kf;
}

// Test-only switch (ref_set_accum_f64): accumulate energy / gradient / hessian in f64 instead of the reference's
// sequential f32.  Used to separate "GPU differs from the reference's arithmetic" from "the reference's own
// sequential-f32 sums are noisy" (they absorb small terms once the running sum passes 2^24..2^25).
thread_local bool g_accum_f64 = false;

struct EvalData {
    Mat6 hessian;
    Float gradient[6];
    Float energy;
    Iso model;
};
struct Precomputed {
    Float energy;
    std::vector<uint32_t> inside_indices;
    std::vector<Float> residuals;
};

// lm_optimizer.rs:68-87 `eval_energy`.
Precomputed eval_energy(const Keyframe& kf, int lvl, const uint8_t* image, int rows, int cols, const Iso& model) {
    Precomputed pre;
    Float energy_sum = 0.0f;
    double energy_sum64 = 0.0;
    const auto& coords = kf.coords[lvl];
    const auto& zs = kf.idepth[lvl];
    const Mat<uint8_t>& tmpl = kf.img[lvl];
    const Intrinsics& k = kf.intrinsics[lvl];
    for (size_t idx = 0; idx < coords.size(); ++idx) {
        const uint32_t x = coords[idx].first, y = coords[idx].second;
        Float u, v, im;
        warp(model, Float(x), Float(y), zs[idx], k, u, v);
        if (interpolate(u, v, image, rows, cols, im)) {
            const Float r = im - Float(tmpl(int(y), int(x)));
            const Float e = kf.huber_delta > 0.0f ? huber_rho(r, kf.huber_delta) : r * r;
            energy_sum += e;
            energy_sum64 += kf.huber_delta > 0.0f ? double(e) : double(r) * double(r);
            pre.residuals.push_back(r);
            pre.inside_indices.push_back(uint32_t(idx));
        }
    }
    pre.energy = g_accum_f64 ? Float(energy_sum64 / double(pre.residuals.size())) : energy_sum / Float(pre.residuals.size());
    return pre;
}

// lm_optimizer.rs:90-107 `compute_eval_data`: gradient += jac * r; hessian += hes (sequential f32).
EvalData compute_eval_data(const Keyframe& kf, int lvl, const Iso& model, const Precomputed& pre) {
    EvalData e;
    std::memset(&e.hessian, 0, sizeof(e.hessian));
    for (int a = 0; a < 6; ++a) e.gradient[a] = 0.0f;
    if (g_accum_f64) {
        double gd[6] = {0}, Hd[36] = {0};
        for (size_t i = 0; i < pre.inside_indices.size(); ++i) {
            const auto& jac = kf.jac[lvl][pre.inside_indices[i]];
            const double w = kf.huber_delta > 0.0f ? double(huber_weight(pre.residuals[i], kf.huber_delta)) : 1.0;
            const double r = w * double(pre.residuals[i]);
            for (int a = 0; a < 6; ++a) gd[a] += double(jac[a]) * r;
            for (int a = 0; a < 6; ++a)
                for (int b = 0; b < 6; ++b) Hd[a * 6 + b] += w * double(jac[a]) * double(jac[b]);
        }
        for (int a = 0; a < 6; ++a) e.gradient[a] = Float(gd[a]);
        for (int a = 0; a < 6; ++a)
            for (int b = 0; b < 6; ++b) e.hessian.m[a][b] = Float(Hd[a * 6 + b]);
        e.energy = pre.energy;
        e.model = model;
        return e;
    }
    for (size_t i = 0; i < pre.inside_indices.size(); ++i) {
        const uint32_t idx = pre.inside_indices[i];
        const auto& jac = kf.jac[lvl][idx];
        const auto& hes = kf.hes[lvl][idx];
        const Float w = kf.huber_delta > 0.0f ? huber_weight(pre.residuals[i], kf.huber_delta) : 1.0f;
        const Float r = kf.huber_delta > 0.0f ? w * pre.residuals[i] : pre.residuals[i];
        for (int a = 0; a < 6; ++a) e.gradient[a] += jac[a] * r;
        if (kf.huber_delta > 0.0f) {
            for (int a = 0; a < 6; ++a)
                for (int b = 0; b < 6; ++b) e.hessian.m[a][b] += w * hes[a * 6 + b];
        } else {
            for (int a = 0; a < 6; ++a)
                for (int b = 0; b < 6; ++b) e.hessian.m[a][b] += hes[a * 6 + b];
        }
    }
    e.energy = pre.energy;
    e.model = model;
    return e;
}

struct LMState {
    Float lm_coef;
    EvalData eval_data;
};

// lm_optimizer.rs:123-136 `step`.
bool lm_step(const LMState& st, Iso& new_model) {
    Mat6 hessian = st.eval_data.hessian;
    for (int a = 0; a < 6; ++a) hessian.m[a][a] *= 1.0f + st.lm_coef;
    if (!cholesky6(hessian)) return false;
    Float delta[6];
    for (int a = 0; a < 6; ++a) delta[a] = st.eval_data.gradient[a];
    cholesky6_solve(hessian, delta);
    const Iso delta_warp = se3_exp(delta);
    new_model = renormalize(iso_mul(st.eval_data.model, iso_inv(delta_warp)));
    return true;
}

// math/optimizer.rs:57-70 `iterative_solve` with lm_optimizer.rs:113-192 init / eval / stop_criterion.
// Extension (benchmark only): cfg.fixed_iters = k > 0 runs exactly k step+eval rounds (the stop
// rule becomes nb_iter >= k and the energy-decrease test is skipped); accept/reject unchanged.
int iterative_solve(const ref_config& cfg, const Keyframe& kf, int lvl, const uint8_t* image, int rows, int cols,
                    const Iso& initial_model, Iso& out_model, int& nb_iter_out, Float& energy_out,
                    std::vector<ref_trace_rec>* trace) {
    LMState state;
    state.lm_coef = cfg.lm_coef_init;
    {
        const Precomputed p0 = eval_energy(kf, lvl, image, rows, cols, initial_model);
        state.eval_data = compute_eval_data(kf, lvl, initial_model, p0);
        if (trace) trace->push_back({lvl, 0, state.eval_data.energy, int(p0.residuals.size()), state.lm_coef, 1});
    }
    int nb_iter = 0;
    for (;;) {
        ++nb_iter;
        Iso new_model;
        if (!lm_step(state, new_model)) {
            out_model = state.eval_data.model;
            nb_iter_out = nb_iter;
            energy_out = state.eval_data.energy;
            return 1;  // "Error at Cholesky decomposition of hessian"
        }
        // eval (:140-149)
        const Precomputed pre = eval_energy(kf, lvl, image, rows, cols, new_model);
        const bool is_err = pre.energy > state.eval_data.energy;
        if (trace) trace->push_back({lvl, nb_iter, pre.energy, int(pre.residuals.size()), state.lm_coef, is_err ? 0 : 1});
        // stop_criterion (:156-192)
        const bool too_many = cfg.fixed_iters ? (uint32_t(nb_iter) >= cfg.fixed_iters) : (uint32_t(nb_iter) > cfg.max_iters);
        bool stop;
        if (is_err) {
            if (too_many) {
                stop = true;
            } else {
                state.lm_coef *= cfg.lm_coef_reject_mult;
                stop = false;
            }
        } else {
            EvalData ed = compute_eval_data(kf, lvl, new_model, pre);
            if (too_many) {
                state.eval_data = ed;
                stop = true;
            } else {
                const Float d_energy = state.eval_data.energy - ed.energy;
                stop = cfg.fixed_iters ? false : !(d_energy > cfg.energy_delta_stop);
                state.lm_coef = cfg.lm_coef_accept_mult * state.lm_coef;
                state.eval_data = ed;
            }
        }
        if (stop) {
            out_model = state.eval_data.model;
            nb_iter_out = nb_iter;
            energy_out = state.eval_data.energy;
            return 0;
        }
    }
}

// ---------------------------------------------------------------------------------------
// candidates/dso.rs:98-325.  T = u16.  The random thinning branch (:140-143) uses thread_rng in
// the reference (not reproducible); here it is a seeded splitmix64 — parity unpinned for it.
struct MaxG { uint16_t g; int i, j; };

inline uint64_t splitmix64(uint64_t& s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// dso.rs:307-325 `region_median_gradients` (upper median sorted[len/2]).
Mat<uint16_t> region_median_gradients(const Mat<uint16_t>& g, int size) {
    const int nr = g.rows / size + (g.rows % size ? 1 : 0), nc = g.cols / size + (g.cols % size ? 1 : 0);
    Mat<uint16_t> out(nr, nc);
    std::vector<uint16_t> tmp;
    for (int j = 0; j < nc; ++j)
        for (int i = 0; i < nr; ++i) {
            const int h = std::min(size, g.rows - i * size), w = std::min(size, g.cols - j * size);
            tmp.clear();
            for (int c = 0; c < w; ++c)
                for (int r = 0; r < h; ++r) tmp.push_back(g(i * size + r, j * size + c));
            std::sort(tmp.begin(), tmp.end());
            out(i, j) = tmp[tmp.size() / 2];
        }
    return out;
}

// dso.rs:284-303 `region_thresholds`: a * (mean3x3(median) + b)^2, sum in u16, cast back to u16
// (NumCast: truncation; out-of-range panics in the reference -> ok=false here).
bool region_thresholds(const Mat<uint16_t>& med, Float a, uint16_t b, Mat<uint16_t>& out) {
    out = Mat<uint16_t>(med.rows, med.cols);
    for (int j = 0; j < med.cols; ++j)
        for (int i = 0; i < med.rows; ++i) {
            const int si = std::max(0, i - 1), sj = std::max(0, j - 1);
            const int ei = std::min(med.rows, i + 2), ej = std::min(med.cols, j + 2);
            uint16_t sum = 0;
            int nb = 0;
            for (int jj = sj; jj < ej; ++jj)
                for (int ii = si; ii < ei; ++ii) {
                    sum = uint16_t(sum + med(ii, jj));
                    ++nb;
                }
            const Float t = Float(sum) / Float(nb) + Float(b);
            const Float val = a * t * t;
            if (!(val > -1.0f && val < 65536.0f)) return false;
            out(i, j) = uint16_t(val);
        }
    return true;
}

// dso.rs:193-222 `init_max_gradients`: per block, first strict maximum in column-major scan.
Mat<MaxG> init_max_gradients(const Mat<uint16_t>& g, int bs) {
    const int nr = g.rows / bs + (g.rows % bs ? 1 : 0), nc = g.cols / bs + (g.cols % bs ? 1 : 0);
    Mat<MaxG> out(nr, nc);
    for (int bj = 0; bj < nc; ++bj)
        for (int bi = 0; bi < nr; ++bi) {
            const int si = bi * bs, sj = bj * bs;
            const int ei = std::min(si + bs, g.rows), ej = std::min(sj + bs, g.cols);
            MaxG m{g(si, sj), si, sj};
            for (int j = sj; j < ej; ++j)
                for (int i = si; i < ei; ++i)
                    if (g(i, j) > m.g) m = {g(i, j), i, j};
            out(bi, bj) = m;
        }
    return out;
}

// dso.rs:225-240 `max_of_four_gradients`: g_max(g1, g_max(g2, g_max(g3, g4))), ties keep the left.
inline MaxG max_of_four(MaxG g1, MaxG g2, MaxG g3, MaxG g4) {
    auto gmax = [](MaxG a, MaxG b) { return a.g < b.g ? b : a; };
    return gmax(g1, gmax(g2, gmax(g3, g4)));
}

// dso.rs:156-190 `pick_all_block_candidates` + :246-276 `pick_level_block_candidates`.
size_t pick_all_block_candidates(int base_size, int nb_levels, Float threshold_factor, int regions_size,
                                 const Mat<uint16_t>& thresholds, const Mat<uint16_t>& g, Mat<uint8_t>& picked) {
    auto pyr = limited_sequence(nb_levels, init_max_gradients(g, base_size), [](const Mat<MaxG>& m, Mat<MaxG>& o) {
        return halve<MaxG, MaxG>(m, max_of_four, o);
    });
    Float coef = 1.0f;
    size_t total = 0;
    Mat<uint8_t> mask(pyr[0].rows, pyr[0].cols, uint8_t(1));
    picked = Mat<uint8_t>(g.rows, g.cols, uint8_t(0));
    for (size_t level = 0; level < pyr.size(); ++level) {
        const Mat<MaxG>& mg = pyr[level];
        Mat<uint8_t> next(mask.rows / 2, mask.cols / 2, uint8_t(1));
        for (int j = 0; j < mask.cols / 2 * 2; ++j)
            for (int i = 0; i < mask.rows / 2 * 2; ++i) {
                if (mask(i, j)) {
                    const MaxG m = mg(i, j);
                    const uint16_t th = thresholds(m.i / regions_size, m.j / regions_size);
                    if (Float(m.g) >= coef * Float(th)) {
                        next(i / 2, j / 2) = 0;
                        picked(m.i, m.j) = uint8_t(level + 1);
                        ++total;
                    }
                } else {
                    next(i / 2, j / 2) = 0;
                }
            }
        mask = std::move(next);
        coef *= threshold_factor;
    }
    return total;
}

// dso.rs:98-150 `select` with DEFAULT_* configs (:72-90).
int dso_select_rec(const Mat<uint16_t>& g, int base_size, int iterations_left, int nb_target, uint64_t seed,
                   Mat<uint8_t>& mask, int* used_random) {
    const int region_size = 32;
    const Float coef_a = 1.0f;
    const uint16_t coef_b = 3;
    const int nb_levels = 3;
    const Float threshold_factor = 0.5f;
    const Float low_thresh = 0.8f, high_thresh = 4.0f, random_thresh = 1.1f;

    const Mat<uint16_t> med = region_median_gradients(g, region_size);
    Mat<uint16_t> thresholds;
    if (!region_thresholds(med, coef_a, coef_b, thresholds)) return -1;
    Mat<uint8_t> picked;
    const size_t nb_candidates =
        pick_all_block_candidates(base_size, nb_levels, threshold_factor, region_size, thresholds, g, picked);
    const Float ratio = Float(nb_candidates) / Float(nb_target);
    const Float target_size_f = std::sqrt(ratio) * (Float(base_size) + 1.0f) - 1.0f;
    const int target_size = std::max(1, int(std::round(target_size_f)));
    auto to_mask = [&]() {
        mask = Mat<uint8_t>(picked.rows, picked.cols);
        for (size_t k = 0; k < picked.size(); ++k) mask.d[k] = picked.d[k] > 0;
    };
    if (ratio < low_thresh || ratio > high_thresh) {
        if (target_size != base_size && iterations_left > 0)
            return dso_select_rec(g, target_size, iterations_left - 1, nb_target, seed, mask, used_random);
        to_mask();
    } else if (ratio > random_thresh) {
        if (used_random) *used_random = 1;
        uint64_t s = seed;
        const uint8_t lim = uint8_t(255.0f / ratio);
        mask = Mat<uint8_t>(picked.rows, picked.cols);
        for (size_t k = 0; k < picked.size(); ++k)  // `picked.map` walks column-major
            mask.d[k] = (picked.d[k] > 0) && (uint8_t(splitmix64(s) & 0xFF) <= lim);
    } else {
        to_mask();
    }
    return int(nb_candidates);
}

int dso_select_impl(const Mat<uint16_t>& gradients, int nb_target, int nb_iterations_left, uint64_t seed,
                    Mat<uint8_t>& mask, int* used_random) {
    if (used_random) *used_random = 0;
    return dso_select_rec(gradients, /*base_size*/ 4, nb_iterations_left, nb_target, seed, mask, used_random);
}

template <typename T>
Mat<T> to_col_major(const T* p, int rows, int cols, int layout) {
    Mat<T> m(rows, cols);
    if (layout == 0) {
        std::memcpy(m.d.data(), p, sizeof(T) * m.size());
    } else {  // DMatrix::from_row_slice (vors_track.rs:142, interop.rs:55)
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) m(r, c) = p[size_t(r) * cols + c];
    }
    return m;
}

}  // namespace

// inverse_compositional.rs:30-60 `Tracker`, `State`.
struct ref_keyframe {
    std::unique_ptr<Keyframe> k;
};
struct ref_tracker {
    ref_config cfg;
    int rows, cols, layout;
    ref_keyframe kf;
    double keyframe_depth_ts, keyframe_img_ts;
    Iso keyframe_pose;
    double cur_depth_ts, cur_img_ts;
    Iso cur_pose;
};

extern "C" {

void ref_set_accum_f64(int on) { g_accum_f64 = on != 0; }

void ref_config_default(ref_config* c) {
    std::memset(c, 0, sizeof(*c));
    c->nb_levels = 6;                  // src/bin/vors_track.rs:35
    c->candidates_diff_threshold = 7;  // :36
    c->depth_scale = 5000.0f;          // src/dataset/tum_rgbd.rs:15
    c->fx = 517.306408f; c->fy = 516.469215f; c->cx = 318.643040f; c->cy = 255.313989f; c->skew = 0.0f;  // :31-35
    c->idepth_variance = 0.0001f;      // src/bin/vors_track.rs:39
    c->candidate_mode = 0;
    c->fixed_iters = 0;
    c->lm_coef_init = 0.1f;            // lm_optimizer.rs:115
    c->lm_coef_reject_mult = 10.0f;    // :173
    c->lm_coef_accept_mult = 0.1f;     // :186
    c->energy_delta_stop = 1.0f;       // :179
    c->max_iters = 20;                 // :157
    c->keyframe_flow_threshold = 1.0f; // inverse_compositional.rs:224
    c->device = -1;
    c->team_size = 0;
    c->dso_nb_target = 2000;           // examples/candidates_dso.rs
}

int ref_pyramid_shapes(int rows, int cols, int max_levels, int* out_rows, int* out_cols) {
    int n = 0, r = rows, c = cols;
    for (;;) {
        if (out_rows) out_rows[n] = r;
        if (out_cols) out_cols[n] = c;
        ++n;
        if (!(n < max_levels)) break;
        if (r / 2 == 0 || c / 2 == 0) break;
        r /= 2;
        c /= 2;
    }
    return n;
}

int ref_mean_pyramid(const uint8_t* img, int rows, int cols, int max_levels, uint8_t* out) {
    auto pyr = mean_pyramid(max_levels, mat_from(img, rows, cols));
    size_t off = 0;
    for (auto& m : pyr) {
        std::memcpy(out + off, m.d.data(), m.size());
        off += m.size();
    }
    return int(pyr.size());
}

void ref_gradient_centered(const uint8_t* img, int rows, int cols, int16_t* gx, int16_t* gy) {
    Mat<int16_t> x, y;
    gradient_centered(mat_from(img, rows, cols), x, y);
    std::memcpy(gx, x.d.data(), x.size() * 2);
    std::memcpy(gy, y.d.data(), y.size() * 2);
}

void ref_gradient_scharr(const uint8_t* img, int rows, int cols, int16_t* gx, int16_t* gy) {
    Mat<int16_t> x, y;
    gradient_scharr(mat_from(img, rows, cols), x, y);
    std::memcpy(gx, x.d.data(), x.size() * 2);
    std::memcpy(gy, y.d.data(), y.size() * 2);
}

static std::vector<Mat<uint8_t>> split_levels_u8(const uint8_t* concat, int rows, int cols, int n_levels) {
    std::vector<Mat<uint8_t>> v;
    size_t off = 0;
    int r = rows, c = cols;
    for (int l = 0; l < n_levels; ++l) {
        v.push_back(mat_from(concat + off, r, c));
        off += size_t(r) * c;
        r /= 2;
        c /= 2;
    }
    return v;
}

void ref_gradients_tracker(const uint8_t* pyr_concat, int rows, int cols, int n_levels, int16_t* gx_concat,
                           int16_t* gy_concat, uint16_t* g2_concat) {
    auto pyr = split_levels_u8(pyr_concat, rows, cols, n_levels);
    std::vector<Mat<int16_t>> gxs, gys;
    std::vector<Mat<uint16_t>> g2s;
    gradients_tracker(pyr, gxs, gys, g2s);
    size_t off = 0;
    for (int l = 0; l < n_levels; ++l) {
        const size_t n = gxs[l].size();
        if (gx_concat) std::memcpy(gx_concat + off, gxs[l].d.data(), n * 2);
        if (gy_concat) std::memcpy(gy_concat + off, gys[l].d.data(), n * 2);
        if (g2_concat) std::memcpy(g2_concat + off, g2s[l].d.data(), n * 2);
        off += n;
    }
}

void ref_squared_norm_direct(const uint8_t* img, int rows, int cols, uint16_t* out) {
    auto m = squared_norm_direct(mat_from(img, rows, cols));
    std::memcpy(out, m.d.data(), m.size() * 2);
}

// examples/candidates_coarse-to-fine.rs:55-69: [squared_norm_direct(I0)] ++ gradients_squared_norm(pyr)
// (multires.rs:96-106).
void ref_gradients_squared_norm_example(const uint8_t* pyr_concat, int rows, int cols, int n_levels,
                                        uint16_t* g2_concat) {
    auto pyr = split_levels_u8(pyr_concat, rows, cols, n_levels);
    size_t off = 0;
    auto m0 = squared_norm_direct(pyr[0]);
    std::memcpy(g2_concat, m0.d.data(), m0.size() * 2);
    off += m0.size();
    for (int l = 1; l < n_levels; ++l) {
        Mat<uint16_t> o;
        halve<uint8_t, uint16_t>(pyr[l - 1], bloc_squared_norm, o);
        std::memcpy(g2_concat + off, o.d.data(), o.size() * 2);
        off += o.size();
    }
}

void ref_prune_with_thresh(uint16_t thresh, uint16_t a, uint16_t b, uint16_t c, uint16_t d, uint8_t out[4]) {
    bool ok[4];
    prune_with_thresh(thresh, a, b, c, d, ok);
    for (int i = 0; i < 4; ++i) out[i] = ok[i];
}

void ref_c2f_select(uint16_t thresh, const uint16_t* g2_concat, int rows, int cols, int n_levels,
                    uint8_t* masks_concat) {
    std::vector<Mat<uint16_t>> g2;
    size_t off = 0;
    int r = rows, c = cols;
    for (int l = 0; l < n_levels; ++l) {
        g2.push_back(mat_from(g2_concat + off, r, c));
        off += size_t(r) * c;
        r /= 2;
        c /= 2;
    }
    auto masks = c2f_select(thresh, g2);  // coarsest first
    off = 0;
    for (int l = 0; l < n_levels; ++l) {
        const Mat<uint8_t>& m = masks[size_t(n_levels - 1 - l)];
        std::memcpy(masks_concat + off, m.d.data(), m.size());
        off += m.size();
    }
}

int ref_dso_select(const uint16_t* gradients, int rows, int cols, int nb_target, int nb_iterations_left,
                   uint64_t seed, uint8_t* mask_out, int* used_random_branch) {
    Mat<uint8_t> mask;
    const int n = dso_select_impl(mat_from(gradients, rows, cols), nb_target, nb_iterations_left, seed, mask,
                                  used_random_branch);
    if (n >= 0) std::memcpy(mask_out, mask.d.data(), mask.size());
    return n;
}

ref_keyframe* ref_keyframe_create(const ref_config* cfg, const uint16_t* depth, const uint8_t* img, int rows,
                                  int cols) {
    auto pyr = mean_pyramid(int(cfg->nb_levels), mat_from(img, rows, cols));
    if (pyr.size() < cfg->nb_levels) return nullptr;  // the reference would index out of bounds in track()
    auto intr = intrinsics_multi_res(Intrinsics{cfg->cx, cfg->cy, cfg->fx, cfg->fy, cfg->skew}, int(cfg->nb_levels));
    auto* kf = new ref_keyframe;
    kf->k = precompute_multires_data(*cfg, mat_from(depth, rows, cols), std::move(intr), std::move(pyr));
    return kf;
}
void ref_keyframe_destroy(ref_keyframe* kf) { delete kf; }
int ref_keyframe_levels(const ref_keyframe* kf) { return int(kf->k->img.size()); }
void ref_keyframe_level_shape(const ref_keyframe* kf, int level, int* rows, int* cols) {
    *rows = kf->k->img[level].rows;
    *cols = kf->k->img[level].cols;
}
void ref_keyframe_intrinsics(const ref_keyframe* kf, int level, float out5[5]) {
    const Intrinsics& k = kf->k->intrinsics[level];
    out5[0] = k.fx; out5[1] = k.fy; out5[2] = k.cx; out5[3] = k.cy; out5[4] = k.s;
}
int ref_keyframe_n_points(const ref_keyframe* kf, int level) { return int(kf->k->coords[level].size()); }
void ref_keyframe_points(const ref_keyframe* kf, int level, uint32_t* xy, float* idepth, float* jac) {
    const auto& c = kf->k->coords[level];
    for (size_t p = 0; p < c.size(); ++p) {
        if (xy) { xy[2 * p] = c[p].first; xy[2 * p + 1] = c[p].second; }
        if (idepth) idepth[p] = kf->k->idepth[level][p];
        if (jac) for (int a = 0; a < 6; ++a) jac[6 * p + a] = kf->k->jac[level][p][a];
    }
}
void ref_keyframe_image(const ref_keyframe* kf, int level, uint8_t* out) {
    std::memcpy(out, kf->k->img[level].d.data(), kf->k->img[level].size());
}
void ref_keyframe_mask0(const ref_keyframe* kf, uint8_t* out) {
    std::memcpy(out, kf->k->mask0.d.data(), kf->k->mask0.size());
}
void ref_keyframe_idepth_map(const ref_keyframe* kf, int level, float* idepth, float* weight) {
    const Mat<IDepth>& m = kf->k->idepth_maps[level];
    for (size_t k = 0; k < m.size(); ++k) {
        const bool known = m.d[k].kind == 2;
        if (idepth) idepth[k] = known ? m.d[k].d : std::numeric_limits<float>::quiet_NaN();
        if (weight) weight[k] = known ? m.d[k].v : 0.0f;
    }
}

int ref_eval(const ref_keyframe* kf, int level, const uint8_t* image, int rows, int cols, const ref_pose* model,
             int accum, float* energy, float g[6], float H[36]) {
    const Iso m = to_iso(*model);
    const Precomputed pre = eval_energy(*kf->k, level, image, rows, cols, m);
    if (accum == 0) {
        const EvalData e = compute_eval_data(*kf->k, level, m, pre);
        *energy = e.energy;
        for (int a = 0; a < 6; ++a) g[a] = e.gradient[a];
        for (int a = 0; a < 6; ++a)
            for (int b = 0; b < 6; ++b) H[a * 6 + b] = e.hessian.m[a][b];
    } else {
        double es = 0, gd[6] = {0}, Hd[36] = {0};
        for (size_t i = 0; i < pre.inside_indices.size(); ++i) {
            const uint32_t idx = pre.inside_indices[i];
            const Float hd = kf->k->huber_delta;
            const double w = hd > 0.0f ? double(huber_weight(pre.residuals[i], hd)) : 1.0;
            es += hd > 0.0f ? double(huber_rho(pre.residuals[i], hd)) : double(pre.residuals[i]) * double(pre.residuals[i]);
            const double r = w * double(pre.residuals[i]);
            const auto& J = kf->k->jac[level][idx];
            for (int a = 0; a < 6; ++a) gd[a] += double(J[a]) * r;
            for (int a = 0; a < 6; ++a)
                for (int b = 0; b < 6; ++b) Hd[a * 6 + b] += w * double(J[a]) * double(J[b]);
        }
        *energy = float(es / double(pre.residuals.size()));
        for (int a = 0; a < 6; ++a) g[a] = float(gd[a]);
        for (int a = 0; a < 36; ++a) H[a] = float(Hd[a]);
    }
    return int(pre.residuals.size());
}

int ref_iterative_solve(const ref_config* cfg, const ref_keyframe* kf, int level, const uint8_t* image, int rows,
                        int cols, const ref_pose* init, ref_pose* out, int* n_iter, float* final_energy,
                        ref_trace_rec* trace, int trace_cap, int* trace_len) {
    std::vector<ref_trace_rec> tr;
    Iso om;
    int it = 0;
    Float en = 0;
    const int st = iterative_solve(*cfg, *kf->k, level, image, rows, cols, to_iso(*init), om, it, en,
                                   trace ? &tr : nullptr);
    *out = from_iso(om);
    if (n_iter) *n_iter = it;
    if (final_energy) *final_energy = en;
    if (trace) {
        const int n = std::min<int>(int(tr.size()), trace_cap);
        std::memcpy(trace, tr.data(), sizeof(ref_trace_rec) * size_t(n));
        if (trace_len) *trace_len = n;
    }
    return st;
}

// inverse_compositional.rs:74-100 `Config::init`.
ref_tracker* ref_tracker_create(const ref_config* cfg, double depth_ts, const uint16_t* depth, double img_ts,
                                const uint8_t* img, int rows, int cols, int layout) {
    auto pyr = mean_pyramid(int(cfg->nb_levels), to_col_major(img, rows, cols, layout));
    if (pyr.size() < cfg->nb_levels) return nullptr;
    if (pyr.back().rows < 3 || pyr.back().cols < 3) return nullptr;  // `width - 2` underflow, lm_optimizer.rs:231
    auto* t = new ref_tracker;
    t->cfg = *cfg;
    t->rows = rows; t->cols = cols; t->layout = layout;
    auto intr = intrinsics_multi_res(Intrinsics{cfg->cx, cfg->cy, cfg->fx, cfg->fy, cfg->skew}, int(cfg->nb_levels));
    t->kf.k = precompute_multires_data(*cfg, to_col_major(depth, rows, cols, layout), std::move(intr), std::move(pyr));
    t->keyframe_depth_ts = depth_ts;
    t->keyframe_img_ts = img_ts;
    t->keyframe_pose = Iso{};
    t->cur_depth_ts = depth_ts;
    t->cur_img_ts = img_ts;
    t->cur_pose = Iso{};
    return t;
}

// inverse_compositional.rs:170-240 `Tracker::track`.
int ref_tracker_track(ref_tracker* t, double depth_ts, const uint16_t* depth, double img_ts, const uint8_t* img,
                      ref_track_stats* stats, ref_trace_rec* trace, int trace_cap, int* trace_len) {
    const Keyframe& kf = *t->kf.k;
    const int L = int(t->cfg.nb_levels);
    Iso lm_model = iso_mul(iso_inv(t->cur_pose), t->keyframe_pose);  // :177
    auto img_multires = mean_pyramid(L, to_col_major(img, t->rows, t->cols, t->layout));  // :178
    bool went_well = true;
    std::vector<ref_trace_rec> tr;
    if (stats) std::memset(stats, 0, sizeof(*stats));
    for (int lvl = L - 1; lvl >= 0; --lvl) {  // :181
        Iso om;
        int it = 0;
        Float en = 0;
        const Mat<uint8_t>& im = img_multires[size_t(lvl)];
        const int st = iterative_solve(t->cfg, kf, lvl, im.d.data(), im.rows, im.cols, lm_model, om, it, en,
                                       trace ? &tr : nullptr);
        if (stats && lvl < REF_MAX_LEVELS) {
            stats->n_iters[lvl] = it;
            stats->energy[lvl] = en;
            stats->n_points[lvl] = int(kf.coords[size_t(lvl)].size());
        }
        if (st == 0) {
            lm_model = om;  // :193
        } else {
            went_well = false;  // :195-199
            break;
        }
    }
    t->cur_depth_ts = depth_ts;  // :203-204
    t->cur_img_ts = img_ts;
    if (went_well) t->cur_pose = iso_mul(t->keyframe_pose, iso_inv(lm_model));  // :206-208

    // :210-221 optical flow over the COARSEST level's candidates (`.last()`).
    const auto& coords = kf.coords.back();
    const auto& zs = kf.idepth.back();
    const Intrinsics& k = kf.intrinsics.back();
    Float flow_sum = 0.0f;
    for (size_t p = 0; p < coords.size(); ++p) {
        Float u, v;
        const Float x = Float(coords[p].first), y = Float(coords[p].second);
        warp(lm_model, x, y, zs[p], k, u, v);
        flow_sum += std::fabs(x - u) + std::fabs(y - v);
    }
    const Float optical_flow = flow_sum / Float(zs.size());
    const bool change_keyframe = optical_flow >= t->cfg.keyframe_flow_threshold;  // :224
    if (change_keyframe) {  // :227-239
        auto intr = kf.intrinsics;
        t->kf.k = precompute_multires_data(t->cfg, to_col_major(depth, t->rows, t->cols, t->layout), std::move(intr),
                                           std::move(img_multires));
        t->keyframe_depth_ts = depth_ts;
        t->keyframe_img_ts = img_ts;
        t->keyframe_pose = t->cur_pose;
    }
    if (stats) {
        stats->status = went_well ? 0 : 1;
        stats->keyframe_changed = change_keyframe ? 1 : 0;
        stats->optical_flow = optical_flow;
    }
    if (trace) {
        const int n = std::min<int>(int(tr.size()), trace_cap);
        std::memcpy(trace, tr.data(), sizeof(ref_trace_rec) * size_t(n));
        if (trace_len) *trace_len = n;
    }
    return went_well ? 0 : 1;
}

// inverse_compositional.rs:243-248 `current_frame` (depth timestamp, pose).
void ref_tracker_current_frame(const ref_tracker* t, double* depth_ts, ref_pose* pose) {
    if (depth_ts) *depth_ts = t->cur_depth_ts;
    if (pose) *pose = from_iso(t->cur_pose);
}
void ref_tracker_keyframe_pose(const ref_tracker* t, ref_pose* pose) { *pose = from_iso(t->keyframe_pose); }
const ref_keyframe* ref_tracker_keyframe(const ref_tracker* t) { return &t->kf; }
void ref_tracker_destroy(ref_tracker* t) { delete t; }

void ref_so3_hat(const float w[3], float out9[9]) {
    const M3 m = so3_hat(V3{w[0], w[1], w[2]});
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) out9[r * 3 + c] = m.m[r][c];
}
void ref_so3_hat2(const float w[3], float out9[9]) {
    const M3 m = so3_hat2(V3{w[0], w[1], w[2]});
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) out9[r * 3 + c] = m.m[r][c];
}
// so3.rs:54-56 `vee`: (m32, m13, m21).
void ref_so3_vee(const float m9[9], float w[3]) { w[0] = m9[2 * 3 + 1]; w[1] = m9[0 * 3 + 2]; w[2] = m9[1 * 3 + 0]; }
void ref_so3_exp(const float w[3], float q[4]) {
    const Quat r = so3_exp(V3{w[0], w[1], w[2]});
    q[0] = r.i; q[1] = r.j; q[2] = r.k; q[3] = r.w;
}
void ref_so3_log(const float q[4], float w[3]) {
    const V3 r = so3_log(Quat{q[0], q[1], q[2], q[3]});
    w[0] = r.x; w[1] = r.y; w[2] = r.z;
}
// se3.rs:45-56 `hat`.
void ref_se3_hat(const float xi[6], float o[16]) {
    const float w1 = xi[3], w2 = xi[4], w3 = xi[5];
    const float m[16] = {0.0f, -w3, w2, xi[0], w3, 0.0f, -w1, xi[1], -w2, w1, 0.0f, xi[2], 0.0f, 0.0f, 0.0f, 0.0f};
    std::memcpy(o, m, sizeof(m));
}
// se3.rs:60-62 `vee`: (m14, m24, m34, m32, m13, m21).
void ref_se3_vee(const float m[16], float xi[6]) {
    xi[0] = m[0 * 4 + 3]; xi[1] = m[1 * 4 + 3]; xi[2] = m[2 * 4 + 3];
    xi[3] = m[2 * 4 + 1]; xi[4] = m[0 * 4 + 2]; xi[5] = m[1 * 4 + 0];
}
void ref_se3_exp(const float xi[6], ref_pose* out) { *out = from_iso(se3_exp(xi)); }
void ref_se3_log(const ref_pose* p, float xi[6]) { se3_log(to_iso(*p), xi); }
void ref_pose_mul(const ref_pose* a, const ref_pose* b, ref_pose* out) { *out = from_iso(iso_mul(to_iso(*a), to_iso(*b))); }
void ref_pose_inverse(const ref_pose* a, ref_pose* out) { *out = from_iso(iso_inv(to_iso(*a))); }
void ref_pose_transform(const ref_pose* a, const float p[3], float out[3]) {
    const V3 r = iso_apply(to_iso(*a), V3{p[0], p[1], p[2]});
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
// nalgebra `UnitQuaternion::from_euler_angles(roll, pitch, yaw)` (used by the reference's test
// generators, so3.rs:146, se3.rs:177).
void ref_quat_from_euler(float roll, float pitch, float yaw, float q[4]) {
    const float sr = std::sin(roll * 0.5f), cr = std::cos(roll * 0.5f);
    const float sp = std::sin(pitch * 0.5f), cp = std::cos(pitch * 0.5f);
    const float sy = std::sin(yaw * 0.5f), cy = std::cos(yaw * 0.5f);
    q[3] = cr * cp * cy + sr * sp * sy;
    q[0] = sr * cp * cy - cr * sp * sy;
    q[1] = cr * sp * cy + sr * cp * sy;
    q[2] = cr * cp * sy - sr * sp * cy;
}
int ref_cholesky_solve6(const float H[36], const float g[6], float x[6]) {
    Mat6 A;
    for (int a = 0; a < 6; ++a) for (int b = 0; b < 6; ++b) A.m[a][b] = H[a * 6 + b];
    if (!cholesky6(A)) return 0;
    for (int a = 0; a < 6; ++a) x[a] = g[a];
    cholesky6_solve(A, x);
    return 1;
}
void ref_warp(const ref_pose* model, float x, float y, float idepth, const float k5[5], float uv[2]) {
    warp(to_iso(*model), x, y, idepth, Intrinsics{k5[2], k5[3], k5[0], k5[1], k5[4]}, uv[0], uv[1]);
}
void ref_warp_jacobian_at(float gu, float gv, float u, float v, float idepth, const float k5[5], float out6[6]) {
    warp_jacobian_at(gu, gv, u, v, idepth, Intrinsics{k5[2], k5[3], k5[0], k5[1], k5[4]}, out6);
}
int ref_interpolate(float x, float y, const uint8_t* image, int rows, int cols, float* out) {
    Float o = 0;
    const bool ok = interpolate(x, y, image, rows, cols, o);
    *out = o;
    return ok ? 1 : 0;
}

}  // extern "C"
