// image_kernels.cu — keyframe / frame precompute kernels, sm_100a.
//
// All integer stages are bit-exact restatements of the reference semantics (truncating signed
// division, u16 wrap of `third + thresh`, stable tie-break of the 4-element sort).  Every kernel
// is batched over streams through blockIdx.y (`items` maps the launch index to the stream slab).
#include <algorithm>

#include "vors_device.cuh"

namespace vors {

namespace {

__device__ __forceinline__ int item_of(const int* items, int j) { return items ? items[j] : j; }

struct LevelIntrinsics {
    Intrinsics k[kMaxLevels];
};

// ------------------------------------------------------------------------------------------------
// Row-major (decoder output) -> column-major (internal, nalgebra) transposition; what
// `DMatrix::from_row_slice` does on the CPU in the reference (src/misc/interop.rs:53-56).
template <typename T>
__global__ void k_transpose(const T* __restrict__ in, T* __restrict__ out_slab, size_t out_stride, const int* __restrict__ items,
                            int rows, int cols) {
    __shared__ T tile[32][33];
    const int j = blockIdx.z;
    const T* src = in + size_t(j) * rows * cols;
    T* dst = out_slab + size_t(item_of(items, j)) * out_stride;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
        const int r = r0 + dy, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[dy][threadIdx.x] = src[size_t(r) * cols + c];
    }
    __syncthreads();
    for (int dx = threadIdx.y; dx < 32; dx += blockDim.y) {
        const int c = c0 + dx, r = r0 + threadIdx.x;
        if (r < rows && c < cols) dst[size_t(c) * rows + r] = tile[threadIdx.x][dx];
    }
}

__global__ void k_copy_items_u16(const uint16_t* __restrict__ in, uint16_t* __restrict__ out_slab, size_t out_stride,
                                 const int* __restrict__ items, size_t count) {
    const int j = blockIdx.y;
    const uint16_t* src = in + size_t(j) * count;
    uint16_t* dst = out_slab + size_t(item_of(items, j)) * out_stride;
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += size_t(gridDim.x) * blockDim.x) dst[i] = src[i];
}

// ------------------------------------------------------------------------------------------------
// Row A: one level of `multires::mean_pyramid` (multires.rs:21-31): ((a+b+c+d)/4) as u8 over 2x2
// blocks a=(2i,2j) b=(2i+1,2j) c=(2i,2j+1) d=(2i+1,2j+1); odd last row/col dropped.
__global__ void k_halve_mean(const Geom g, int l, uint8_t* __restrict__ pyr_slab, const int* __restrict__ items) {
    uint8_t* pyr = pyr_slab + size_t(item_of(items, blockIdx.y)) * g.pix_stride;
    const uint8_t* in = pyr + g.off[l - 1];
    uint8_t* out = pyr + g.off[l];
    const int R = g.rows[l], C = g.cols[l], Rin = g.rows[l - 1];
    const int n = R * C;
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < n; o += gridDim.x * blockDim.x) {
        const int x = o / R, y = o - x * R;
        const uint8_t* p = in + size_t(2 * x) * Rin + 2 * y;
        const unsigned a = p[0], b = p[1], c = p[Rin], d = p[Rin + 1];
        out[o] = uint8_t((a + b + c + d) >> 2);
    }
}

// Row A, fused: ALL pyramid levels of a frame in one launch.  One CTA stages a 64x64 level-0 tile in shared memory and
// halves it in place level by level (up to 6 halvings: 64 -> 1), writing every level's part of the tile; the level-0
// image is read from HBM exactly once and the per-frame pyramid costs one launch instead of L-1.
// Floor-halving keeps tiles independent: a level-l pixel exists iff both its children rows/cols exist at level l-1.
constexpr int kPyrTile = 64;
__global__ void __launch_bounds__(256) k_pyramid_fused(const Geom g, int l_first, int l_count, uint8_t* __restrict__ pyr_slab,
                                                        const int* __restrict__ items) {
    __shared__ uint8_t buf[2][kPyrTile * kPyrTile];  // ping-pong: level l in buf[l & 1], column-major tile (y fastest)
    uint8_t* pyr = pyr_slab + size_t(item_of(items, blockIdx.z)) * g.pix_stride;
    const int ty0 = blockIdx.x * kPyrTile, tx0 = blockIdx.y * kPyrTile;  // tile origin at level l_first
    {
        const uint8_t* in = pyr + g.off[l_first];
        const int R = g.rows[l_first], C = g.cols[l_first];
        for (int i = threadIdx.x; i < kPyrTile * kPyrTile; i += blockDim.x) {
            const int x = i / kPyrTile, y = i - x * kPyrTile;
            const int gy = ty0 + y, gx = tx0 + x;
            buf[0][i] = (gy < R && gx < C) ? in[size_t(gx) * R + gy] : uint8_t(0);
        }
    }
    __syncthreads();
    int side = kPyrTile;
    for (int k = 1; k <= l_count; ++k) {
        const int l = l_first + k;
        const int half = side >> 1;
        const uint8_t* src = buf[(k - 1) & 1];
        uint8_t* dst = buf[k & 1];
        uint8_t* out = pyr + g.off[l];
        const int R = g.rows[l], C = g.cols[l];
        const int oy0 = ty0 >> k, ox0 = tx0 >> k;
        for (int i = threadIdx.x; i < half * half; i += blockDim.x) {
            const int x = i / half, y = i - x * half;
            const uint8_t* p = src + (2 * x) * side + 2 * y;
            const unsigned v = (unsigned(p[0]) + p[1] + p[side] + p[side + 1]) >> 2;  // ((a+b+c+d)/4) as u8, multires.rs:24-29
            dst[x * half + y] = uint8_t(v);
            const int gy = oy0 + y, gx = ox0 + x;
            if (gy < R && gx < C) out[size_t(gx) * R + gy] = uint8_t(v);
        }
        __syncthreads();
        side = half;
    }
}

// Rows B, C, D: the Tracker's gradient recipe (inverse_compositional.rs:112-117) for every level in
// one launch.  Level 0: gradient::centered (gradient.rs:15-33), i16 division truncating toward zero,
// 1-px border 0.  Level l >= 1: bloc_x / bloc_y of the level l-1 image (gradient.rs:74-93).
// g2 = (gx*gx + gy*gy) as u16 (gradient.rs:38-44).
// `scharr` (extension, vors_config.gradient_operator = 1): 3x3 Scharr / 32 on every level's own image instead.
// The (gx, gy) pair of pixel (x, y) of level l as the gradient slab stores it: gx | gy << 16, both i16.
__device__ __forceinline__ uint32_t grad_pair_at(const Geom& g, const uint8_t* __restrict__ pyr, int scharr, int l, int x, int y, int* g2 = nullptr) {
    const int R = g.rows[l], C = g.cols[l];
    int gx = 0, gy = 0;
    if (scharr) {
        if (x > 0 && x < C - 1 && y > 0 && y < R - 1) {
            const uint8_t* p = pyr + g.off[l] + size_t(x) * R + y;  // p[dc * R + dr]
            const int tl = p[-R - 1], ml = p[-R], bl = p[-R + 1], tc = p[-1], bc = p[1], tr = p[R - 1], mr = p[R], br = p[R + 1];
            gx = (3 * (tr - tl) + 10 * (mr - ml) + 3 * (br - bl)) / 32;
            gy = (3 * (bl - tl) + 10 * (bc - tc) + 3 * (br - tr)) / 32;
        }
    } else if (l == 0) {
        if (x > 0 && x < C - 1 && y > 0 && y < R - 1) {
            const uint8_t* p = pyr + size_t(x) * R + y;
            gx = (int(p[R]) - int(p[-R])) / 2;  // right - left, C++ `/` truncates like Rust
            gy = (int(p[1]) - int(p[-1])) / 2;  // bottom - top
        }
    } else {
        const int Rin = g.rows[l - 1];
        const uint8_t* p = pyr + g.off[l - 1] + size_t(2 * x) * Rin + 2 * y;
        const int a = p[0], b = p[1], c = p[Rin], d = p[Rin + 1];
        gx = (c + d - a - b) / 2;
        gy = (b - a + d - c) / 2;
    }
    if (g2) *g2 = gx * gx + gy * gy;
    return (uint32_t(gx) & 0xFFFFu) | (uint32_t(gy) << 16);
}

__global__ void k_gradients(const Geom g, int scharr, const uint8_t* __restrict__ pyr_slab, uint32_t* __restrict__ grad_slab,
                            uint16_t* __restrict__ g2_slab, const int* __restrict__ items) {
    const size_t base = size_t(item_of(items, blockIdx.y)) * g.pix_stride;
    const uint8_t* pyr = pyr_slab + base;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < g.pix_total; i += gridDim.x * blockDim.x) {
        int l = 0;
#pragma unroll
        for (int k = 1; k < kMaxLevels; ++k)
            if (k < g.L && i >= g.off[k]) l = k;
        const int o = i - g.off[l];
        const int R = g.rows[l];
        const int x = o / R, y = o - x * R;
        int g2;
        const uint32_t pair = grad_pair_at(g, pyr, scharr, l, x, y, &g2);
        if (grad_slab) grad_slab[base + i] = pair;
        if (g2_slab) g2_slab[base + i] = uint16_t(g2);
    }
}

// ------------------------------------------------------------------------------------------------
// Row E: one level of candidates::coarse_to_fine::select (coarse_to_fine.rs:15-89).  One thread per
// (ceil) parent of level l+1: a selected parent keeps its max child and, if second > third + thresh
// (u16 arithmetic), the second max; everything else (unselected parents, odd border) is false.
// A stable ascending sort of (value, index) pairs is a plain sort of value<<2|index keys.
__device__ __forceinline__ void cswap(unsigned& a, unsigned& b) {
    const unsigned lo = min(a, b), hi = max(a, b);
    a = lo;
    b = hi;
}

__global__ void k_c2f_level(const Geom g, int l, uint16_t thresh, const uint16_t* __restrict__ g2_slab,
                            uint8_t* __restrict__ mask_slab, const int* __restrict__ items) {
    const size_t base = size_t(item_of(items, blockIdx.y)) * g.pix_stride;
    const uint16_t* g2 = g2_slab + base + g.off[l];
    uint8_t* mask = mask_slab + base + g.off[l];
    const uint8_t* pre = (l + 1 == g.L - 1) ? nullptr : mask_slab + base + g.off[l + 1];  // coarsest: all true
    const int R = g.rows[l], C = g.cols[l], Rp = g.rows[l + 1], Cp = g.cols[l + 1];
    const int PR = (R + 1) / 2, PC = (C + 1) / 2;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < PR * PC; t += gridDim.x * blockDim.x) {
        const int px = t / PR, py = t - px * PR;
        const bool real_parent = (py < Rp) && (px < Cp);
        uint8_t ok[4] = {0, 0, 0, 0};
        if (real_parent && (pre == nullptr || pre[size_t(px) * Rp + py])) {
            const uint16_t* p = g2 + size_t(2 * px) * R + 2 * py;
            unsigned k0 = (unsigned(p[0]) << 2) | 0u, k1 = (unsigned(p[1]) << 2) | 1u;
            unsigned k2 = (unsigned(p[R]) << 2) | 2u, k3 = (unsigned(p[R + 1]) << 2) | 3u;
            cswap(k0, k1); cswap(k2, k3); cswap(k0, k2); cswap(k1, k3); cswap(k1, k2);
            const unsigned first = k3 & 3u, second = k2 & 3u;
            const uint16_t x = uint16_t(k2 >> 2), y = uint16_t(k1 >> 2);
            ok[first] = 1;
            if (x > uint16_t(y + thresh)) ok[second] = 1;
        }
        const int y0 = 2 * py, x0 = 2 * px;
        if (y0 < R && x0 < C) mask[size_t(x0) * R + y0] = ok[0];
        if (y0 + 1 < R && x0 < C) mask[size_t(x0) * R + y0 + 1] = ok[1];
        if (y0 < R && x0 + 1 < C) mask[size_t(x0 + 1) * R + y0] = ok[2];
        if (y0 + 1 < R && x0 + 1 < C) mask[size_t(x0 + 1) * R + y0 + 1] = ok[3];
    }
}

// ------------------------------------------------------------------------------------------------
// Row F: helper::zip_mask_map + inverse_depth::from_depth (helper.rs:40-47, inverse_depth.rs:24-29):
// idepth = depth_scale / depth where mask and depth != 0; Unknown is encoded as NaN (weight 0).
__global__ void k_idepth0(const Geom g, const uint16_t* __restrict__ depth_slab, size_t depth_stride,
                          const uint8_t* __restrict__ mask_slab, int dense, float scale, float variance,
                          float* __restrict__ idepth_slab, float* __restrict__ weight_slab, const int* __restrict__ items) {
    const int it = item_of(items, blockIdx.y);
    const size_t base = size_t(it) * g.pix_stride;
    const uint16_t* depth = depth_slab + size_t(it) * depth_stride;
    const int n = g.rows[0] * g.cols[0];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint16_t d = depth[i];
        const bool known = (dense || mask_slab[base + i]) && d != 0;
        idepth_slab[base + i] = known ? scale / float(d) : __int_as_float(0x7fc00000);
        weight_slab[base + i] = known ? variance : 0.0f;
    }
}

// Row G: one halving of the idepth pyramid with inverse_depth::fuse + strategy_dso_mean
// (inverse_depth.rs:49-98): weighted mean of the known children in (a,b,c,d) order, evaluated left
// to right in f32; one known child is copied verbatim; none -> Unknown.
// `similar` selects inverse_depth.rs:105-152 `strategy_statistically_similar` instead: the known children merge only if each
// lies within one fused standard deviation of the fused value, else the bloc is Discarded (stored like Unknown: NaN).
__global__ void k_idepth_halve(const Geom g, int l, int similar, float* __restrict__ idepth_slab, float* __restrict__ weight_slab,
                               const int* __restrict__ items) {
    const size_t base = size_t(item_of(items, blockIdx.y)) * g.pix_stride;
    const float* din = idepth_slab + base + g.off[l - 1];
    const float* win = weight_slab + base + g.off[l - 1];
    float* dout = idepth_slab + base + g.off[l];
    float* wout = weight_slab + base + g.off[l];
    const int R = g.rows[l], C = g.cols[l], Rin = g.rows[l - 1];
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < R * C; o += gridDim.x * blockDim.x) {
        const int x = o / R, y = o - x * R;
        const size_t p = size_t(2 * x) * Rin + 2 * y;
        const float dd[4] = {din[p], din[p + 1], din[p + Rin], din[p + Rin + 1]};
        const float ww[4] = {win[p], win[p + 1], win[p + Rin], win[p + Rin + 1]};
        float ds[4], vs[4];
        int n = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (!isnan(dd[k])) {
                ds[n] = dd[k];
                vs[n] = ww[k];
                ++n;
            }
        float d = __int_as_float(0x7fc00000), w = 0.0f;
        // explicit __fmul_rn/__fadd_rn: the reference rounds every product and sum (no FMA contraction)
        if (similar) {
            float nd = 0.0f, nv = 0.0f;
            if (n == 1) {
                d = ds[0];
                w = __fmul_rn(2.0f, vs[0]);
            } else if (n == 2) {
                nd = __fdiv_rn(__fadd_rn(__fmul_rn(ds[0], vs[1]), __fmul_rn(ds[1], vs[0])), __fadd_rn(vs[0], vs[1]));
                nv = __fdiv_rn(__fadd_rn(vs[0], vs[1]), 2.0f);
            } else if (n == 3) {
                const float v12 = __fmul_rn(vs[0], vs[1]), v13 = __fmul_rn(vs[0], vs[2]), v23 = __fmul_rn(vs[1], vs[2]);
                nd = __fdiv_rn(__fadd_rn(__fadd_rn(__fmul_rn(ds[0], v23), __fmul_rn(ds[1], v13)), __fmul_rn(ds[2], v12)),
                               __fadd_rn(__fadd_rn(v12, v13), v23));
                nv = __fdiv_rn(__fmul_rn(2.0f, __fadd_rn(__fadd_rn(vs[0], vs[1]), vs[2])), 9.0f);
            } else if (n == 4) {
                const float v123 = __fmul_rn(__fmul_rn(vs[0], vs[1]), vs[2]), v234 = __fmul_rn(__fmul_rn(vs[1], vs[2]), vs[3]);
                const float v341 = __fmul_rn(__fmul_rn(vs[2], vs[3]), vs[0]), v412 = __fmul_rn(__fmul_rn(vs[3], vs[0]), vs[1]);
                const float sum = __fadd_rn(__fadd_rn(__fadd_rn(v123, v234), v341), v412);
                nd = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(ds[0], v234), __fmul_rn(ds[1], v341)), __fmul_rn(ds[2], v412)),
                                         __fmul_rn(ds[3], v123)),
                               sum);
                nv = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(vs[0], vs[1]), vs[2]), vs[3]), 8.0f);
            }
            if (n >= 2) {
                bool ok = true;
                for (int k = 0; k < n; ++k) {
                    const float e = __fsub_rn(ds[k], nd);
                    ok = ok && (__fmul_rn(e, e) < nv);
                }
                if (ok) {
                    d = nd;
                    w = nv;
                }
            }
        } else if (n == 1) {
            d = ds[0];
            w = vs[0];
        } else if (n == 2) {
            w = __fadd_rn(vs[0], vs[1]);
            d = __fdiv_rn(__fadd_rn(__fmul_rn(ds[0], vs[0]), __fmul_rn(ds[1], vs[1])), w);
        } else if (n == 3) {
            w = __fadd_rn(__fadd_rn(vs[0], vs[1]), vs[2]);
            d = __fdiv_rn(__fadd_rn(__fadd_rn(__fmul_rn(ds[0], vs[0]), __fmul_rn(ds[1], vs[1])), __fmul_rn(ds[2], vs[2])), w);
        } else if (n == 4) {
            w = __fadd_rn(__fadd_rn(__fadd_rn(vs[0], vs[1]), vs[2]), vs[3]);
            d = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(ds[0], vs[0]), __fmul_rn(ds[1], vs[1])), __fmul_rn(ds[2], vs[2])),
                                    __fmul_rn(ds[3], vs[3])),
                          w);
        }
        dout[o] = d;
        wout[o] = w;
    }
}

// ------------------------------------------------------------------------------------------------
// Row H: `extract_z` (inverse_compositional.rs:260-279) — ordered stream compaction of the known
// idepths of every level, in memory (= column-major scan) order.  Three launches for all levels
// and all streams: per-block counts, per-(stream, level) exclusive scan, ordered scatter.  The
// scatter also gathers what the align kernel needs per candidate (template value, gradient pair).
__device__ __forceinline__ int level_of_block(const Geom& g, int blk) {
    int l = 0;
#pragma unroll
    for (int k = 1; k < kMaxLevels; ++k)
        if (k < g.L && blk >= g.blk_off[k]) l = k;
    return l;
}

__global__ void __launch_bounds__(kCompactBlock) k_compact_count(const Geom g, const float* __restrict__ idepth_slab,
                                                                 int* __restrict__ blk_count, const int* __restrict__ items) {
    const int it = item_of(items, blockIdx.y);
    const int blk = blockIdx.x;
    const int l = level_of_block(g, blk);
    const int i = (blk - g.blk_off[l]) * kCompactBlock + threadIdx.x;
    const int n = g.rows[l] * g.cols[l];
    const bool known = (i < n) && !isnan(idepth_slab[size_t(it) * g.pix_stride + g.off[l] + i]);
    const int c = __syncthreads_count(known);
    if (threadIdx.x == 0) blk_count[size_t(it) * g.blk_total + blk] = c;
}

// One CTA per (level, stream): exclusive scan of that level's block counts (in place) + total.
__global__ void __launch_bounds__(1024) k_compact_scan(const Geom g, int* __restrict__ blk_count, int* __restrict__ n_points,
                                                       uint32_t* __restrict__ pts_slab, const int* __restrict__ items) {
    __shared__ int warp_sum[32];
    __shared__ int carry;
    const int it = item_of(items, blockIdx.y);
    const int l = blockIdx.x;
    int* cnt = blk_count + size_t(it) * g.blk_total + g.blk_off[l];
    const int nb = g.blk_off[l + 1] - g.blk_off[l];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nb; b0 += 1024) {
        const int i = b0 + threadIdx.x;
        const int v = i < nb ? cnt[i] : 0;
        int s = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= d) s += t;
        }
        if (lane == 31) warp_sum[w] = s;
        __syncthreads();
        if (w == 0) {
            int ws = warp_sum[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, ws, d);
                if (lane >= d) ws += t;
            }
            warp_sum[lane] = ws;  // inclusive scan of warp totals
        }
        __syncthreads();
        const int before = carry + (w ? warp_sum[w - 1] : 0) + (s - v);
        if (i < nb) cnt[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) n_points[it * kMaxLevels + l] = carry;
    // padding up to the next whole ring stage of the align kernel: (pk, idepth, grad) = (0, NaN, 0).  A NaN inverse
    // depth fails the kernel's inside test, which sends the slot to its rare path where padding is recognised by index.
    const int total = carry, padded = (total + kPtAlign - 1) / kPtAlign * kPtAlign;
    uint32_t* lvl = pts_slab + 3 * (size_t(it) * g.pt_total + g.pt_off[l]);
    for (int i = total + threadIdx.x; i < padded; i += blockDim.x) {
        lvl[pt_word(i, 0)] = 0u;
        lvl[pt_word(i, 1)] = 0x7FC00000u;
        lvl[pt_word(i, 2)] = 0u;
    }
}

__global__ void __launch_bounds__(kCompactBlock) k_compact_scatter(const Geom g, const float* __restrict__ idepth_slab,
                                                                   const uint8_t* __restrict__ pyr_slab,
                                                                   const uint32_t* __restrict__ grad_slab,
                                                                   const int* __restrict__ blk_base, uint32_t* __restrict__ pts_slab,
                                                                   const int* __restrict__ items) {
    __shared__ int warp_cnt[32];
    const int it = item_of(items, blockIdx.y);
    const size_t base = size_t(it) * g.pix_stride;
    const int blk = blockIdx.x;
    const int l = level_of_block(g, blk);
    const int i = (blk - g.blk_off[l]) * kCompactBlock + threadIdx.x;
    const int R = g.rows[l];
    const int n = R * g.cols[l];
    const size_t src = base + g.off[l] + i;
    float d = 0.0f;
    bool known = false;
    if (i < n) {
        d = idepth_slab[src];
        known = !isnan(d);
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, known);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) warp_cnt[w] = __popc(ballot);
    __syncthreads();
    if (w == 0) {
        const int v = warp_cnt[lane];
        int s = v;
#pragma unroll
        for (int k = 1; k < 32; k <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, s, k);
            if (lane >= k) s += t;
        }
        warp_cnt[lane] = s - v;  // exclusive
    }
    __syncthreads();
    if (known) {
        const int pos = blk_base[size_t(it) * g.blk_total + blk] + warp_cnt[w] + __popc(ballot & ((1u << lane) - 1u));
        const int x = i / R, y = i - x * R;
        uint32_t* lvl = pts_slab + 3 * (size_t(it) * g.pt_total + g.pt_off[l]);
        lvl[pt_word(pos, 0)] = rec_pack_pk(x, y, uint32_t(pyr_slab[src]));
        lvl[pt_word(pos, 1)] = __float_as_uint(d);
        lvl[pt_word(pos, 2)] = rec_pack_grad(grad_slab[src]);
    }
}

}  // namespace

namespace {

// Sum over ALL candidates of a level of J J^T (21 unique entries), once per keyframe.  The align kernel then
// only accumulates J J^T for the candidates that fall OUTSIDE the frame in a pass and forms
// H = H_total - H_outside (compute_eval_data's `hessian += hes`, lm_optimizer.rs:100, over the inside set).
// J is the reference's, bit for bit (jacobian_exact); its products are exact in f64 and summed in f64 in a fixed order ->
// deterministic, and independent of how the compiler contracts the surrounding code.
// Two steps, so that a 300 k-candidate level is not one CTA's job: partial sums per CHUNK of a level (one CTA each: strided
// assignment, shuffle + shared-memory tree), then k_h_total_finish adds a level's chunks in order.
constexpr int kHPart = 22;         // doubles per chunk: 21 sums + the number of valid candidates
constexpr int kHChunkCand = 8192;  // compacted records: candidates per chunk
constexpr int kHChunkTiles = 128;  // tiled records: tiles per chunk (8 tiles per warp of the 512-thread CTA)
struct HChunks {
    int off[kMaxLevels + 1];  // first chunk of level l; off[l >= L] = chunks per stream
};

__device__ __forceinline__ int level_of_chunk(const Geom& g, const HChunks& hc, int chunk) {
    int l = 0;
#pragma unroll
    for (int k = 1; k < kMaxLevels; ++k)
        if (k < g.L && chunk >= hc.off[k]) l = k;
    return l;
}

__device__ __forceinline__ void h_accumulate(const float (&J)[6], double (&acc)[21]) {
    int t = 0;
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int b = a; b < 6; ++b, ++t) acc[t] += double(J[a]) * double(J[b]);  // exact in f64
}

// block sum (blockDim.x = 512) of the per-thread partials -> out[0..20], out[21] = number of valid candidates
__device__ __forceinline__ void h_block_reduce(double (&acc)[21], int valid, double* __restrict__ out) {
    __shared__ double part[16][21];
    __shared__ int cnt[16];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < 21; ++c) {
        double v = acc[c];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        if (lane == 0) part[w][c] = v;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) valid += __shfl_xor_sync(0xffffffffu, valid, d);
    if (lane == 0) cnt[w] = valid;
    __syncthreads();
    if (threadIdx.x < 21) {
        double s = 0.0;
        for (int ww = 0; ww < int(blockDim.x >> 5); ++ww) s += part[ww][threadIdx.x];
        out[threadIdx.x] = s;
    }
    if (threadIdx.x == 32) {
        int s = 0;
        for (int ww = 0; ww < int(blockDim.x >> 5); ++ww) s += cnt[ww];
        out[21] = double(s);
    }
}

__global__ void __launch_bounds__(512) k_h_total(const Geom g, const LevelIntrinsics li, const HChunks hc,
                                                 const uint32_t* __restrict__ pts_slab, const int* __restrict__ n_points,
                                                 double* __restrict__ part, const int* __restrict__ items) {
    const int it = item_of(items, blockIdx.y), l = level_of_chunk(g, hc, blockIdx.x);
    const int n = n_points[it * kMaxLevels + l];
    const int i0 = (int(blockIdx.x) - hc.off[l]) * kHChunkCand, i1 = min(n, i0 + kHChunkCand);
    if (i0 >= n) return;  // (the whole CTA) chunks beyond the level's candidates: k_h_total_finish does not read them
    const uint32_t* lvl = pts_slab + 3 * (size_t(it) * g.pt_total + g.pt_off[l]);
    const Intrinsics k = li.k[l];
    double acc[21];
#pragma unroll
    for (int c = 0; c < 21; ++c) acc[c] = 0.0;
    int valid = 0;
    for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
        const uint32_t p = lvl[pt_word(i, 0)], gr = lvl[pt_word(i, 2)];
        float J[6];
        jacobian_exact(rec_gx(gr), rec_gy(gr), float(rec_x(p)), float(rec_y(p)), __uint_as_float(lvl[pt_word(i, 1)]), k, J);
        h_accumulate(J, acc);
        ++valid;
    }
    h_block_reduce(acc, valid, part + (size_t(blockIdx.y) * hc.off[kMaxLevels] + blockIdx.x) * kHPart);
}

// `chunk_cand` > 0: compacted records, a level's chunks in use follow from its candidate count; 0: tiled records, every
// chunk was written and the valid slots they counted become the level's n_points.
__global__ void k_h_total_finish(const Geom g, const HChunks hc, int chunk_cand, const double* __restrict__ part,
                                 int* __restrict__ n_points, double* __restrict__ h_total, const int* __restrict__ items) {
    const int it = item_of(items, blockIdx.y), l = blockIdx.x;
    int used = hc.off[l + 1] - hc.off[l];
    if (chunk_cand) used = min(used, (n_points[it * kMaxLevels + l] + chunk_cand - 1) / chunk_cand);
    const double* p = part + (size_t(blockIdx.y) * hc.off[kMaxLevels] + hc.off[l]) * kHPart;
    if (threadIdx.x < kHPart) {
        double s = 0.0;
        for (int c = 0; c < used; ++c) s += p[size_t(c) * kHPart + threadIdx.x];
        if (threadIdx.x < 21)
            h_total[(size_t(it) * kMaxLevels + l) * kHStride + threadIdx.x] = s;
        else if (!chunk_cand)
            n_points[it * kMaxLevels + l] = int(s);
    }
}

__global__ void k_jacobians(const uint32_t* __restrict__ lvl, int n, Intrinsics k, float* __restrict__ out6) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t p = lvl[pt_word(i, 0)], gr = lvl[pt_word(i, 2)];
    float J[6];
    jacobian_at<true>(rec_gx(gr), rec_gy(gr), float(rec_x(p)), float(rec_y(p)),
                      __uint_as_float(lvl[pt_word(i, 1)]), k, J);
#pragma unroll
    for (int a = 0; a < 6; ++a) out6[size_t(i) * 6 + a] = J[a];
}

__global__ void k_se3_exp(const float* __restrict__ xi6, Pose* __restrict__ out) {
    float xi[6];
    for (int a = 0; a < 6; ++a) xi[a] = xi6[a];
    *out = se3_exp(xi);
}

// op: 0 se3::exp (6 -> 7), 1 se3::log (7 -> 6), 2 so3::exp (3 -> 4), 3 so3::log (4 -> 3); one thread, f32 like the reference
__global__ void k_lie(int op, const float* __restrict__ in, float* __restrict__ out) {
    if (op == 1) {
        const Pose p{{in[0], in[1], in[2]}, {in[3], in[4], in[5], in[6]}};
        float xi[6];
        se3_log(p, xi);
        for (int a = 0; a < 6; ++a) out[a] = xi[a];
    } else if (op == 2) {
        const Quat q = so3_exp(Vec3{in[0], in[1], in[2]});
        out[0] = q.i; out[1] = q.j; out[2] = q.k; out[3] = q.w;
    } else if (op == 3) {
        const Vec3 w = so3_log(Quat{in[0], in[1], in[2], in[3]});
        out[0] = w.x; out[1] = w.y; out[2] = w.z;
    }
}

// ------------------------------------------------------------------------------------------------
// Tiled dense records (vors_device.cuh): one warp writes one tile = one 3840-byte ring stage of the align
// kernel, slot (j, lane) = pixel (x = 12 tx + j, y = 32 ty + lane) of the tile's level.  No compaction: a pixel outside the
// image or without a known inverse depth gets a NaN inverse depth (extract_z, inverse_compositional.rs:260-279, keeps
// exactly the pixels whose inverse depth is known; here they keep their place and the others are marked).
__device__ __forceinline__ int level_of_tile(const Geom& g, int tile) {
    int l = 0;
#pragma unroll
    for (int k = 1; k < kMaxLevels; ++k)
        if (k < g.L && tile >= g.tile_off[k]) l = k;
    return l;
}

constexpr int kTileCtaWarps = 8;  // tiles per CTA of k_tile_records: a warp per tile
__global__ void __launch_bounds__(kTileCtaWarps * 32) k_tile_records(const Geom g, int scharr, const float* __restrict__ idepth_slab,
                                                                     const uint8_t* __restrict__ pyr_slab,
                                                                     const uint32_t* __restrict__ grad_slab,
                                                                     uint32_t* __restrict__ pts_slab, const int* __restrict__ items) {
    // a warp assembles its tile here (lane = row, the twelve columns unrolled: the tile's index arithmetic is paid once per
    // lane and all of a lane's loads are in flight together); the tile then leaves with 16-byte stores
    __shared__ __align__(16) uint32_t s_tile[kTileCtaWarps][kTileWords];
    const int it = item_of(items, blockIdx.y);
    const size_t base = size_t(it) * g.pix_stride;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x * kTileCtaWarps + warp;
    if (tile >= g.tile_total) return;  // (a whole warp; nothing below synchronises the block)
    const int l = level_of_tile(g, tile);
    const int t = tile - g.tile_off[l], tx = t / g.tiles_y[l], ty = t - tx * g.tiles_y[l];
    const int R = g.rows[l], C = g.cols[l];
    const int y = kTileRows * ty + lane;
    uint32_t* s = s_tile[warp];
#pragma unroll
    for (int j = 0; j < kTileCols; ++j) {
        const int x = kTileCols * tx + j;
        // rows below the image: zero inverse depth (the align kernel's dead lanes); columns beyond it and pixels without depth: NaN
        float rho = y < R ? __int_as_float(0x7fc00000) : 0.0f;
        uint32_t gr = 0u;
        unsigned short tm = 0;
        if (x < C && y < R) {
            const size_t src = base + g.off[l] + size_t(x) * R + y;
            const float d = idepth_slab[src];
            // (no gradient slab: the Tracker's gradient recipe evaluated here, same function as k_gradients)
            const uint32_t pair = grad_slab ? grad_slab[src] : grad_pair_at(g, pyr_slab + base, scharr, l, x, y);
            const uint8_t t8 = pyr_slab[src];
            if (!isnan(d)) {
                rho = d;
                gr = rec_pack_grad(pair);
                tm = __half_as_ushort(__float2half_rn(float(t8)));  // 0..255: exact in f16
            }
        }
        s[tile_rho_word(j, lane)] = __float_as_uint(rho);
        s[tile_grad_word(j, lane)] = gr;
        reinterpret_cast<unsigned short*>(s)[tile_tmpl_half(j, lane)] = tm;
    }
    __syncwarp();
    uint4* st = reinterpret_cast<uint4*>(pts_slab + (size_t(it) * g.tile_total + tile) * kTileWords);
    for (int i = lane; i < kTileWords / 4; i += 32) st[i] = reinterpret_cast<const uint4*>(s)[i];
}
static_assert(kTileWords % 4 == 0, "16-byte stores move a tile");

// Row K for tiled records: partial sums of J J^T over the valid slots of a chunk of kHChunkTiles tiles and their count.
static_assert(kTileHalfCols % 2 == 0 && (kTileRows * kTileHalfCols) % 2 == 0 && kTileSlots % 2 == 0, "8-byte loads of a lane's half row");
__global__ void __launch_bounds__(512) k_h_total_tiled(const Geom g, const LevelIntrinsics li, const HChunks hc,
                                                       const uint32_t* __restrict__ pts_slab, double* __restrict__ part,
                                                       const int* __restrict__ items) {
    const int it = item_of(items, blockIdx.y), l = level_of_chunk(g, hc, blockIdx.x);
    const int tiles_y = g.tiles_y[l], tiles_l = tiles_y * g.tiles_x[l];
    const int t0 = (int(blockIdx.x) - hc.off[l]) * kHChunkTiles, t1 = min(tiles_l, t0 + kHChunkTiles);
    const uint32_t* lvl = pts_slab + (size_t(it) * g.tile_total + g.tile_off[l]) * kTileWords;
    const Intrinsics k = li.k[l];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc[21];
#pragma unroll
    for (int c = 0; c < 21; ++c) acc[c] = 0.0;
    int valid = 0;
    // a warp takes whole tiles, lane = row, the twelve columns unrolled: no per-slot index arithmetic, and the six inverse
    // depths / gradient pairs of a lane's half row arrive with three 8-byte loads each (the align kernel's access pattern)
    for (int st = t0 + warp; st < t1; st += int(blockDim.x >> 5)) {
        const int tx = st / tiles_y, ty = st - tx * tiles_y;
        const uint32_t* w = lvl + size_t(st) * kTileWords;
        const float y = float(kTileRows * ty + lane);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            uint2 r2[kTileHalfCols / 2], g2[kTileHalfCols / 2];
#pragma unroll
            for (int q = 0; q < kTileHalfCols / 2; ++q) {
                r2[q] = *reinterpret_cast<const uint2*>(w + tile_rho_word(kTileHalfCols * h + 2 * q, lane));
                g2[q] = *reinterpret_cast<const uint2*>(w + tile_grad_word(kTileHalfCols * h + 2 * q, lane));
            }
#pragma unroll
            for (int c = 0; c < kTileHalfCols; ++c) {
                const float rho = __uint_as_float((c & 1) ? r2[c >> 1].y : r2[c >> 1].x);
                const uint32_t gr = (c & 1) ? g2[c >> 1].y : g2[c >> 1].x;
                if (isnan(rho) || rho == 0.0f) continue;
                ++valid;
                float J[6];
                jacobian_exact(rec_gx(gr), rec_gy(gr), float(kTileCols * tx + kTileHalfCols * h + c), y, rho, k, J);
                h_accumulate(J, acc);
            }
        }
    }
    h_block_reduce(acc, valid, part + (size_t(blockIdx.y) * hc.off[kMaxLevels] + blockIdx.x) * kHPart);
}

// Frame pyramids -> atlas page (vors_device.cuh): every level of every listed stream, through surface stores.
// Texture x = image row (the fast axis of the column-major pyramid).  A thread moves a QUAD = four consecutive rows of one
// column: one 4-byte load and one 4-byte surface store (8-byte for f16 texels) where both are aligned - cells start at
// multiples of four texels - else texel by texel (last rows of a level whose height is not a multiple of four, odd slabs).
struct AtlasQuads {
    int off[kMaxLevels + 1];  // first quad of level l; off[l >= L] = quads per stream
};

__device__ __forceinline__ void atlas_store1(const AtlasPages& pages, cudaSurfaceObject_t surf, uint8_t v, int tx, int ty) {
    if (pages.f16)
        surf2Dwrite<unsigned short>(__half_as_ushort(__float2half_rn(float(v))), surf, tx * 2, ty);
    else
        surf2Dwrite<unsigned char>(v, surf, tx, ty);
}

__global__ void k_atlas_fill(const Geom g, const AtlasQuads aq, const uint8_t* __restrict__ pyr_slab, const AtlasPages pages,
                             const int* __restrict__ items, int first) {
    // stream `it`; its slab sits at index items[j] of pyr_slab, or at index j when pyr_slab already points at stream `first`
    const int it = items ? items[blockIdx.y] : first + int(blockIdx.y);
    const uint8_t* pyr = pyr_slab + size_t(items ? items[blockIdx.y] : int(blockIdx.y)) * g.pix_stride;
    const int page = it / g.per_page, cell = it - page * g.per_page;
    const int ox = (cell % g.per_row) * g.cell_w, oy = (cell / g.per_row) * g.cell_h;
    const cudaSurfaceObject_t surf = pages.surf[page];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < aq.off[kMaxLevels]; i += gridDim.x * blockDim.x) {
        // (the level's constants by unrolled selects: dynamic indexing would copy the parameter structs to local memory)
        int q0 = 0, R = g.rows[0], p0 = 0, ly = 0;
#pragma unroll
        for (int k = 1; k < kMaxLevels; ++k)
            if (k < g.L && i >= aq.off[k]) {
                q0 = aq.off[k];
                R = g.rows[k];
                p0 = g.off[k];
                ly = g.lvl_y[k];
            }
        const int o = i - q0, Q = (R + 3) >> 2;
        const int x = o / Q, y = (o - x * Q) * 4;
        const uint8_t* src = pyr + p0 + size_t(x) * R + y;
        const int tx = ox + y, ty = oy + ly + x;
        const int nv = min(4, R - y);
        if (nv == 4 && ((reinterpret_cast<uintptr_t>(src) | uintptr_t(tx)) & 3u) == 0) {
            const uchar4 v = *reinterpret_cast<const uchar4*>(src);
            if (pages.f16) {
                const ushort4 h = make_ushort4(__half_as_ushort(__float2half_rn(float(v.x))), __half_as_ushort(__float2half_rn(float(v.y))),
                                               __half_as_ushort(__float2half_rn(float(v.z))), __half_as_ushort(__float2half_rn(float(v.w))));
                surf2Dwrite<ushort4>(h, surf, tx * 2, ty);
            } else {
                surf2Dwrite<uchar4>(v, surf, tx, ty);
            }
        } else {
            for (int k = 0; k < nv; ++k) atlas_store1(pages, surf, src[k], tx + k, ty);
        }
    }
}

inline int grid_for(int n, int block, int cap = 148 * 16) { return max(1, min((n + block - 1) / block, cap)); }
}  // namespace

// ---- launchers -----------------------------------------------------------------------------------
void launch_transpose_u8(Launcher& L, const uint8_t* in, uint8_t* out_slab, size_t out_stride, const int* items, int m, int rows,
                         int cols) {
    dim3 grid((cols + 31) / 32, (rows + 31) / 32, m), block(32, 8);
    k_transpose<uint8_t><<<grid, block, 0, L.stream>>>(in, out_slab, out_stride, items, rows, cols);
    ++L.launches;
}
void launch_transpose_u16(Launcher& L, const uint16_t* in, uint16_t* out_slab, size_t out_stride, const int* items, int m, int rows,
                          int cols) {
    dim3 grid((cols + 31) / 32, (rows + 31) / 32, m), block(32, 8);
    k_transpose<uint16_t><<<grid, block, 0, L.stream>>>(in, out_slab, out_stride, items, rows, cols);
    ++L.launches;
}
void launch_copy_items_u16(Launcher& L, const uint16_t* in, uint16_t* out_slab, size_t out_stride, const int* items, int m,
                           size_t count) {
    dim3 grid(grid_for(int(count), 256, 148 * 4), m);
    k_copy_items_u16<<<grid, 256, 0, L.stream>>>(in, out_slab, out_stride, items, count);
    ++L.launches;
}
void launch_pyramid(Launcher& L, const Geom& g, uint8_t* pyr_slab, const int* items, int m) {
    // fused: up to 6 halvings per launch (a 64x64 tile collapses to 1 pixel); deeper pyramids chain a second launch
    for (int l0 = 0; l0 + 1 < g.L; l0 += 6) {
        const int count = std::min(6, g.L - 1 - l0);
        dim3 grid((g.rows[l0] + kPyrTile - 1) / kPyrTile, (g.cols[l0] + kPyrTile - 1) / kPyrTile, m);
        k_pyramid_fused<<<grid, 256, 0, L.stream>>>(g, l0, count, pyr_slab, items);
        ++L.launches;
    }
}
// one level at a time (kept for comparison / debugging)
void launch_pyramid_levelwise(Launcher& L, const Geom& g, uint8_t* pyr_slab, const int* items, int m) {
    for (int l = 1; l < g.L; ++l) {
        dim3 grid(grid_for(g.rows[l] * g.cols[l], 256), m);
        k_halve_mean<<<grid, 256, 0, L.stream>>>(g, l, pyr_slab, items);
        ++L.launches;
    }
}
void launch_gradients(Launcher& L, const Geom& g, int scharr, const uint8_t* pyr_slab, uint32_t* grad_slab, uint16_t* g2_slab, const int* items,
                      int m) {
    dim3 grid(grid_for(g.pix_total, 256), m);
    k_gradients<<<grid, 256, 0, L.stream>>>(g, scharr, pyr_slab, grad_slab, g2_slab, items);
    ++L.launches;
}
void launch_c2f(Launcher& L, const Geom& g, uint16_t thresh, const uint16_t* g2_slab, uint8_t* mask_slab, const int* items, int m) {
    for (int l = g.L - 2; l >= 0; --l) {
        const int parents = ((g.rows[l] + 1) / 2) * ((g.cols[l] + 1) / 2);
        dim3 grid(grid_for(parents, 256), m);
        k_c2f_level<<<grid, 256, 0, L.stream>>>(g, l, thresh, g2_slab, mask_slab, items);
        ++L.launches;
    }
}
void launch_idepth(Launcher& L, const Geom& g, const uint16_t* depth_slab, size_t depth_stride, const uint8_t* mask_slab, int dense,
                   float scale, float variance, int similar, float* idepth_slab, float* weight_slab, const int* items, int m) {
    {
        dim3 grid(grid_for(g.rows[0] * g.cols[0], 256), m);
        k_idepth0<<<grid, 256, 0, L.stream>>>(g, depth_slab, depth_stride, mask_slab, dense, scale, variance, idepth_slab,
                                              weight_slab, items);
        ++L.launches;
    }
    for (int l = 1; l < g.L; ++l) {
        dim3 grid(grid_for(g.rows[l] * g.cols[l], 256), m);
        k_idepth_halve<<<grid, 256, 0, L.stream>>>(g, l, similar, idepth_slab, weight_slab, items);
        ++L.launches;
    }
}
void launch_compact(Launcher& L, const Geom& g, const float* idepth_slab, const uint8_t* pyr_slab, const uint32_t* grad_slab,
                    int* blk_count, int* n_points, uint32_t* pts_slab, const int* items, int m) {
    dim3 gridb(g.blk_total, m);
    k_compact_count<<<gridb, kCompactBlock, 0, L.stream>>>(g, idepth_slab, blk_count, items);
    dim3 grids(g.L, m);
    k_compact_scan<<<grids, 1024, 0, L.stream>>>(g, blk_count, n_points, pts_slab, items);
    k_compact_scatter<<<gridb, kCompactBlock, 0, L.stream>>>(g, idepth_slab, pyr_slab, grad_slab, blk_count, pts_slab, items);
    L.launches += 3;
}
namespace {
HChunks h_chunks(const Geom& g, bool tiled) {
    HChunks hc;
    int off = 0;
    for (int l = 0; l <= kMaxLevels; ++l) {
        hc.off[l] = off;
        if (l < g.L)
            off += tiled ? (g.tiles_y[l] * g.tiles_x[l] + kHChunkTiles - 1) / kHChunkTiles
                         : (g.rows[l] * g.cols[l] + kHChunkCand - 1) / kHChunkCand;
    }
    return hc;
}
LevelIntrinsics level_intrinsics(const Geom& g, const Intrinsics* intr) {
    LevelIntrinsics li;
    for (int l = 0; l < kMaxLevels; ++l) li.k[l] = intr[l < g.L ? l : g.L - 1];
    return li;
}
}  // namespace
size_t h_total_scratch_doubles(const Geom& g, bool tiled) { return size_t(h_chunks(g, tiled).off[kMaxLevels]) * kHPart; }

void launch_h_total(Launcher& L, const Geom& g, const Intrinsics* intr, const uint32_t* pts_slab, int* n_points, double* h_part,
                    double* h_total, const int* items, int m) {
    const HChunks hc = h_chunks(g, false);
    dim3 grid(hc.off[kMaxLevels], m), gridf(g.L, m);
    k_h_total<<<grid, 512, 0, L.stream>>>(g, level_intrinsics(g, intr), hc, pts_slab, n_points, h_part, items);
    k_h_total_finish<<<gridf, 32, 0, L.stream>>>(g, hc, kHChunkCand, h_part, n_points, h_total, items);
    L.launches += 2;
}
void launch_tile_records(Launcher& L, const Geom& g, const Intrinsics* intr, int scharr, const float* idepth_slab, const uint8_t* pyr_slab,
                         const uint32_t* grad_slab, int* n_points, uint32_t* pts_slab, double* h_part, double* h_total, const int* items,
                         int m) {
    dim3 grid((g.tile_total + kTileCtaWarps - 1) / kTileCtaWarps, m);
    k_tile_records<<<grid, kTileCtaWarps * 32, 0, L.stream>>>(g, scharr, idepth_slab, pyr_slab, grad_slab, pts_slab, items);
    const HChunks hc = h_chunks(g, true);
    dim3 gridh(hc.off[kMaxLevels], m), gridf(g.L, m);
    k_h_total_tiled<<<gridh, 512, 0, L.stream>>>(g, level_intrinsics(g, intr), hc, pts_slab, h_part, items);
    k_h_total_finish<<<gridf, 32, 0, L.stream>>>(g, hc, 0, h_part, n_points, h_total, items);
    L.launches += 3;
}
void launch_atlas_fill(Launcher& L, const Geom& g, const uint8_t* pyr_slab, const AtlasPages& pages, const int* items, int m, int first) {
    AtlasQuads aq;
    int off = 0;
    for (int l = 0; l <= kMaxLevels; ++l) {
        aq.off[l] = off;
        if (l < g.L) off += ((g.rows[l] + 3) / 4) * g.cols[l];
    }
    dim3 grid(grid_for(off, 256, std::max(64, 148 * 8 / std::max(m, 1))), m);  // a single stream still fills the device
    k_atlas_fill<<<grid, 256, 0, L.stream>>>(g, aq, pyr_slab, pages, items, first);
    ++L.launches;
}
void launch_jacobians(Launcher& L, const uint32_t* pts_level, int n, Intrinsics k, float* out6) {
    if (n <= 0) return;
    k_jacobians<<<(n + 255) / 256, 256, 0, L.stream>>>(pts_level, n, k, out6);
    ++L.launches;
}
void launch_lie(Launcher& L, int op, const float* in, float* out) {
    k_lie<<<1, 1, 0, L.stream>>>(op, in, out);
    ++L.launches;
}
void launch_se3_exp(Launcher& L, const float* xi6, Pose* out) {
    k_se3_exp<<<1, 1, 0, L.stream>>>(xi6, out);
    ++L.launches;
}

}  // namespace vors
