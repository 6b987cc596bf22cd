// dso_kernels.cu — SURVEY.md §8a rows R and S on the device, sm_100a.
//
// Row R: candidates::dso::select (src/core/candidates/dso.rs:98-325) — region medians, smoothed thresholds,
// per-block arg-max pyramid, level-wise picking, and the seeded replacement of the reference's
// non-reproducible `thread_rng` thinning branch (dso.rs:140-143).  The recursion over block sizes (dso.rs:117-139)
// needs the candidate count, so it is driven from the host (one small readback per iteration).
// Row S: the example gradient-norm recipe (gradient.rs:49-65 squared_norm_direct, :102-111 bloc_squared_norm,
// multires.rs:96-106) — numerically different from the Tracker recipe (no intermediate truncation).
#include <algorithm>
#include <cmath>

#include "vors_device.cuh"

namespace vors {

namespace {

// gradient.rs:49-65: ((gx^2 + gy^2) / 4) as u16 on the un-truncated differences; 1-px border 0.
// `as_magnitude`: examples/candidates_dso.rs:42 feeds DSO with sqrt(g2) as u16 (f32 sqrt, truncating cast).
__global__ void k_sqnorm_direct(const uint8_t* __restrict__ img, int R, int C, int as_magnitude, uint16_t* __restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < R * C; i += gridDim.x * blockDim.x) {
        const int x = i / R, y = i - x * R;
        uint16_t v = 0;
        if (x > 0 && x < C - 1 && y > 0 && y < R - 1) {
            const uint8_t* p = img + size_t(x) * R + y;
            const int gx = int(p[R]) - int(p[-R]), gy = int(p[1]) - int(p[-1]);
            v = uint16_t((gx * gx + gy * gy) / 4);
        }
        out[i] = as_magnitude ? uint16_t(__fsqrt_rn(float(v))) : v;
    }
}

// gradient.rs:102-111 `bloc_squared_norm` through multires::halve: level l (>= 1) from the level l-1 image.
__global__ void k_bloc_sqnorm(const uint8_t* __restrict__ fine, int Rin, int R, int C, uint16_t* __restrict__ out) {
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < R * C; o += gridDim.x * blockDim.x) {
        const int x = o / R, y = o - x * R;
        const uint8_t* p = fine + size_t(2 * x) * Rin + 2 * y;
        const int a = p[0], b = p[1], c = p[Rin], d = p[Rin + 1];
        const int dx = c + d - a - b, dy = b - a + d - c;
        out[o] = uint16_t((dx * dx + dy * dy) / 4);
    }
}

// dso.rs:307-325 `region_median_gradients`: upper median sorted[len/2] of each size x size region (smaller at the
// right / bottom), by a two-pass radix select over the u16 values.  One CTA per region.
__global__ void __launch_bounds__(256) k_region_median(const uint16_t* __restrict__ g, int rows, int cols, int size, int nrr,
                                                       uint16_t* __restrict__ med) {
    __shared__ int hist[256];
    __shared__ int sel, krem;
    const int ri = blockIdx.x % nrr, rj = blockIdx.x / nrr;
    const int h = min(size, rows - ri * size), w = min(size, cols - rj * size);
    const int n = h * w;
    const uint16_t* base = g + size_t(rj * size) * rows + ri * size;
    for (int pass = 0; pass < 2; ++pass) {
        hist[threadIdx.x] = 0;
        __syncthreads();
        const int hi = sel;  // valid in pass 1
        for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
            const int c = idx / h, r = idx - c * h;
            const unsigned v = base[size_t(c) * rows + r];
            if (pass == 0)
                atomicAdd(&hist[v >> 8], 1);
            else if (int(v >> 8) == hi)
                atomicAdd(&hist[v & 255u], 1);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int k = pass == 0 ? n / 2 : krem, cum = 0, b = 0;
            for (; b < 256; ++b) {
                if (cum + hist[b] > k) break;
                cum += hist[b];
            }
            if (pass == 0) {
                sel = b;
                krem = k - cum;
            } else {
                med[blockIdx.x] = uint16_t((hi << 8) | b);
            }
        }
        __syncthreads();
    }
}

// dso.rs:284-303 `region_thresholds`: a * (mean3x3(median) + b)^2; the 3x3 sum is accumulated in u16 (wraps like
// release-mode Rust), the result is cast back to u16 (truncation); out-of-range panics in the reference -> *err = 1.
__global__ void k_region_thresholds(const uint16_t* __restrict__ med, int nrr, int nrc, float a, uint16_t b,
                                    uint16_t* __restrict__ th, int* __restrict__ err) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nrr * nrc) return;
    const int j = t / nrr, i = t - j * nrr;
    const int si = max(0, i - 1), sj = max(0, j - 1), ei = min(nrr, i + 2), ej = min(nrc, j + 2);
    uint16_t sum = 0;
    int nb = 0;
    for (int jj = sj; jj < ej; ++jj)
        for (int ii = si; ii < ei; ++ii) {
            sum = uint16_t(sum + med[jj * nrr + ii]);
            ++nb;
        }
    const float tmp = __fadd_rn(__fdiv_rn(float(sum), float(nb)), float(b));
    const float val = __fmul_rn(__fmul_rn(a, tmp), tmp);
    if (!(val > -1.0f && val < 65536.0f)) {
        *err = 1;
        th[t] = 0;
    } else {
        th[t] = uint16_t(val);
    }
}

struct BlockMax {
    uint16_t g, i, j, pad;
};

// dso.rs:193-222 `init_max_gradients`: per block the FIRST strict maximum in column-major scan.
__global__ void k_block_max0(const uint16_t* __restrict__ g, int rows, int cols, int bs, int nbr, int nbc, BlockMax* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nbr * nbc) return;
    const int bj = t / nbr, bi = t - bj * nbr;
    const int si = bi * bs, sj = bj * bs, ei = min(si + bs, rows), ej = min(sj + bs, cols);
    BlockMax m{g[size_t(sj) * rows + si], uint16_t(si), uint16_t(sj), 0};
    for (int j = sj; j < ej; ++j)
        for (int i = si; i < ei; ++i) {
            const uint16_t v = g[size_t(j) * rows + i];
            if (v > m.g) m = BlockMax{v, uint16_t(i), uint16_t(j), 0};
        }
    out[t] = m;
}

// dso.rs:225-240 `max_of_four_gradients` through multires::halve: g_max(g1, g_max(g2, g_max(g3, g4))), ties keep the left.
__global__ void k_block_max_halve(const BlockMax* __restrict__ in, int hin, int h, int w, BlockMax* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= h * w) return;
    const int j = t / h, i = t - j * h;
    const BlockMax a = in[size_t(2 * j) * hin + 2 * i], b = in[size_t(2 * j) * hin + 2 * i + 1];
    const BlockMax c = in[size_t(2 * j + 1) * hin + 2 * i], d = in[size_t(2 * j + 1) * hin + 2 * i + 1];
    BlockMax m = c.g < d.g ? d : c;
    m = b.g < m.g ? m : b;
    m = a.g < m.g ? m : a;
    out[t] = m;
}

// dso.rs:246-276 `pick_level_block_candidates`, one thread per parent of the next mask level: a block whose max
// gradient reaches coef * region threshold is picked (level tag written at the pixel); the parent stays eligible
// only if all four children were eligible and none was picked.  Blocks beyond the even-cropped range are skipped.
__global__ void k_pick_level(const BlockMax* __restrict__ mg, const uint8_t* __restrict__ mask, int h, int w, float coef,
                             uint8_t level, const uint16_t* __restrict__ th, int nrr, int region, int rows,
                             uint8_t* __restrict__ picked, uint8_t* __restrict__ mask_next, int* __restrict__ count) {
    const int ph = h / 2, pw = w / 2;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ph * pw) return;
    const int pj = t / ph, pi = t - pj * ph;
    bool parent_ok = true;
    int n = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = 2 * pi + (k & 1), j = 2 * pj + (k >> 1);
        const size_t idx = size_t(j) * h + i;
        if (mask == nullptr || mask[idx]) {
            const BlockMax m = mg[idx];
            const uint16_t thr = th[(m.j / region) * nrr + (m.i / region)];
            if (float(m.g) >= __fmul_rn(coef, float(thr))) {
                parent_ok = false;
                picked[size_t(m.j) * rows + m.i] = level;
                ++n;
            }
        } else {
            parent_ok = false;
        }
    }
    if (mask_next) mask_next[t] = parent_ok ? 1 : 0;
    if (n) atomicAdd(count, n);
}

__device__ __forceinline__ unsigned long long splitmix64_at(unsigned long long seed, unsigned long long index) {
    unsigned long long z = seed + (index + 1ull) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// dso.rs:138-145: picked -> bool mask; in the thinning branch (:140-143) the r-th picked pixel in column-major order
// keeps its place iff the r-th draw of the seeded generator is <= (255 / ratio) as u8.  One CTA walks the image in
// order with a running rank.
__global__ void __launch_bounds__(1024) k_dso_to_mask(const uint8_t* __restrict__ picked, int n, int thin, unsigned lim,
                                                      unsigned long long seed, uint8_t* __restrict__ mask) {
    __shared__ int warp_cnt[32];
    __shared__ int carry;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < n; b0 += 1024) {
        const int i = b0 + threadIdx.x;
        const bool p = i < n && picked[i] > 0;
        const unsigned ballot = __ballot_sync(0xffffffffu, p);
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
            warp_cnt[lane] = s - v;
        }
        __syncthreads();
        const int base = carry;
        if (i < n) {
            bool keep = p;
            if (p && thin) {
                const unsigned long long rank = (unsigned long long)(base + warp_cnt[w] + __popc(ballot & ((1u << lane) - 1u)));
                keep = (unsigned(splitmix64_at(seed, rank)) & 0xFFu) <= lim;
            }
            mask[i] = keep ? 1 : 0;
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry = base + warp_cnt[31] + __popc(ballot);
        __syncthreads();
    }
}

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

}  // namespace

void launch_sqnorm_direct(Launcher& L, const uint8_t* img, int rows, int cols, int as_magnitude, uint16_t* out) {
    k_sqnorm_direct<<<std::max(1, std::min(cdiv(rows * cols, 256), 148 * 8)), 256, 0, L.stream>>>(img, rows, cols, as_magnitude, out);
    ++L.launches;
}
void launch_bloc_sqnorm(Launcher& L, const uint8_t* fine, int rows_in, int rows, int cols, uint16_t* out) {
    k_bloc_sqnorm<<<std::max(1, std::min(cdiv(rows * cols, 256), 148 * 8)), 256, 0, L.stream>>>(fine, rows_in, rows, cols, out);
    ++L.launches;
}

// Host driver of dso::select with the DEFAULT_* configs (dso.rs:72-90).  `ws` must hold dso_workspace_bytes().
size_t dso_workspace_bytes(int rows, int cols) {
    const size_t px = size_t(rows) * cols;
    // picked + mask_a + mask_b (u8, block grids are <= px) + 3 block-max levels (8 B, <= px each at base size 1) + regions
    return px * 3 + px * 8 * 2 + 4096 * 2 * 2 + 1024;
}

int dso_select_device(Launcher& L, const uint16_t* d_grad, int rows, int cols, int nb_target, int nb_iterations_left,
                      unsigned long long seed, uint8_t* d_mask_out, uint8_t* ws, int* h_pinned_scratch, int* used_random,
                      int* nb_candidates_out) {
    const int region = 32;
    const float coef_a = 1.0f;
    const uint16_t coef_b = 3;
    const int nb_levels = 3;
    const float threshold_factor = 0.5f, low = 0.8f, high = 4.0f, random_thresh = 1.1f;
    const size_t px = size_t(rows) * cols;
    const int nrr = cdiv(rows, region), nrc = cdiv(cols, region);
    if (size_t(nrr) * nrc > 4096) return VORS_E_INVALID;

    uint8_t* picked = ws;
    uint8_t* mask_a = ws + px;
    uint8_t* mask_b = ws + 2 * px;
    BlockMax* bm0 = reinterpret_cast<BlockMax*>(ws + 3 * px);  // px is a multiple of ... ensure 8-byte alignment below
    bm0 = reinterpret_cast<BlockMax*>((reinterpret_cast<uintptr_t>(bm0) + 7) & ~uintptr_t(7));
    BlockMax* bm1 = bm0 + px;
    uint16_t* med = reinterpret_cast<uint16_t*>(bm1 + px / 2 + 8);
    uint16_t* th = med + 4096;
    int* d_flags = reinterpret_cast<int*>((reinterpret_cast<uintptr_t>(th + 4096) + 7) & ~uintptr_t(7));  // [0] count, [1] err

    k_region_median<<<nrr * nrc, 256, 0, L.stream>>>(d_grad, rows, cols, region, nrr, med);
    cudaMemsetAsync(d_flags, 0, 8, L.stream);
    k_region_thresholds<<<cdiv(nrr * nrc, 128), 128, 0, L.stream>>>(med, nrr, nrc, coef_a, coef_b, th, d_flags + 1);
    L.launches += 2;
    if (used_random) *used_random = 0;

    int base_size = 4, left = nb_iterations_left;
    for (;;) {
        // pick_all_block_candidates (dso.rs:156-190)
        int h[4], w[4];
        h[0] = cdiv(rows, base_size);
        w[0] = cdiv(cols, base_size);
        int n_lv = 1;
        while (n_lv < nb_levels && h[n_lv - 1] / 2 > 0 && w[n_lv - 1] / 2 > 0) {
            h[n_lv] = h[n_lv - 1] / 2;
            w[n_lv] = w[n_lv - 1] / 2;
            ++n_lv;
        }
        BlockMax* lv[3] = {bm0, bm1, bm1 + size_t(h[0] / 2 + 1) * (w[0] / 2 + 1)};
        k_block_max0<<<cdiv(h[0] * w[0], 128), 128, 0, L.stream>>>(d_grad, rows, cols, base_size, h[0], w[0], lv[0]);
        ++L.launches;
        for (int l = 1; l < n_lv; ++l) {
            k_block_max_halve<<<cdiv(h[l] * w[l], 128), 128, 0, L.stream>>>(lv[l - 1], h[l - 1], h[l], w[l], lv[l]);
            ++L.launches;
        }
        cudaMemsetAsync(picked, 0, px, L.stream);
        cudaMemsetAsync(d_flags, 0, 4, L.stream);
        float coef = 1.0f;
        const uint8_t* mask = nullptr;  // level 0: all blocks eligible
        uint8_t* next = mask_a;
        for (int l = 0; l < n_lv; ++l) {
            const int parents = (h[l] / 2) * (w[l] / 2);
            if (parents > 0) {
                k_pick_level<<<cdiv(parents, 128), 128, 0, L.stream>>>(lv[l], mask, h[l], w[l], coef, uint8_t(l + 1), th, nrr, region, rows,
                                                                     picked, next, d_flags);
                ++L.launches;
            }
            mask = next;
            next = (next == mask_a) ? mask_b : mask_a;
            coef *= threshold_factor;
        }
        cudaMemcpyAsync(h_pinned_scratch, d_flags, 8, cudaMemcpyDeviceToHost, L.stream);
        if (cudaStreamSynchronize(L.stream) != cudaSuccess) return VORS_E_CUDA;
        if (h_pinned_scratch[1]) return VORS_E_INVALID;  // the reference's `expect("woops")` panic
        const int nb_candidates = h_pinned_scratch[0];
        if (nb_candidates_out) *nb_candidates_out = nb_candidates;
        // dso.rs:115-150
        const float ratio = float(nb_candidates) / float(nb_target);
        const float target_size_f = std::sqrt(ratio) * (float(base_size) + 1.0f) - 1.0f;
        const int target_size = std::max(1, int(std::round(target_size_f)));
        int thin = 0;
        if (ratio < low || ratio > high) {
            if (target_size != base_size && left > 0) {
                base_size = target_size;
                --left;
                continue;
            }
        } else if (ratio > random_thresh) {
            thin = 1;
            if (used_random) *used_random = 1;
        }
        const unsigned lim = thin ? unsigned(uint8_t(255.0f / ratio)) : 0u;
        k_dso_to_mask<<<1, 1024, 0, L.stream>>>(picked, int(px), thin, lim, seed, d_mask_out);
        ++L.launches;
        return VORS_OK;
    }
}

}  // namespace vors
