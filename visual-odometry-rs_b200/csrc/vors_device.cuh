// vors_device.cuh — device-side data structures shared by the kernels and the host engine.
//
// HBM layout (everything column-major like nalgebra::DMatrix: pixel (row y, col x) at x*rows + y):
//   * an image pyramid is one slab of `pix_total` elements, level l at pixel offset off[l];
//   * a keyframe's candidate points are three 4-byte streams per level in the reference's scan order
//     (extract_z, inverse_compositional.rs:260-279), chunk-blocked (64 candidates = one 768-byte record
//     pk[64] | idepth[64] | grad[64]) at offsets pt_off[l] aligned to kPtAlign candidates with capacity rows_l*cols_l,
//     so the align kernel stages whole ring stages with 16-byte aligned bulk copies:
//     pk = x | y<<12 | template<<24,  idepth (f32),  grad = half2(gx, gy) (integers |g| <= 255: exact in f16).
//     Padding up to the next kPtAlign multiple is (0, NaN, 0): a NaN inverse depth can never pass the align
//     kernel's inside test.  12 B per candidate; the align kernel recomputes the Jacobian and J J^T in registers.
//   * n streams (trackers) of a batch own consecutive slabs: base + stream * pix_stride (pt_total for candidates);
//     pix_stride = pix_total + a zero page of rows_0 + 2 bytes (rounded up to 16) that is never written in the
//     frame-pyramid slab: the align kernel points candidates that fall outside the frame at it (all four texels 0).
//   * DENSE keyframes with zero skew and plain L2 (the benchmarked configuration) use TILED records with IMPLICIT coordinates
//     instead: a level is cut into tiles of 32 rows x 12 columns; one tile = one ring stage of 384 slots, slot (j, lane) =
//     pixel (x = 12 tx + j, y = 32 ty + lane), tiles ordered column-of-tiles major (stage c = tx * tiles_y + ty).  A stage is
//     3840 bytes: inverse depths f32 [2][32][6] (the six values a lane needs for half a tile are contiguous: three 8-byte
//     shared loads), gradients half2 [2][32][6], template values f16 [2][32][6] = 10 B per slot; pixels outside the image
//     or without depth carry a NaN inverse depth (never inside, never in H_total), rows below the image a zero one (the
//     align kernel's dead lanes).  x, y never travel: they follow from the stage and lane.
//     (12 columns = two half tiles of six words: a multiple of the three-slot software pipeline of the align kernel's hot
//     loop, whose unrolled body of six words still fits the instruction cache.)
//   * the frame pyramids the align kernel SAMPLES live in gather-enabled 2D CUDA arrays ("atlas pages", u8 texels read as
//     texel / 255, or f16 with -DVORS_TEX_F16=1): texture x = image y.  Stream s, level l occupies the cell at
//     (ox, oy) = ((s % per_row) * cell_w, (s / per_row) * cell_h + lvl_y[l]) with cell_w = rows_0 + 2 rounded up to a multiple of 4, lvl_y[l] = sum_{k<l}
//     (cols_k + 2): the two texels after every level's last row / column are never written (zero) and serve as the zero page.
//     k_atlas_fill writes four texels per surface store, which is why cells start at multiples of four texels.
//     One tld4 returns the 2x2 footprint of a warped candidate.  The linear pyramids stay: keyframe build and the exact
//     re-evaluation of boundary-band candidates read them.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vors_b200.h"
#include "lie.cuh"

// compile-time variants (scripts/build_variants.sh passes them to every translation unit)
#ifndef VORS_TEX
#define VORS_TEX 1      // 1: the align kernel samples the current frame with one texture gather (tld4) per candidate from the
                        // atlas pages; 0: four byte loads from the linear pyramid
#endif
#ifndef VORS_TEX_F16
#define VORS_TEX_F16 0  // atlas texels: 0 = u8 read as texel / 255 (normalised float), 1 = f16 (exact values, twice the bytes)
#endif

namespace vors {

constexpr int kMaxLevels = VORS_MAX_LEVELS;
constexpr int kCompactBlock = 1024;  // elements per ordered-compaction block
constexpr int kTraceCap = 256;       // trace records kept per alignment
constexpr int kMaxTeam = 160;        // CTAs cooperating on one alignment (<= one per SM)
constexpr int kChunk = 64;          // candidates per chunk-blocked record (768 B)
#ifndef VORS_STAGE_CHUNKS
#define VORS_STAGE_CHUNKS 4
#endif
constexpr int kPtAlign = VORS_STAGE_CHUNKS * kChunk;  // a level's candidate block is padded to this many candidates (one align-kernel ring stage)
constexpr int kHStride = 24;         // doubles per (stream, level) in the H_total table (21 used)
constexpr int kNumAcc = 29;          // finished pass: sum r^2, n_inside, g[6], H[21] (upper triangle)
constexpr int kNumRaw = 32;          // raw pass accumulators: sum r^2, n_inside, 9 gradient moments (or g[6]), H_outside[21]

constexpr int kTileRows = 32;        // tiled dense records: a tile is 32 rows (lanes) x kTileCols columns (words of a stage)
constexpr int kTileCols = 12;
constexpr int kTileSlots = kTileRows * kTileCols;
constexpr int kTileBytes = kTileSlots * 10;  // rho f32 | grad half2 | template f16
constexpr int kTileWords = kTileBytes / 4;
constexpr int kTileHalfCols = kTileCols / 2;  // words per unrolled half tile
static_assert(kTileCols % 2 == 0 && kTileHalfCols % 3 == 0 && kTileHalfCols % 2 == 0, "half tiles of a multiple of three words (pipeline slots), even (8-byte loads)");

struct Geom {
    int L;
    int rows[kMaxLevels], cols[kMaxLevels];
    int off[kMaxLevels];         // pixel offset of level l inside a slab
    int pt_off[kMaxLevels];      // element offset of level l inside a candidate-stream slab (kPtAlign-aligned for TMA)
    int pt_total;                // candidate-stream slab extent per stream (kPtAlign-aligned)
    int blk_off[kMaxLevels + 1]; // compaction-block offset of level l (kCompactBlock px per block)
    int pix_total;               // pixels of all levels
    int pix_stride;              // per-stream slab stride: pix_total + zero page (see above)
    int blk_total;
    // tiled dense records (see the header comment)
    int tiles_y[kMaxLevels], tiles_x[kMaxLevels];
    int tile_off[kMaxLevels];    // first tile (= stage) of level l inside a stream's tiled record slab
    int tile_total;              // tiles per stream
    // texture atlas cell of one stream
    int cell_w, cell_h;          // rows_0 + 2 rounded up to a multiple of 4, sum (cols_l + 2)
    int lvl_y[kMaxLevels];       // texel row of level l inside the cell
    int per_row, per_page;       // cells per atlas row / per atlas page
};

// word offsets of slot (j, lane) of a tile inside its kTileWords-word stage
__host__ __device__ __forceinline__ int tile_rho_word(int j, int lane) {
    return (j / kTileHalfCols) * (kTileRows * kTileHalfCols) + lane * kTileHalfCols + j % kTileHalfCols;
}
__host__ __device__ __forceinline__ int tile_grad_word(int j, int lane) { return kTileSlots + tile_rho_word(j, lane); }
__host__ __device__ __forceinline__ int tile_tmpl_half(int j, int lane) { return 4 * kTileSlots + tile_rho_word(j, lane); }  // index in halves

// ---- candidate record fields -----------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t rec_pack_pk(int x, int y, uint32_t tmpl) { return uint32_t(x) | (uint32_t(y) << 12) | (tmpl << 24); }
__host__ __device__ __forceinline__ uint32_t rec_x(uint32_t pk) { return pk & 0xFFFu; }
__host__ __device__ __forceinline__ uint32_t rec_y(uint32_t pk) { return (pk >> 12) & 0xFFFu; }
__host__ __device__ __forceinline__ uint32_t rec_tmpl(uint32_t pk) { return pk >> 24; }
// gradient word: i16 pair of the gradient slab (gx | gy<<16) -> half2(gx, gy)
__device__ __forceinline__ uint32_t rec_pack_grad(uint32_t grad_i16x2) {
    const __half2 h = __floats2half2_rn(float(int(short(grad_i16x2 & 0xFFFFu))), float(int(short(grad_i16x2 >> 16))));
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float rec_gx(uint32_t gr) { return __half2float(__ushort_as_half((unsigned short)(gr & 0xFFFFu))); }
__device__ __forceinline__ float rec_gy(uint32_t gr) { return __half2float(__ushort_as_half((unsigned short)(gr >> 16))); }

// One pyramid level of one alignment, as the align kernel sees it.
// word index of field f (0 pk, 1 idepth, 2 grad) of candidate i inside a level's chunk-blocked block
__host__ __device__ __forceinline__ size_t pt_word(int i, int f) { return size_t(i / kChunk) * (3 * kChunk) + size_t(f) * kChunk + size_t(i % kChunk); }

struct LevelJob {
    const uint32_t* pts;  // chunk-blocked candidates of this level (tiled dense keyframes: the level's tiles, kTileWords each)
    const uint8_t* img;  // current frame, this level
    const int* n_ptr;    // number of candidates (device memory: written by the compaction kernels)
    const double* h_total;  // sum over ALL candidates of J J^T, 21 upper-triangle entries (k_h_total)
    uint32_t* defer;     // bitmap of the level's candidate slots the align kernel's hot loop deferred (all zero between passes)
    uint32_t* defer_far; // same, for the slots it found outside for sure
    int rows, cols;
    float zero_u, zero_v;  // integer-valued image coordinates whose 2x2 footprint is the slab's zero page
    float tex_ku, tex_kv;  // texture path: (2^23 + floor(u)) - tex_ku = atlas texel row of the footprint's shared corner (same for v)
    int tiles_y, n_tiles;  // tiled dense records: tiles per tile column, tiles of the level
    Intrinsics k;
};

struct AlignJob {
    LevelJob lv[kMaxLevels];
    int lvl_first;    // coarsest level to run (nb_levels-1 for a full alignment)
    int lvl_last;     // finest level to run (0)
    int flow_level;   // level whose candidates feed the optical-flow test (nb_levels-1); <0 = skip
    int pass_only;    // 1: single evaluation at `init` on lvl_first, no LM loop (vors_align_pass)
    unsigned long long tex;  // cudaTextureObject_t of the atlas page holding this stream's current frame pyramid
};

struct AlignResult {
    Pose model;  // lm_model after the last successful level
    int status;  // VORS_OK / VORS_OPTIMIZATION_FAILED
    float optical_flow;
    int n_iters[kMaxLevels];
    float energy[kMaxLevels];
    int n_points[kMaxLevels];
    int n_passes;
    int trace_len;
    unsigned long long point_passes;
    unsigned long long t_begin_ns, t_end_ns;  // %globaltimer when the team picked the job up / finished it (load-balance diagnosis)
    // pass_only outputs
    float pass_energy;
    int pass_n_inside;
    float pass_g[6];
    float pass_H[21];
};

struct TeamScratch {
    double part[2][kMaxTeam][40];
    unsigned int counter;
    unsigned int pad[31];
};

struct AlignParams {
    const AlignJob* jobs;
    const Pose* init;  // per job lm_model prior: current_frame_pose^-1 * keyframe_pose (inverse_compositional.rs:177)
    AlignResult* results;
    vors_trace_rec* trace;  // n_jobs * kTraceCap, or nullptr
    TeamScratch* scratch;   // one per team (only touched when team > 1)
    unsigned int* job_counter;  // zeroed before the launch: jobs beyond the first wave are handed out through it (team == 1)
    int n_jobs;
    int team;
    // LM constants (lm_optimizer.rs:115,157,173,179,186)
    float lm_coef_init, lm_coef_reject_mult, lm_coef_accept_mult, energy_delta_stop;
    int max_iters, fixed_iters;
    int has_skew;  // 0 selects the zero-skew Jacobian specialisation
    int tiled;     // 1: the keyframes hold tiled dense records with implicit coordinates (dense, zero skew, plain L2)
    float huber_delta;  // > 0 selects the Huber-weighted kernel variant (extension)
};

// Row J: warp_jacobian_at (inverse_compositional.rs:313-341), used by the align kernel (J is recomputed
// per pass in registers instead of being stored), by k_h_total and by the jacobian export kernel.  Same
// expression order as the reference except that `c / fu` is evaluated as c * (1/fu) (loop-invariant
// reciprocal).  kSkew = false is the zero-skew specialisation (every intrinsics set the reference ships,
// src/dataset/tum_rgbd.rs:23-51, has skew 0): algebraically identical, 7 fewer instructions.
template <bool kSkew = true>
__device__ __forceinline__ void jacobian_centred(float gu, float gv, float a, float b, float rho, const Intrinsics& k, float J[6]);
template <bool kSkew = true>
__device__ __forceinline__ void jacobian_at(float gu, float gv, float u, float v, float rho, const Intrinsics& k, float J[6]) {
    jacobian_centred<kSkew>(gu, gv, u - k.cx, v - k.cy, rho, k, J);
}
// same, from a = u - cu, b = v - cv
template <bool kSkew>
__device__ __forceinline__ void jacobian_centred(float gu, float gv, float a, float b, float rho, const Intrinsics& k, float J[6]) {
    const float _fv = 1.0f / k.fy;
    const float _fu = 1.0f / k.fx;
    if (kSkew) {
        const float c = a * k.fy - k.s * b;
        const float _fuv = 1.0f / (k.fx * k.fy);
        J[0] = gu * rho * k.fx;
        J[1] = rho * (gu * k.s + gv * k.fy);
        J[2] = -rho * (gu * a + gv * b);
        J[3] = gu * (-a * b * _fv - k.s) + gv * (-b * b * _fv - k.fy);
        J[4] = gu * (a * c * _fuv + k.fx) + gv * (b * c * _fuv);
        J[5] = gu * (-k.fx * k.fx * b + k.s * c) * _fuv + gv * (c * _fu);
    } else {
        const float bf = b * _fv, af = a * _fu;  // b / fv, a / fu  (c = a fv, c / (fu fv) = a / fu)
        J[0] = gu * (rho * k.fx);
        J[1] = gv * (rho * k.fy);
        J[2] = -rho * (gu * a + gv * b);
        J[3] = -(gu * (a * bf) + gv * (b * bf + k.fy));
        J[4] = gu * (a * af + k.fx) + gv * (b * af);
        J[5] = gv * (a * (k.fy * _fu)) - gu * (b * (k.fx * _fv));
    }
}

// The reference's own expression (inverse_compositional.rs:326-340) with every operation rounded separately (no FMA
// contraction, true divisions): the reference's J bit for bit, whatever code surrounds the call.  For the places that
// evaluate J once per keyframe (H_total): there the contraction the compiler picks for jacobian_at would otherwise change
// with the loop structure of the kernel, and with it the last bit of J.
__device__ __forceinline__ void jacobian_exact(float gu, float gv, float u, float v, float z, const Intrinsics& k, float J[6]) {
    const float a = __fsub_rn(u, k.cx), b = __fsub_rn(v, k.cy);
    const float c = __fsub_rn(__fmul_rn(a, k.fy), __fmul_rn(k.s, b));
    const float _fv = __fdiv_rn(1.0f, k.fy);
    const float _fuv = __fdiv_rn(1.0f, __fmul_rn(k.fx, k.fy));
    J[0] = __fmul_rn(__fmul_rn(gu, z), k.fx);
    J[1] = __fmul_rn(z, __fadd_rn(__fmul_rn(gu, k.s), __fmul_rn(gv, k.fy)));
    J[2] = __fmul_rn(-z, __fadd_rn(__fmul_rn(gu, a), __fmul_rn(gv, b)));
    J[3] = __fadd_rn(__fmul_rn(gu, __fsub_rn(__fmul_rn(__fmul_rn(-a, b), _fv), k.s)),
                     __fmul_rn(gv, __fsub_rn(__fmul_rn(__fmul_rn(-b, b), _fv), k.fy)));
    J[4] = __fadd_rn(__fmul_rn(gu, __fadd_rn(__fmul_rn(__fmul_rn(a, c), _fuv), k.fx)), __fmul_rn(gv, __fmul_rn(__fmul_rn(b, c), _fuv)));
    J[5] = __fadd_rn(__fmul_rn(__fmul_rn(gu, __fadd_rn(__fmul_rn(__fmul_rn(-k.fx, k.fx), b), __fmul_rn(k.s, c))), _fuv),
                     __fmul_rn(gv, __fdiv_rn(c, k.fx)));
}

// ---- launchers implemented in image_kernels.cu -------------------------------------------------
// `items`: device array of stream indices the kernels operate on (m of them); nullptr = 0..m-1.
struct Launcher {
    cudaStream_t stream;
    unsigned long long launches = 0;
};

void launch_transpose_u8(Launcher& L, const uint8_t* in_rowmajor, uint8_t* out_slab, size_t out_stride, const int* items,
                         int m, int rows, int cols);
void launch_transpose_u16(Launcher& L, const uint16_t* in_rowmajor, uint16_t* out_slab, size_t out_stride, const int* items,
                          int m, int rows, int cols);
void launch_copy_items_u16(Launcher& L, const uint16_t* in, uint16_t* out_slab, size_t out_stride, const int* items, int m,
                           size_t count);
void launch_pyramid(Launcher& L, const Geom& g, uint8_t* pyr_slab, const int* items, int m);
void launch_gradients(Launcher& L, const Geom& g, int scharr, const uint8_t* pyr_slab, uint32_t* grad_slab, uint16_t* g2_slab,
                      const int* items, int m);
void launch_c2f(Launcher& L, const Geom& g, uint16_t thresh, const uint16_t* g2_slab, uint8_t* mask_slab, const int* items,
                int m);
void launch_idepth(Launcher& L, const Geom& g, const uint16_t* depth_slab, size_t depth_stride, const uint8_t* mask_slab,
                   int dense, float scale, float variance, int similar, float* idepth_slab, float* weight_slab, const int* items,
                   int m);
void launch_compact(Launcher& L, const Geom& g, const float* idepth_slab, const uint8_t* pyr_slab, const uint32_t* grad_slab,
                    int* blk_count, int* n_points, uint32_t* pts_slab, const int* items, int m);
// `h_part`: scratch of h_total_scratch_doubles() doubles per listed stream (partial sums per chunk of a level)
size_t h_total_scratch_doubles(const Geom& g, bool tiled);
void launch_h_total(Launcher& L, const Geom& g, const Intrinsics* intr, const uint32_t* pts_slab, int* n_points, double* h_part,
                    double* h_total, const int* items, int m);
// tiled dense keyframes: records + H_total + candidate counts of all levels (replaces launch_compact + launch_h_total);
// grad_slab == nullptr: the gradients are computed from the pyramid on the fly (`scharr` as for launch_gradients)
void launch_tile_records(Launcher& L, const Geom& g, const Intrinsics* intr, int scharr, const float* idepth_slab, const uint8_t* pyr_slab,
                         const uint32_t* grad_slab, int* n_points, uint32_t* pts_slab, double* h_part, double* h_total, const int* items,
                         int m);
// linear frame pyramids of m streams (items, or first .. first + m - 1 when items == nullptr and pyr_slab points at stream
// `first`'s slab) -> their cells of the atlas pages
constexpr int kMaxAtlasPages = 8;
struct AtlasPages {
    cudaSurfaceObject_t surf[kMaxAtlasPages];
    int f16;
};
void launch_atlas_fill(Launcher& L, const Geom& g, const uint8_t* pyr_slab, const AtlasPages& pages, const int* items, int m, int first);
void launch_jacobians(Launcher& L, const uint32_t* pts_level, int n, Intrinsics k, float* out6);
void launch_se3_exp(Launcher& L, const float* xi6, Pose* out);
void launch_lie(Launcher& L, int op, const float* in, float* out);

// ---- implemented in dso_kernels.cu ------------------------------------------------------------------
void launch_sqnorm_direct(Launcher& L, const uint8_t* img, int rows, int cols, int as_magnitude, uint16_t* out);
void launch_bloc_sqnorm(Launcher& L, const uint8_t* fine, int rows_in, int rows, int cols, uint16_t* out);
size_t dso_workspace_bytes(int rows, int cols);
int dso_select_device(Launcher& L, const uint16_t* d_grad, int rows, int cols, int nb_target, int nb_iterations_left,
                      unsigned long long seed, uint8_t* d_mask_out, uint8_t* ws, int* h_pinned_scratch, int* used_random,
                      int* nb_candidates_out);

// ---- implemented in align_kernel.cu ---------------------------------------------------------------
struct AlignLaunchInfo {
    int block;
    int max_resident_ctas;  // co-resident CTAs of the align kernel on this device
    int sm_count;
};
cudaError_t align_query(AlignLaunchInfo* info);
cudaError_t launch_align(Launcher& L, const AlignParams& p, int n_teams);

}  // namespace vors
