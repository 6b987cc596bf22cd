// align_kernel.cu — the persistent direct-alignment kernel, sm_100a.
//
// One launch runs COMPLETE alignments: for every job (keyframe, frame, prior) a team of `team` CTAs
// walks the pyramid coarse to fine and, per level, runs the reference's Levenberg-Marquardt loop
// (src/math/optimizer.rs:57-70 + src/core/track/lm_optimizer.rs:113-192) entirely on the device:
//
//   pass      every warp streams its share of the level's candidate records (12 B per candidate, 256 candidates per
//             stage) through its own shared-memory ring with TMA bulk copies (cp.async.bulk, SASS UBLKCP) issued
//             kStages - 1 stages ahead, so HBM latency is covered by bytes in flight instead of by occupancy, and
//             evaluates eight candidates per lane per stage, branch-free on the common path: warp with the folded
//             3x4 matrix of lie.cuh (lm_optimizer.rs:213-219), inside test with a safety margin, f32 bilinear sample
//             of u8 texels (lm_optimizer.rs:227-251), residual against the template value, sum r^2 and nine
//             moments of (gu r, gv r) from which g = sum J r is assembled once per pass (inverse_compositional.rs:
//             313-341; J itself is never formed on the common path).  Candidates that fall outside read a zero page
//             (exact zero contributions) and add J J^T to per-thread shared-memory sums: H over the inside set is
//             H_total - H_outside, H_total precomputed per keyframe level (k_h_total).  Slots within 1/128 px of an
//             inside-test boundary are deferred through a bitmap and re-evaluated after the hot loop with the
//             reference's own operation order (eval_energy + compute_eval_data, lm_optimizer.rs:68-107, fused);
//   reduce    warp shuffles -> shared memory -> f64 per-CTA partials -> (team > 1) peer partials
//             through global memory with one counter barrier per pass; fixed order => deterministic;
//   decide    the last warp of every CTA redundantly replays accept / reject / stop (lm_optimizer.rs:
//             140-192), damps, solves the 6x6 system by Cholesky, applies se3::exp and the first
//             order renormalisation (lm_optimizer.rs:123-136, 198-209) and publishes the next
//             candidate model's warp matrix; no host round trip anywhere in the loop.
//
// team == 1 is the throughput configuration (one alignment per CTA, two CTAs per SM so one CTA's
// serial solve overlaps the other's pass); team > 1 trades efficiency for latency on few streams.
// No tensor cores: the work is ~80 scalar instructions per 12-byte candidate, bounded by
// instruction issue, not by a dense contraction.
#include <cooperative_groups.h>
#include <cstdio>

#include "vors_device.cuh"

namespace vors {

namespace {

#ifndef VORS_CHUNK_UNROLL
#define VORS_CHUNK_UNROLL 4
#endif
#ifndef VORS_TIMING
#define VORS_TIMING 0
#endif
#ifndef VORS_WARPS
#define VORS_WARPS 10
#endif
#ifndef VORS_MIN_CTAS
#define VORS_MIN_CTAS 2
#endif
#ifndef VORS_X_MAGIC
#define VORS_X_MAGIC 0  // 1: x through the 2^23 magic number (2 LOP3 + FADD) instead of LOP3 + I2FP
#endif
#ifndef VORS_STAGES
#define VORS_STAGES 2
#endif
#ifndef VORS_OLD_SERIAL
#define VORS_OLD_SERIAL 0  // 1: the round-1 serial LM round (one lane, IEEE divisions) instead of serial_round
#endif
#ifndef VORS_FRND
#define VORS_FRND 0     // 1: floor by FRND.FLOOR (one XU-pipe instruction) instead of the round-down magic-number add + subtract
#endif
constexpr int kWarps = VORS_WARPS;            // warps per CTA; every warp refills its own TMA ring
constexpr int kConsumers = kWarps * 32;
constexpr int kBlock = kConsumers;
constexpr int kChunkUnroll = VORS_CHUNK_UNROLL;  // chunks of a stage unrolled in the hot loop (instruction-cache footprint)
constexpr int kSmallWordsPerWarp = 4;           // levels of at most this many 32-candidate words per warp skip the TMA ring
constexpr int kSerialWarp = kWarps - 1;       // runs the serial part of every LM round
constexpr int kMinCtasPerSm = VORS_MIN_CTAS;  // 10 warps x 2 CTAs: register cap 96, 20 warps per SM
constexpr int kStageChunks = VORS_STAGE_CHUNKS;  // chunk-blocked records per ring stage
constexpr int kStageCand = kStageChunks * kChunk;  // candidates per stage, 2 per lane per chunk (== kPtAlign)
constexpr int kStageWordsBm = kStageCand / 32;     // bitmap words per stage
constexpr int kStages = VORS_STAGES;          // TMA ring depth per warp
constexpr int kStageWords = kStageChunks * 3 * kChunk;  // per chunk: pk[64] | idepth[64] | grad[64]
constexpr uint32_t kStageBytes = kStageWords * 4;
constexpr int kRingWords = kStageWords > kTileWords ? kStageWords : kTileWords;  // ring slot: a stage of either record layout
constexpr uint32_t kRingBytes = kRingWords * 4;
constexpr int kTileWordsBm = kTileSlots / 32;  // bitmap words per tile
constexpr int kHsmStride = 21;  // odd: scalar accesses of a warp hit 32 distinct banks
constexpr int kStashCap = 16;   // boundary-band candidates a warp stashes per pass
constexpr int kNearWords = 32;  // beyond that: flagged bitmap words a warp remembers per pass; beyond that: bitmap scan
static_assert(kStageCand == kPtAlign, "levels are padded to whole ring stages");

struct LmShared {
    float M[12];
    // kept state = last accepted evaluation (lm_optimizer.rs:31-40 `EvalData`)
    float keptH[21];
    float keptg[6];
    float keptE;
    Pose kept_model;
    Pose cand_model;
    Pose out_model;  // lm_model of Tracker::track: result of the last successful level
    float lam;
    int iter;
    int init_phase;
    int cont;
    int failed;
    int n_passes;
    int trace_len;
    unsigned long long point_passes;
    float warp_part[kWarps][32];   // per-warp sums of the pass accumulators (E, n, 9 moments / g[6] / Huber: g[6] and H[21])
    double hout[kWarps][21];       // per-warp sum of J J^T over the candidates outside for sure, cumulative over a level's passes
    double h_total[21];            // the level's H_total (k_h_total), cached for the per-pass serial part
    double raw[kNumRaw];           // CTA / team totals of the raw accumulators
    double tot[32];                // finished pass: sum r^2, n_inside, g[6], H[21]
#if VORS_TIMING
    long long dbg[8][4];
    long long dbgw[8][4];
    long long dbgs[8][4];
#endif
    // boundary-band candidates flagged by each warp in this pass, stashed with what their exact re-evaluation needs
    // (slot, a = x - cx, b = y - cy, inverse depth, gradient bits, template value): no record is read twice
    uint32_t stash[kWarps][kStashCap][6];
    uint32_t near_words[kWarps][2 * kNearWords];  // (word, mask) pairs flagged by each warp in this pass (second tier, see deferred_pass)
    alignas(8) unsigned long long full_bar[kWarps][kStages];
    alignas(128) float ring[kWarps][kStages * kRingWords];
    // per-thread J J^T accumulators of the hot loop (candidates that changed sides of the frame border): 21 floats at an
    // odd stride (conflict-free scalar accesses; this rare path trades vector accesses for 9 KB of shared memory); zero
    // between passes
    alignas(16) float hsm[kConsumers][kHsmStride];
};

// ---- TMA bulk copy + mbarrier primitives (PTX ISA: cp.async.bulk, mbarrier) ------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// candidates are streamed once per pass: evict-first in L2 so the frame images (re-read by every pass) stay resident
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}

// int -> float conversions: one I2FP each (the loop is issue-bound: the two-instruction 2^23 magic-number conversion
// only pays where it fuses with work that is needed anyway, i.e. the floor of the warped coordinates).
__device__ __forceinline__ float u2f(uint32_t v) { return float(v); }
// 1/x: MUFU.RCP seed + one Newton step (2 FMAs) = correctly rounded to within 1 ulp without the slow-path range
// checks of an IEEE division; x = 0 / inf / NaN give inf / NaN, which the inside test rejects like the reference.
#ifndef VORS_RCP_NEWTON
#define VORS_RCP_NEWTON 1  // 0: the raw MUFU.RCP seed (1 ulp) without the Newton step (measured: no faster, not the default)
#endif
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
#if VORS_RCP_NEWTON
    return fmaf(r, fmaf(-x, r, 1.0f), r);
#else
    return r;
#endif
}

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// index of (a, b), a <= b, in the packed upper triangle
__device__ __host__ constexpr int tri(int a, int b) { return a * 6 - a * (a - 1) / 2 + (b - a); }

// Raw per-thread accumulators of one pass.  Zero skew (every intrinsics set the reference ships): instead of the
// six Jacobian entries, nine moments of (p, q) = (gu r, gv r) and s = a p + b q from which g = sum J r is assembled once
// per pass in f64 (J is linear in (gu, gv) with coefficients polynomial in a = x - cu, b = y - cv and idepth,
// inverse_compositional.rs:326-340; a b p + b^2 q = b s and a^2 p + a b q = a s) - 13 instead of 29 instructions per
// candidate.  With skew the plain g[6].
// kHuber (extension, vors_config.huber_delta > 0): the weights depend on the residual, so neither the moments nor
// H = H_total - H_outside apply; g[6] and the 21 entries of H are summed directly (s[0..5], s[6..26]).
template <bool kHuber>
struct Acc {
    float e;
    float s[kHuber ? 27 : 9];
};
enum { kSrp, kSrq, kSrs, kSbs, kSas, kSq, kSp, kSbp, kSaq };
constexpr int kNumMoments = 9;
constexpr int kRawH = 2 + kNumMoments;  // raw[kRawH ..]: H_outside[21]
static_assert(kRawH + 21 == kNumRaw, "raw totals: sum r^2, n_inside, moments, H_outside");

__device__ __forceinline__ void accumulate_moments(Acc<false>& acc, float gu, float gv, float a, float b, float rho, float r) {
    const float p = gu * r, q = gv * r;
    const float s = fmaf(a, p, b * q);
    acc.s[kSrp] = fmaf(rho, p, acc.s[kSrp]);
    acc.s[kSrq] = fmaf(rho, q, acc.s[kSrq]);
    acc.s[kSrs] = fmaf(rho, s, acc.s[kSrs]);
    acc.s[kSbs] = fmaf(b, s, acc.s[kSbs]);
    acc.s[kSas] = fmaf(a, s, acc.s[kSas]);
    acc.s[kSp] += p;
    acc.s[kSq] += q;
    acc.s[kSbp] = fmaf(b, p, acc.s[kSbp]);
    acc.s[kSaq] = fmaf(a, q, acc.s[kSaq]);
}

// Huber-weighted contribution of one inside candidate: rho_delta(r) to the energy, w J r to g, w J J^T to H.
__device__ __forceinline__ void accumulate_huber(Acc<true>& acc, const float (&J)[6], float r, float delta) {
    const float ar = fabsf(r);
    const bool quad = ar <= delta;
    const float w = quad ? 1.0f : delta / ar;
    acc.e += quad ? r * r : delta * (2.0f * ar - delta);
    const float wr = w * r;
#pragma unroll
    for (int c = 0; c < 6; ++c) acc.s[c] = fmaf(J[c], wr, acc.s[c]);
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        const float wj = w * J[c];
#pragma unroll
        for (int d = c; d < 6; ++d) acc.s[6 + tri(c, d)] = fmaf(wj, J[d], acc.s[6 + tri(c, d)]);
    }
}


// The reference's own warp arithmetic (lm_optimizer.rs:213-219 with camera.rs:126-140 and nalgebra's
// `UnitQuaternion * Vector3` = t*w + v x t + p, t = 2 (v x p)), every product / sum / quotient rounded
// separately like rustc emits it.  Only used for candidates that land within kBandPx of an inside-test
// boundary, where the sign of the last ulp decides membership (e.g. the x = 0 column under an identity
// model): there the folded matrix and the reference may disagree, so the reference's order decides.
__device__ __noinline__ float2 warp_exact(const Pose& m, const Intrinsics& k, float x, float y, float rho) {
    const float z = __fdiv_rn(1.0f, rho);
    const float Y = __fdiv_rn(__fmul_rn(__fsub_rn(y, k.cy), z), k.fy);
    const float X = __fdiv_rn(__fsub_rn(__fmul_rn(__fsub_rn(x, k.cx), z), __fmul_rn(k.s, Y)), k.fx);
    const float qi = m.q.i, qj = m.q.j, qk = m.q.k, qw = m.q.w;
    // t = (v x p) * 2
    const float tx = __fmul_rn(__fsub_rn(__fmul_rn(qj, z), __fmul_rn(qk, Y)), 2.0f);
    const float ty = __fmul_rn(__fsub_rn(__fmul_rn(qk, X), __fmul_rn(qi, z)), 2.0f);
    const float tz = __fmul_rn(__fsub_rn(__fmul_rn(qi, Y), __fmul_rn(qj, X)), 2.0f);
    // c = v x t
    const float cx = __fsub_rn(__fmul_rn(qj, tz), __fmul_rn(qk, ty));
    const float cy = __fsub_rn(__fmul_rn(qk, tx), __fmul_rn(qi, tz));
    const float cz = __fsub_rn(__fmul_rn(qi, ty), __fmul_rn(qj, tx));
    // (t*w + c + p) + translation
    const float X2 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(tx, qw), cx), X), m.t.x);
    const float Y2 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(ty, qw), cy), Y), m.t.y);
    const float Z2 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(tz, qw), cz), z), m.t.z);
    const float px = __fadd_rn(__fadd_rn(__fmul_rn(k.fx, X2), __fmul_rn(k.s, Y2)), __fmul_rn(k.cx, Z2));
    const float py = __fadd_rn(__fmul_rn(k.fy, Y2), __fmul_rn(k.cy, Z2));
    return make_float2(__fdiv_rn(px, Z2), __fdiv_rn(py, Z2));
}

constexpr float kBandPx = 1.0f / 128.0f;
// floor without F2I: for 0 <= x < 2^22 and an integer-valued float C in [2^23, 2^23 + 2^22], x + C rounded DOWN is the
// float whose bit pattern is bits(C) + floor(x).  Texel offsets are formed from those raw bit patterns with 32-bit
// arithmetic; the constant they carry, bits(Cu) * rows + bits(Cv), is folded (mod 2^32) into the image base pointer.
constexpr uint32_t kMagicBits = 0x4B000000u;  // 2^23

// Per-level constants of the pass (warp-uniform).
struct LevelConst {
    float cx, cy;
    float hu, hv;        // centre of the inside range [0, W-2) x [0, H-2): (W-2) / 2, (H-2) / 2
    float lo_u, lo_v;    // |u - hu| < lo_u and |v - hv| < lo_v: inside for sure (half extent - band)
    float hi_u, hi_v;    // |u - hu| > hi_u or  |v - hv| > hi_v: outside for sure (half extent + band)
    float wm2, hm2;
    float zero_u, zero_v;
    float magic_u, magic_v;     // floor constants Cu, Cv (see kMagicBits)
    float tex_ku, tex_kv;       // texture path: (2^23 + floor) - tex_k = atlas coordinate of the footprint's shared corner
    unsigned long long tex;     // atlas page of this job's frame
    int tiles_y;                // tiled records: tiles per tile column
    float inv_tiles_y;
    uint32_t rows;
    int n;
    const uint8_t* img_biased;  // img - ((bits(Cu) * rows + bits(Cv)) mod 2^32)
    const uint8_t* img;
    const uint32_t* pts;        // the level's chunk-blocked candidates (deferred pass)
    Intrinsics k;
    double inv_fx, inv_fy, k01;  // 1/fx, 1/fy, -s/(fx fy): the f64 divisions of the per-pass serial part, done once per level
    double ga[6], gb[6];         // g from the gradient moments (serial_round)
    float huber_delta;
};

// Chooses Cu so that (bits(Cu) * rows + bits(Cv)) mod 2^32 plus any texel offset the level can produce (its image and the
// slab's zero page: at most zero_u * rows + zero_v + rows + 1) cannot wrap.  The shift Cu - 2^23 stays far below 2^22
// (it is about the slab extent divided by the level's rows), so u + Cu stays inside [2^23, 2^24) where floats are integers.
__device__ __forceinline__ void choose_floor_magic(LevelConst& c, const uint8_t* img) {
    const uint32_t max_off = uint32_t(c.zero_u) * c.rows + uint32_t(c.zero_v) + c.rows + 2u;
    uint32_t du = 0u;
    const uint32_t bv = kMagicBits;
    uint32_t c32 = kMagicBits * c.rows + bv;
    if (c32 > 0xFFFFFFFFu - max_off) {  // shift the constant past the wrap-around
        du = (max_off + c.rows - 1u) / c.rows + 1u;
        c32 = (kMagicBits + du) * c.rows + bv;
    }
    c.magic_u = __uint_as_float(kMagicBits + du);
    c.magic_v = __uint_as_float(bv);
    c.img_biased = img - size_t(c32);
}

// A candidate between its gather issue (front) and its arithmetic (back).  The consumer loop runs the front of
// candidate q+1 before the back of candidate q, so texel latency hides behind ~80 instructions of the same warp.
struct Front {
    uint32_t pk, gr;
    float a, b, rho, fa, fb;
#if VORS_TEX
    float t00, t10, t01, t11;  // texels as the texture unit returns them (texel / 255, or the exact value with f16 texels)
#else
    uint32_t t00, t10, t01, t11;
#endif
};
#if VORS_TEX
#if VORS_TEX_F16
constexpr float kTexScale = 1.0f;
#else
constexpr float kTexScale = 255.0f;
#endif
// One texture gather = the 2x2 footprint.  (tx, ty) is the corner shared by the four texels (integer-valued, half a texel
// away from every footprint boundary: the texture unit's fixed-point coordinates cannot select another footprint than the
// kernel's own floor).  Texture x = image y: .w = (y, x), .z = (y+1, x), .x = (y, x+1), .y = (y+1, x+1).
__device__ __forceinline__ void tex_gather(unsigned long long tex, float tx, float ty, float& t00, float& t10, float& t01, float& t11) {
    asm("tld4.r.2d.v4.f32.f32 {%0, %1, %2, %3}, [%4, {%5, %6}];" : "=f"(t01), "=f"(t11), "=f"(t10), "=f"(t00) : "l"(tex), "f"(tx), "f"(ty));
}
#endif

// The kernel's shared state: the LM / ring block in dynamic shared memory, the per-level constants in static.
extern __shared__ __align__(128) unsigned char smem_raw[];
__device__ __forceinline__ LmShared& lm_shared() { return *reinterpret_cast<LmShared*>(smem_raw); }
__shared__ LevelConst s_lc;

// Deferred candidates.  The hot loop only evaluates candidates that are inside the frame for sure (fast warp, margin of
// kBandPx); every other slot - padding, candidates near an inside-test boundary, candidates that fall outside - is
// redirected to the zero page there (exact zero contributions) and flagged in a per-level bitmap in global memory
// (one bit per candidate slot: word = slot / 32, bit = slot % 32 = the lane that owned it).
// After the hot loop each warp revisits the flagged slots of its own stages: the reference's own arithmetic
// (lm_optimizer.rs:213-231) decides membership; inside candidates are evaluated in full, outside ones contribute
// J J^T to H_outside.  Keeping all of this (and its function calls) out of the hot loop keeps that loop call-free.

// Fields of candidate slot `i` of a level, from its record in global memory (deferred pass, optical flow).
struct SlotRec {
    float x, y, rho, gu, gv, tmpl;
    uint32_t gr;
};
template <bool kTiled>
__device__ __forceinline__ SlotRec load_slot(const uint32_t* __restrict__ pts, int tiles_y, int i) {
    SlotRec r;
    if constexpr (kTiled) {
        const int st = i / kTileSlots, j = (i / kTileRows) % kTileCols, ln = i % kTileRows;
        const int tx = st / tiles_y, ty = st - tx * tiles_y;
        const uint32_t* w = pts + size_t(st) * kTileWords;
        r.x = float(kTileCols * tx + j);
        r.y = float(kTileRows * ty + ln);
        r.rho = __uint_as_float(__ldg(w + tile_rho_word(j, ln)));
        const uint32_t gr = __ldg(w + tile_grad_word(j, ln));
        r.gr = gr;
        r.gu = rec_gx(gr);
        r.gv = rec_gy(gr);
        r.tmpl = __half2float(__ushort_as_half(__ldg(reinterpret_cast<const unsigned short*>(w) + tile_tmpl_half(j, ln))));
    } else {
        const uint32_t pk = __ldg(pts + pt_word(i, 0)), gr = __ldg(pts + pt_word(i, 2));
        r.rho = __uint_as_float(__ldg(pts + pt_word(i, 1)));
        r.x = float(rec_x(pk));
        r.y = float(rec_y(pk));
        r.gr = gr;
        r.gu = rec_gx(gr);
        r.gv = rec_gy(gr);
        r.tmpl = float(rec_tmpl(pk));
    }
    return r;
}

template <bool kSkew>
__device__ __forceinline__ void add_outside(float sign, uint32_t gr, float a, float b, float rho, const Intrinsics& k, float* hs);

// One deferred candidate: the reference's own warp decides (lm_optimizer.rs:213-231).  Inside -> full evaluation into `acc`.
// Its side is kept in the level's far bitmap like that of the candidates that are outside for sure: when the exact verdict
// differs from the bit, the bit flips and -+J J^T goes to the thread's shared-memory accumulators (H = H_total - H_outside).
template <bool kSkew, bool kHuber>
__device__ __forceinline__ void eval_deferred(int slot, float ca, float cb, float rho, uint32_t gr, float tmpl, const LevelConst& lc,
                                              const Pose& model, Acc<kHuber>& acc, int& fixed, uint32_t* far_bitmap, float* hs, int& any_flip) {
    const Intrinsics k = lc.k;
    const int rows = int(lc.rows);
    const float gu = rec_gx(gr), gv = rec_gy(gr);
    // x, y are integers and a = fl(x - cx) is at most 2^-17 away from x - cx: rounding a + cx to the nearest integer restores them
    const float2 uv = warp_exact(model, k, rintf(ca + lc.cx), rintf(cb + lc.cy), rho);
    // 0 <= floor(u) < W-2  <=>  0 <= u < W-2 (W-2 is an integer); NaN compares false -> outside
    const bool inside = (uv.x >= 0.0f) && (uv.x < lc.wm2) && (uv.y >= 0.0f) && (uv.y < lc.hm2);
    if (!kHuber) {
        const unsigned bit = 1u << (slot & 31);
        const bool was_outside = (__ldcg(far_bitmap + (slot >> 5)) & bit) != 0u;
        if (was_outside == inside) {  // changed sides
            if (inside)
                atomicAnd(far_bitmap + (slot >> 5), ~bit);
            else
                atomicOr(far_bitmap + (slot >> 5), bit);
            add_outside<kSkew>(inside ? -1.0f : 1.0f, gr, ca, cb, rho, k, hs);
            any_flip = 1;
        }
    }
    if (inside) {
        const float fu = floorf(uv.x), fv = floorf(uv.y);
        const float a = uv.x - fu, b = uv.y - fv;
        const uint8_t* p = lc.img + (size_t(int(fu)) * size_t(rows) + size_t(int(fv)));
        const float t00 = float(__ldg(p)), t10 = float(__ldg(p + 1)), t01 = float(__ldg(p + rows)), t11 = float(__ldg(p + rows + 1));
        const float top = fmaf(a, t01 - t00, t00), bot = fmaf(a, t11 - t10, t10);
        const float val = fmaf(b, bot - top, top);  // same lerp form as `back`
        const float r = val - tmpl;
        ++fixed;
        if constexpr (kHuber) {
            float J[6];
            jacobian_centred<kSkew>(gu, gv, ca, cb, rho, k, J);
            accumulate_huber(acc, J, r, lc.huber_delta);
        } else {
            acc.e = fmaf(r, r, acc.e);
            if constexpr (kSkew) {
                float J[6];
                jacobian_centred<true>(gu, gv, ca, cb, rho, k, J);
#pragma unroll
                for (int q = 0; q < 6; ++q) acc.s[q] = fmaf(J[q], r, acc.s[q]);
            } else {
                accumulate_moments(acc, gu, gv, ca, cb, rho, r);
            }
        }
    }
}

// The boundary-band candidates of this warp's stages, after its hot loop.  Common case: they were stashed (at most kStashCap
// of them): one per lane, side by side.  Otherwise (e.g. a static camera: the x = 0 column and the y = 0 row sit exactly on
// the inside-test boundary) this warp's part of the near bitmap is scanned and the records are read back.
// `scratch`: this warp's ring memory (idle between passes), used as the compacted candidate list of the scan.
// `small_words` > 0: the level went through the small-level path (its `small_words` bitmap words dealt round-robin to the warps).
template <bool kSkew, bool kHuber, bool kTiled>
__device__ __noinline__ void deferred_pass(int warp, int lane, int first_stage, int stage_stride, int n_stages, int small_words,
                                           uint32_t* __restrict__ bitmap, uint32_t* __restrict__ far_bitmap, int n_near, const uint32_t* wlist,
                                           int n_flagged_words, uint32_t* scratch, float* hs, Acc<kHuber>* acc_io, int* n_fix, int* any_flip_io) {
    LmShared& S = lm_shared();
    const LevelConst& lc = s_lc;
    Acc<kHuber> acc = *acc_io;
    int fixed = 0, any_flip = 0;
    if (n_near <= kStashCap) {
        if (lane < n_near) {
            const uint32_t* e = S.stash[warp][lane];
            const int slot = int(e[0]);
            bitmap[slot >> 5] = 0u;  // leave the near bitmap all-zero for the next pass
            eval_deferred<kSkew, kHuber>(slot, __uint_as_float(e[1]), __uint_as_float(e[2]), __uint_as_float(e[3]), e[4], __uint_as_float(e[5]), lc,
                                         S.cand_model, acc, fixed, far_bitmap, hs, any_flip);
        }
    } else {
        // The flagged slots are first gathered into one list (`scratch`, at most kListCap entries between two flushes) and then
        // evaluated 32 at a time, whatever their spread over the bitmap words: a boundary ROW puts one slot into every word it
        // crosses, a boundary COLUMN 32 slots into one word.
        constexpr int kListCap = 1024;
        int n_list = 0;  // warp-uniform
        auto flush = [&]() {
            __syncwarp();
            for (int e = lane; e < n_list; e += 32) {
                const int slot = int(scratch[e]);
                const SlotRec sr = load_slot<kTiled>(lc.pts, lc.tiles_y, slot);
                eval_deferred<kSkew, kHuber>(slot, sr.x - lc.cx, sr.y - lc.cy, sr.rho, sr.gr, sr.tmpl, lc, S.cand_model, acc, fixed, far_bitmap, hs,
                                             any_flip);
            }
            __syncwarp();
            n_list = 0;
        };
        // every lane brings one flagged bitmap word: `mask` over the 32 slots starting at slot `base`
        auto round = [&](uint32_t mask, int base) {
            const int cnt = __popc(mask);
            int incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            if (n_list + total > kListCap) flush();
            int off = n_list + incl - cnt;
            while (mask) {
                scratch[off++] = uint32_t(base) + uint32_t(__ffs(mask) - 1);
                mask &= mask - 1u;
            }
            n_list += total;
        };
        if (n_flagged_words <= kNearWords) {
            // second tier (e.g. a whole boundary column exactly on the inside-test boundary: one word, 32 candidates): the flagged
            // words were remembered, one per lane
            uint32_t word = 0u, mask = 0u;
            if (lane < n_flagged_words) {
                word = wlist[2 * lane];
                mask = wlist[2 * lane + 1];
                bitmap[word] = 0u;
            }
            round(mask, int(word) * 32);
            flush();
        } else {
        // third tier: this warp's bitmap words, one per lane and round: the words of its stages, or its share of a small level's words
        constexpr int kW = kTiled ? kTileWordsBm : kStageWordsBm;
        const int my_words = small_words > 0 ? (small_words > first_stage ? (small_words - first_stage + stage_stride - 1) / stage_stride : 0)
                                             : (first_stage < n_stages ? ((n_stages - first_stage + stage_stride - 1) / stage_stride) * kW : 0);
        for (int k0 = 0; k0 < my_words; k0 += 32) {  // warp-uniform trip count
            const int kk = k0 + lane;
            int word = 0;
            uint32_t mask = 0u;
            if (kk < my_words) {
                word = small_words > 0 ? first_stage + kk * stage_stride : (first_stage + (kk / kW) * stage_stride) * kW + kk % kW;
                mask = bitmap[word];
                if (mask) bitmap[word] = 0u;  // leave the near bitmap all-zero for the next pass
            }
            if (!__any_sync(0xffffffffu, mask != 0u)) continue;
            round(mask, word * 32);
        }
        flush();
        }
        // `scratch` is ring memory: order these generic-proxy accesses before the next pass's bulk copies (async proxy)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    *acc_io = acc;
    *n_fix = fixed;
    *any_flip_io |= __any_sync(0xffffffffu, any_flip != 0) ? 1 : 0;
}

// Per-level constants the common path keeps in registers.
struct PassConst {
    float cx, cy, hu, hv, lo_u, lo_v, magic_u, magic_v, zero_u, zero_v;
    uint32_t rows;
    const uint8_t* img_biased;
    float tex_ku, tex_kv;
    unsigned long long tex;
};

// floor / fraction of a warped coordinate and the texel fetch of its 2x2 footprint (common tail of every front).
// (`ku`, `kv`: the texture offsets lc.tex_ku / lc.tex_kv, or a lane's own copies of them: see TileLane)
__device__ __forceinline__ void sample_front(float u, float v, const PassConst& lc, Front& f, float ku, float kv) {
#if VORS_TEX
#if VORS_FRND
    const float fu = floorf(u), fv = floorf(v);
    f.fa = u - fu;
    f.fb = v - fv;
    tex_gather(lc.tex, fv + (8388608.0f - kv), fu + (8388608.0f - ku), f.t00, f.t10, f.t01, f.t11);
#else
    // tu = 2^23 + floor(u) exactly (round-down add; 0 <= u < 2^22 on this path)
    const float tu = __fadd_rd(u, 8388608.0f), tv = __fadd_rd(v, 8388608.0f);
    f.fa = u - (tu - 8388608.0f);
    f.fb = v - (tv - 8388608.0f);
    tex_gather(lc.tex, tv - kv, tu - ku, f.t00, f.t10, f.t01, f.t11);
#endif
#else
    // floor and fraction without F2I / I2F (see kMagicBits)
    const float tu = __fadd_rd(u, lc.magic_u), tv = __fadd_rd(v, lc.magic_v);
    f.fa = u - (tu - lc.magic_u);
    f.fb = v - (tv - lc.magic_v);
    const uint8_t* p = lc.img_biased + (__float_as_uint(tu) * lc.rows + __float_as_uint(tv));
    f.t00 = __ldg(p);
    f.t10 = __ldg(p + 1);
    f.t01 = __ldg(p + lc.rows);
    f.t11 = __ldg(p + lc.rows + 1);
#endif
}
__device__ __forceinline__ void sample_front(float u, float v, const PassConst& lc, Front& f) { sample_front(u, v, lc, f, lc.tex_ku, lc.tex_kv); }

// bilinear sample in lerp form minus the template value: the same interpolant as lm_optimizer.rs:241-246 (a along x, b along
// y) with 6 instead of 10 operations; it differs from the reference's four-product expression by ~1 ulp of the value, far
// inside the 1e-5 bar on the pass energy (tests/test_gpu_parity.py).  With normalised u8 texels the sample comes back in
// units of 255 grey levels and is rescaled by the one FFMA that subtracts the template.
__device__ __forceinline__ float residual_of(const Front& f, float tmpl) {
    const float a = f.fa, b = f.fb;
#if VORS_TEX
    const float top = fmaf(a, f.t01 - f.t00, f.t00), bot = fmaf(a, f.t11 - f.t10, f.t10);
    const float val = fmaf(b, bot - top, top);
    return kTexScale == 1.0f ? val - tmpl : fmaf(kTexScale, val, -tmpl);
#else
    const float t00 = u2f(f.t00), t10 = u2f(f.t10);
    const float top = fmaf(a, u2f(f.t01) - t00, t00), bot = fmaf(a, u2f(f.t11) - t10, t10);
    const float val = fmaf(b, bot - top, top);
    return val - tmpl;
#endif
}

// Warp-uniform bookkeeping of the slots of a pass that the hot loop did not evaluate.
struct Defer {
    uint32_t* near;    // this level's bitmap of slots for deferred_pass (global; all zero between passes)
    uint32_t* far;     // this level's bitmap of the slots that were outside for sure in the previous pass of the level
    int n_bad;         // slots of this pass redirected to the zero page so far
    int n_near;        // of which flagged in `near`
    uint32_t* stash;   // this warp's stash of the flagged candidates (shared memory, kStashCap entries of 6 words)
    int n_words;       // bitmap words flagged in `near` during this pass
    uint32_t* wlist;   // (word, mask) of the first kNearWords of them (shared memory)
    int any_flip;      // this warp added to its hsm accumulators during this pass
    int first_pass;    // first pass of the level: the far bitmap holds nothing yet
};

// +-J J^T of this lane's candidate (`sign` = +1 / -1 / 0) added to the thread's shared-memory accumulators.
template <bool kSkew>
__device__ __forceinline__ void add_outside(float sign, uint32_t gr, float a, float b, float rho, const Intrinsics& k, float* hs) {
    float J[6];
    // J is linear in the gradient: zero gradient (and a finite inverse depth: padding slots carry NaN) -> J = 0 exactly
    const bool on = sign != 0.0f;
    jacobian_centred<kSkew>(on ? rec_gx(gr) : 0.0f, on ? rec_gy(gr) : 0.0f, a, b, on ? rho : 0.0f, k, J);
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        const float jc = sign * J[c];
#pragma unroll
        for (int d = c; d < 6; ++d) hs[tri(c, d)] = fmaf(jc, J[d], hs[tri(c, d)]);
    }
}

// front, first half: unpack, warp, inside test - straight-line arithmetic that the compiler can interleave with the
// `back` of the previous candidate (issued between the two halves).
struct FrontA {
    uint32_t pk, gr;
    float rho, a, b, u, v;
    bool ok;  // inside for sure (false for NaN)
};
__device__ __forceinline__ FrontA front_a(uint32_t pk, float rho, uint32_t gr, const float (&M)[12], const PassConst& lc) {
#if VORS_X_MAGIC
    const float x = __uint_as_float((pk & 0xFFFu) | 0x4B000000u) - 8388608.0f;
#else
    const float x = float(pk & 0xFFFu);
#endif
    const float y = float((pk >> 12) & 0xFFFu);
    const float a = x - lc.cx, b = y - lc.cy;  // camera.rs:135-140 starts from these rounded differences too
    const float U = fmaf(M[0], a, fmaf(M[1], b, fmaf(M[3], rho, M[2])));
    const float V = fmaf(M[4], a, fmaf(M[5], b, fmaf(M[7], rho, M[6])));
    const float W = fmaf(M[8], a, fmaf(M[9], b, fmaf(M[11], rho, M[10])));
    const float iw = rcp_approx(W);
    FrontA o;
    o.u = fmaf(U, iw, lc.cx);
    o.v = fmaf(V, iw, lc.cy);
    // lm_optimizer.rs:231: inside iff 0 <= floor(u) < W-2 and 0 <= floor(v) < H-2; here: inside with a margin of kBandPx
    o.ok = (fabsf(o.u - lc.hu) < lc.lo_u) & (fabsf(o.v - lc.hv) < lc.lo_v);
    o.pk = pk;
    o.gr = gr;
    o.rho = rho;
    o.a = a;
    o.b = b;
    return o;
}

// The rare, warp-uniform part of a front: bookkeeping of the slots of this 32-slot word that the common path cannot evaluate.
// `not_ok`: ballot of the lanes not inside for sure; `old_far`: the word's far-bitmap word from the previous pass of the level.
template <bool kSkew, bool kHuber>
__device__ __forceinline__ void rare_slots(unsigned not_ok, float u, float v, float rho, uint32_t gr, float tmpl, float a, float b, int word,
                                           unsigned old_far, const PassConst& lc, const Intrinsics& k, Defer& df, float* hs, int lane) {
    // H over the inside set = H_total - H_outside (lm_optimizer.rs:100 sums J J^T over the inside set).  H_outside is
    // maintained incrementally across the passes of a level: only candidates that were outside for sure in the previous
    // pass and are not now, or the reverse, add -+J J^T (after the first passes the pose barely moves: few flips).
    const bool far = (fabsf(u - lc.hu) > s_lc.hi_u) | (fabsf(v - lc.hv) > s_lc.hi_v);  // outside for sure (false for NaN)
    // padding slots and pixels without depth carry a NaN inverse depth: never `far`, and they need no second look
    const unsigned live_mask = __ballot_sync(0xffffffffu, rho == rho);
    const unsigned far_mask = __ballot_sync(0xffffffffu, far), near_mask = not_ok & ~far_mask & live_mask;
    // candidates in the boundary band keep the side they were on until their exact evaluation decides (deferred_pass): a
    // candidate sitting ON the boundary (static camera) then costs nothing here from its second pass on
    const unsigned new_far = (far_mask & ~near_mask) | (old_far & near_mask);
    const unsigned flips = kHuber ? 0u : new_far ^ old_far;  // (Huber weights: H is summed directly, nothing to maintain)
    if (flips) {
        const float sign = ((flips >> lane) & 1u) ? (far ? 1.0f : -1.0f) : 0.0f;
        add_outside<kSkew>(sign, gr, a, b, rho, k, hs);
        df.any_flip = 1;
    }
    if (lane == 0 && flips) df.far[word] = new_far;
    if (near_mask) {  // boundary band / NaN coordinates of a live candidate: see deferred_pass
        if (lane == 0) {
            df.near[word] = near_mask;
            if (df.n_words < kNearWords) {
                df.wlist[2 * df.n_words] = uint32_t(word);
                df.wlist[2 * df.n_words + 1] = near_mask;
            }
        }
        df.n_words += 1;
        const int at = df.n_near + __popc(near_mask & ((1u << lane) - 1u));
        if (((near_mask >> lane) & 1u) && at < kStashCap) {
            uint32_t* e = df.stash + 6 * at;
            e[0] = uint32_t(word) * 32u + uint32_t(lane);
            e[1] = __float_as_uint(a);
            e[2] = __float_as_uint(b);
            e[3] = __float_as_uint(rho);
            e[4] = gr;
            e[5] = __float_as_uint(tmpl);
        }
    }
    df.n_bad += __popc(not_ok);
    df.n_near += __popc(near_mask);
}

// front, second half: the rare paths (warp-uniform branch), floor / fraction, texel gathers.
// `word` is the bitmap word of this call's 32 slots.
// `old_words`: lane j holds the far-bitmap word of the stage's word j from the previous pass of the level; bit j of
// `old_nz` says whether it is non-zero.  `j` = this call's word within the stage.
template <bool kSkew, bool kHuber>
__device__ __forceinline__ void front_b(const FrontA& x, int word, int j, unsigned old_words, unsigned old_nz, const PassConst& lc,
                                        const Intrinsics& k, Defer& df, float* hs, int lane, Front& f) {
    uint32_t pk = x.pk, gr = x.gr;
    float rho = x.rho, u = x.u, v = x.v;
    const bool ok = x.ok;
    const unsigned not_ok = __ballot_sync(0xffffffffu, !ok);
    if (not_ok | (old_nz & (1u << j))) {  // warp-uniform, rare
        const unsigned old_far = __shfl_sync(0xffffffffu, old_words, j);
        rare_slots<kSkew, kHuber>(not_ok, u, v, rho, gr, u2f(pk >> 24), x.a, x.b, word, old_far, lc, k, df, hs, lane);
        if (!ok) {
            u = lc.zero_u;
            v = lc.zero_v;
            pk = 0u;   // template 0: r = 0 - 0
            rho = 0.0f;
            gr = 0u;   // zero gradient: J = 0 (only the Huber path forms J on the common path)
        }
    }
    sample_front(u, v, lc, f);
    f.pk = pk;
    f.gr = gr;
    f.a = x.a;
    f.b = x.b;
    f.rho = rho;
}

// ---- tiled dense records: implicit coordinates -------------------------------------------------------------------------
// Per-lane constants of one tile (= ring stage): the lane's row, and the part of the folded warp that does not depend on the
// word (column) or on the inverse depth: [Ub Vb Wb] = M[:, 0] a0 + M[:, 1] b + M[:, 2] with a0 = x0 - cx of the tile's first
// column and b = y - cy of the lane's row.  Word j then costs two FFMAs per row of the matrix: M[:, 0] j + . and M[:, 3] rho + .
// Rows of a tile below the image (levels whose height is not a multiple of 32) are DEAD lanes: their records hold a zero
// inverse depth, gradient and template, and they get constants of their own - [Ub Vb Wb] = 2^20 [u* - cx, v* - cy, 1] with
// (u*, v*) the middle of the image, so that they pass the inside test in every word whatever the model (the word-dependent
// terms are 2^-20 of that), and texture offsets (ku, kv) that send their gather to the zero page: residual 0 - 0, zero
// gradient, no trip through the rare path.  They are subtracted from the slot count instead.
struct TileLane {
    float a0, b, Ub, Vb, Wb;
    float M0v, M4v, M8v;
    float ku, kv;
};
// front of word j of a tile (tiled records only exist for zero skew, plain L2: see AlignParams::tiled).
// `tmpl`: the slot's template value; travels in Front::pk as f32 bits.
// J: the word within its half tile (compile time); `t` holds the half tile's constants (a0, Ub, Vb, Wb of its first column),
// `word` its first bitmap word, `jh` its first word within the tile, `old_nz` the tile's non-zero far words shifted so
// that bit J is this word's.
template <int J>
__device__ __forceinline__ void tile_front(const TileLane& t, float rho, uint32_t gr, float tmpl, const float (&M)[12], int word, int jh,
                                           unsigned old_words, unsigned old_nz, const PassConst& lc, const Intrinsics& k, Defer& df,
                                           float* hs, int lane, Front& f) {
    // (M0v, M4v, M8v: the first matrix column in ordinary registers - an FFMA takes one uniform-register or immediate operand,
    // and the word index J is the immediate)
    const float U = fmaf(M[3], rho, J ? fmaf(t.M0v, float(J), t.Ub) : t.Ub);
    const float V = fmaf(M[7], rho, J ? fmaf(t.M4v, float(J), t.Vb) : t.Vb);
    const float W = fmaf(M[11], rho, J ? fmaf(t.M8v, float(J), t.Wb) : t.Wb);
    const float iw = rcp_approx(W);
    float u = fmaf(U, iw, lc.cx), v = fmaf(V, iw, lc.cy);
    const float a = t.a0 + float(J);
    // lm_optimizer.rs:231: inside iff 0 <= floor(u) < W-2 and 0 <= floor(v) < H-2; here: inside with a margin of kBandPx
    const bool ok = (fabsf(u - lc.hu) < lc.lo_u) & (fabsf(v - lc.hv) < lc.lo_v);
    const unsigned not_ok = __ballot_sync(0xffffffffu, !ok);
    if (not_ok | (old_nz & (1u << J))) {  // warp-uniform, rare
        const unsigned old_far = __shfl_sync(0xffffffffu, old_words, jh + J);
        rare_slots<false, false>(not_ok, u, v, rho, gr, tmpl, a, t.b, word + J, old_far, lc, k, df, hs, lane);
        if (!ok) {
            u = lc.zero_u;
            v = lc.zero_v;
            tmpl = 0.0f;
            rho = 0.0f;
            gr = 0u;
        }
    }
    sample_front(u, v, lc, f, t.ku, t.kv);
    f.pk = __float_as_uint(tmpl);
    f.gr = gr;
    f.a = a;
    f.b = t.b;
    f.rho = rho;
}
__device__ __forceinline__ void tile_back(const Front& f, Acc<false>& acc) {
    const float r = residual_of(f, __uint_as_float(f.pk));
    acc.e = fmaf(r, r, acc.e);
    accumulate_moments(acc, rec_gx(f.gr), rec_gy(f.gr), f.a, f.b, f.rho, r);
}
__device__ __forceinline__ float half_lo(uint32_t w) { return __half2float(__ushort_as_half((unsigned short)(w & 0xFFFFu))); }
__device__ __forceinline__ float half_hi(uint32_t w) { return __half2float(__ushort_as_half((unsigned short)(w >> 16))); }

// back: bilinear sample, residual, accumulate.
template <bool kSkew, bool kHuber>
__device__ __forceinline__ void back(const Front& f, const Intrinsics& k, float huber_delta, Acc<kHuber>& acc) {
    const float gu = rec_gx(f.gr), gv = rec_gy(f.gr);
    const float r = residual_of(f, u2f(f.pk >> 24));
    if constexpr (kHuber) {
        float J[6];
        jacobian_centred<kSkew>(gu, gv, f.a, f.b, f.rho, k, J);
        accumulate_huber(acc, J, r, huber_delta);
    } else {
        acc.e = fmaf(r, r, acc.e);
        if constexpr (kSkew) {
            float J[6];
            jacobian_centred<true>(gu, gv, f.a, f.b, f.rho, k, J);
#pragma unroll
            for (int c = 0; c < 6; ++c) acc.s[c] = fmaf(J[c], r, acc.s[c]);
        } else {
            accumulate_moments(acc, gu, gv, f.a, f.b, f.rho, r);
        }
    }
}

// The serial part of one LM round, run by thread 0 of every CTA of the team on identical inputs.
// Mirrors init / eval / stop_criterion / step of lm_optimizer.rs:113-192.
__device__ __noinline__ void lm_decide(LmShared& S, const AlignParams& P, const AlignJob& job, int job_idx, int lvl, bool writer) {
    const double* tot = S.tot;
    const int n_inside = int(tot[1]);
    // energy = energy_sum / residuals.len() as f32 (lm_optimizer.rs:85); 0/0 = NaN when nothing is inside
    const float E = float(tot[0]) / float(n_inside);
    S.n_passes += 1;
    bool stop;
    bool accepted = true;
    const float lam_used = S.init_phase ? P.lm_coef_init : S.lam;
    if (S.init_phase) {
        S.lam = P.lm_coef_init;
        S.iter = 0;
        S.init_phase = 0;
        stop = false;
    } else {
        const bool rejected = E > S.keptE;  // lm_optimizer.rs:144 (NaN compares false -> accepted)
        accepted = !rejected;
        const bool too_many = P.fixed_iters ? (S.iter >= P.fixed_iters) : (S.iter > P.max_iters);
        if (rejected) {
            stop = too_many;
            if (!too_many) S.lam *= P.lm_coef_reject_mult;
        } else if (too_many) {
            stop = true;
        } else {
            const float d_energy = S.keptE - E;
            stop = P.fixed_iters ? false : !(d_energy > P.energy_delta_stop);
            S.lam = P.lm_coef_accept_mult * S.lam;
        }
    }
    if (writer && P.trace && S.trace_len < kTraceCap) {
        vors_trace_rec& t = P.trace[size_t(job_idx) * kTraceCap + S.trace_len];
        t.level = lvl;
        t.iter = S.iter;
        t.energy = E;
        t.n_inside = n_inside;
        t.lm_coef = lam_used;
        t.accepted = accepted ? 1 : 0;
    }
    S.trace_len += 1;
    if (accepted) {
        S.keptE = E;
#pragma unroll
        for (int c = 0; c < 6; ++c) S.keptg[c] = float(tot[2 + c]);
#pragma unroll
        for (int c = 0; c < 21; ++c) S.keptH[c] = float(tot[8 + c]);
        S.kept_model = S.cand_model;
    }
    if (stop) {
        S.cont = 0;
        return;
    }
    // step (lm_optimizer.rs:123-136)
    S.iter += 1;
    float A[36], b[6];
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c < 6; ++c) A[r * 6 + c] = S.keptH[r <= c ? tri(r, c) : tri(c, r)];
#pragma unroll
    for (int c = 0; c < 6; ++c) A[c * 6 + c] *= 1.0f + S.lam;
#pragma unroll
    for (int c = 0; c < 6; ++c) b[c] = S.keptg[c];
    if (!cholesky6_solve(A, b)) {
        S.failed = 1;
        S.cont = 0;
        return;
    }
    const Pose delta = se3_exp(b);
    S.cand_model = pose_renormalize(pose_mul(S.kept_model, pose_inverse(delta)));
    S.cont = 1;  // the caller publishes the candidate model's warp matrix
}

// Entry t of the finished pass (sum r^2, n_inside, g[6], H[21]) from the raw totals, f64; one lane of warp 0 per entry.
template <bool kSkew, bool kHuber>
__device__ __forceinline__ double finish_entry(int t, const double* raw, const LevelConst& lc, const double* h_total) {
    if (kHuber) return raw[t];  // sum rho, n_inside, g[6], H[21] were summed directly
    if (t < 2) return raw[t];
    if (t >= 8) {
        // H over the inside set = H_total (all candidates, per keyframe level) - H_outside; an empty inside set must give an
        // exactly zero H (the reference then fails its Cholesky, lm_optimizer.rs:131-133)
        return raw[1] > 0.0 ? h_total[t - 8] - raw[kRawH + (t - 8)] : 0.0;
    }
    if (kSkew) return raw[t];
    const double* m = raw + 2;
    const double fu = lc.k.fx, fv = lc.k.fy;
    switch (t) {
        case 2: return fu * m[kSrp];
        case 3: return fv * m[kSrq];
        case 4: return -m[kSrs];
        case 5: return -m[kSbs] * lc.inv_fy - fv * m[kSq];
        case 6: return m[kSas] * lc.inv_fx + fu * m[kSp];
        default: return (fv * lc.inv_fx) * m[kSaq] - (fu * lc.inv_fy) * m[kSbp];
    }
}

// Entry e = 4 r + c of the centred warp matrix (lie.cuh `warp_matrix(..., centred = true)`), one lane of warp 0 per entry:
// rows of P [R Ki | t] with P = [[fx, s, 0], [0, fy, 0], [0, 0, 1]], Ki = [[1/fx, -s/(fx fy), 0], [0, 1/fy, 0], [0, 0, 1]].
__device__ __forceinline__ float warp_matrix_entry(int e, const Pose& m, const LevelConst& lc) {
    const int r = e >> 2, c = e & 3;
    const double qi = m.q.i, qj = m.q.j, qk = m.q.k, qw = m.q.w;
    const double fx = lc.k.fx, fy = lc.k.fy, s = lc.k.s;
    // column c of A = [R Ki | t] (rotation matrix of the possibly slightly non-unit quaternion exactly as quat_rotate applies it)
    double a0, a1, a2;
    if (c == 3) {
        a0 = m.t.x; a1 = m.t.y; a2 = m.t.z;
    } else {
        const double r00 = 1.0 - 2.0 * (qj * qj + qk * qk), r10 = 2.0 * (qi * qj + qk * qw), r20 = 2.0 * (qi * qk - qj * qw);
        const double r01 = 2.0 * (qi * qj - qk * qw), r11 = 1.0 - 2.0 * (qi * qi + qk * qk), r21 = 2.0 * (qj * qk + qi * qw);
        const double r02 = 2.0 * (qi * qk + qj * qw), r12 = 2.0 * (qj * qk - qi * qw), r22 = 1.0 - 2.0 * (qi * qi + qj * qj);
        if (c == 0) {
            const double ki = lc.inv_fx;
            a0 = r00 * ki; a1 = r10 * ki; a2 = r20 * ki;
        } else if (c == 1) {
            const double k01 = lc.k01, k11 = lc.inv_fy;
            a0 = r00 * k01 + r01 * k11; a1 = r10 * k01 + r11 * k11; a2 = r20 * k01 + r21 * k11;
        } else {
            a0 = r02; a1 = r12; a2 = r22;
        }
    }
    return float(r == 0 ? fx * a0 + s * a1 : r == 1 ? fy * a1 : a2);
}

// ---- the serial part of an LM round, warp-uniform --------------------------------------------------------------------------
// Every lane of the serial warp runs the same scalar program on the same values (no divergent region, no round trip through
// shared memory between its stages, no call): the chain is latency-bound, so what counts is its length in dependent
// instructions.  Divisions and square roots are a MUFU seed plus Newton steps (<= 1 ulp) instead of the IEEE sequences with
// their slow-path checks: the solver's result is not bit-compatible with nalgebra's anyway (FMA contraction), and its
// round-off is five orders of magnitude below the step it computes.
__device__ __forceinline__ float fast_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return fmaf(r, fmaf(-x, r, 1.0f), r);
}
__device__ __forceinline__ float fast_rsqrt(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r * fmaf(-0.5f * x, r * r, 1.5f);  // one Newton step
}
// nalgebra 0.17 `Matrix6::cholesky()` + `solve` (lm_optimizer.rs:131-134) in the operation order of lie.cuh's
// cholesky6_solve, with the fast reciprocals above.  A: lower triangle, row-major 6x6.
__device__ __forceinline__ bool cholesky6_solve_fast(float (&A)[36], float (&b)[6]) {
#pragma unroll
    for (int j = 0; j < 6; ++j) {
#pragma unroll
        for (int k = 0; k < j; ++k) {
            const float factor = -A[j * 6 + k];
#pragma unroll
            for (int i = j; i < 6; ++i) A[i * 6 + j] = fmaf(factor, A[i * 6 + k], A[i * 6 + j]);
        }
        const float diag = A[j * 6 + j];
        if (!(diag > 0.0f)) return false;
        const float inv = fast_rsqrt(diag);
        A[j * 6 + j] = inv;  // 1 / L_jj: every later use of the diagonal is a division by it
#pragma unroll
        for (int i = j + 1; i < 6; ++i) A[i * 6 + j] *= inv;
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const float coeff = b[i] * A[i * 6 + i];
        b[i] = coeff;
#pragma unroll
        for (int r = i + 1; r < 6; ++r) b[r] = fmaf(-coeff, A[r * 6 + i], b[r]);
    }
#pragma unroll
    for (int i = 5; i >= 0; --i) {
        float dot = 0.0f;
#pragma unroll
        for (int r = i + 1; r < 6; ++r) dot = fmaf(A[r * 6 + i], b[r], dot);
        b[i] = (b[i] - dot) * A[i * 6 + i];
    }
    return true;
}
// src/math/se3.rs:65-95 `exp` (lie.cuh se3_exp) with the fast reciprocals; the trigonometric branch (theta >= 0.01, the first
// rounds of a coarse level at most) keeps the accurate sinf / cosf: (1 - cos theta) / theta^2 cancels.
__device__ __forceinline__ Pose se3_exp_fast(const float (&xi)[6]) {
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
    const float w11 = wx * wx, w12 = wx * wy, w13 = wx * wz, w22 = wy * wy, w23 = wy * wz, w33 = wz * wz;
    Pose out;
    // V = I + c1 Omega + c2 Omega^2 applied to v
    out.t.x = ((1.0f + c2 * (-w22 - w33)) * v0 + (c1 * -wz + c2 * w12) * v1) + (c1 * wy + c2 * w13) * v2;
    out.t.y = ((c1 * wz + c2 * w12) * v0 + (1.0f + c2 * (-w11 - w33)) * v1) + (c1 * -wx + c2 * w23) * v2;
    out.t.z = ((c1 * -wy + c2 * w13) * v0 + (c1 * wx + c2 * w23) * v1) + (1.0f + c2 * (-w11 - w22)) * v2;
    const Quat q{imag_factor * wx, imag_factor * wy, imag_factor * wz, real_factor};
    const float inv_n = fast_rsqrt(quat_norm2(q));  // UnitQuaternion::from_quaternion normalises
    out.q = {q.i * inv_n, q.j * inv_n, q.k * inv_n, q.w * inv_n};
    return out;
}

// g entries of the finished pass from the gradient moments: g[t] = ga[t] * m[gi[t]] + gb[t] * m[gj[t]] (see accumulate_moments;
// inverse_compositional.rs:326-340 with zero skew).  The coefficient pairs are per-level constants (LevelConst::ga, gb).
__constant__ int c_gi[6] = {kSrp, kSrq, kSrs, kSbs, kSas, kSaq};
__constant__ int c_gj[6] = {kSrp, kSrq, kSrs, kSq, kSp, kSbp};

// One LM round after a pass: finish (sum r^2, n_inside, g, H from the raw totals), decide (init / eval / stop_criterion,
// lm_optimizer.rs:113-192), step (:123-136) and the next candidate model's warp matrix.  Run by all 32 lanes of the serial warp.
template <bool kSkew, bool kHuber>
__device__ __noinline__ void serial_round(const AlignParams& P, int job_idx, int lvl, bool writer, int lane, int n_level, int pass_only) {
    LmShared& S = lm_shared();
    const LevelConst& lc = s_lc;
    const double* raw = S.raw;
    // ---- finish: lane t holds entry t of (sum r^2, n_inside, g[6], H[21])
    double mine = 0.0;
    if (lane < kNumAcc) {
        if (kHuber || kSkew || lane < 2) {
            mine = raw[lane];
        } else if (lane < 8) {
            const int t = lane - 2;
            mine = lc.ga[t] * raw[2 + c_gi[t]] + lc.gb[t] * raw[2 + c_gj[t]];
        }
        if (!kHuber && lane >= 8)
            // H over the inside set = H_total (all candidates, per keyframe level) - H_outside; an empty inside set must give an
            // exactly zero H (the reference then fails its Cholesky, lm_optimizer.rs:131-133)
            mine = raw[1] > 0.0 ? S.h_total[lane - 8] - raw[kRawH + (lane - 8)] : 0.0;
        S.tot[lane] = mine;
    }
    if (lane == 0) S.point_passes += (unsigned long long)n_level;
    if (pass_only) {
        if (lane == 0) {
            S.n_passes += 1;
            S.cont = 0;
        }
        return;
    }
    // ---- decide (every lane, same values)
    const float mine_f = float(mine);
    const int n_inside = int(__shfl_sync(0xffffffffu, mine, 1));
    // energy = energy_sum / residuals.len() as f32 (lm_optimizer.rs:85); 0/0 = NaN when nothing is inside
    const float E = __shfl_sync(0xffffffffu, mine_f, 0) / float(n_inside);
    const int init_phase = S.init_phase, iter = S.iter;
    const float lam_old = S.lam, keptE = S.keptE;
    const float lam_used = init_phase ? P.lm_coef_init : lam_old;
    float lam = lam_old;
    bool stop, accepted = true;
    if (init_phase) {
        lam = P.lm_coef_init;
        stop = false;
    } else {
        const bool rejected = E > keptE;  // lm_optimizer.rs:144 (NaN compares false -> accepted)
        accepted = !rejected;
        const bool too_many = P.fixed_iters ? (iter >= P.fixed_iters) : (iter > P.max_iters);
        if (rejected) {
            stop = too_many;
            if (!too_many) lam = lam_old * P.lm_coef_reject_mult;
        } else if (too_many) {
            stop = true;
        } else {
            const float d_energy = keptE - E;
            stop = P.fixed_iters ? false : !(d_energy > P.energy_delta_stop);
            lam = P.lm_coef_accept_mult * lam_old;
        }
    }
    const int trace_len = S.trace_len;
    if (lane == 0 && writer && P.trace && trace_len < kTraceCap) {
        vors_trace_rec& t = P.trace[size_t(job_idx) * kTraceCap + trace_len];
        t.level = lvl;
        t.iter = init_phase ? 0 : iter;
        t.energy = E;
        t.n_inside = n_inside;
        t.lm_coef = lam_used;
        t.accepted = accepted ? 1 : 0;
    }
    // kept state = last accepted evaluation; every lane keeps g, H and the kept model in registers for the step below
    float g[6], H[21];
#pragma unroll
    for (int c = 0; c < 6; ++c) g[c] = accepted ? __shfl_sync(0xffffffffu, mine_f, 2 + c) : S.keptg[c];
#pragma unroll
    for (int c = 0; c < 21; ++c) H[c] = accepted ? __shfl_sync(0xffffffffu, mine_f, 8 + c) : S.keptH[c];
    const Pose kept = accepted ? S.cand_model : S.kept_model;
    // every lane has read the round's state: only now may it be overwritten (the lanes of a warp are not in lock step)
    __syncwarp();
    if (accepted) {
        if (lane >= 2 && lane < 8) S.keptg[lane - 2] = mine_f;
        if (lane >= 8 && lane < kNumAcc) S.keptH[lane - 8] = mine_f;
        if (lane == 0) {
            S.keptE = E;
            S.kept_model = kept;
        }
    }
    if (lane == 0) {
        S.n_passes += 1;
        S.trace_len = trace_len + 1;
        S.lam = lam;
        S.init_phase = 0;
        S.iter = init_phase ? 0 : iter;
    }
    if (stop) {
        if (lane == 0) S.cont = 0;
        return;
    }
    // ---- step (lm_optimizer.rs:123-136)
    float A[36], b[6];
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c <= r; ++c) A[r * 6 + c] = H[tri(c, r)];
#pragma unroll
    for (int c = 0; c < 6; ++c) A[c * 6 + c] *= 1.0f + lam;
#pragma unroll
    for (int c = 0; c < 6; ++c) b[c] = g[c];
    if (!cholesky6_solve_fast(A, b)) {
        if (lane == 0) {
            S.iter = (init_phase ? 0 : iter) + 1;
            S.failed = 1;
            S.cont = 0;
        }
        return;
    }
    const Pose cand = pose_renormalize(pose_mul(kept, pose_inverse(se3_exp_fast(b))));
    if (lane == 0) {
        S.iter = (init_phase ? 0 : iter) + 1;
        S.cand_model = cand;
        S.cont = 1;
    }
    if (lane < 12) S.M[lane] = warp_matrix_entry(lane, cand, lc);
}

template <bool kSkew, bool kHuber, bool kTiled>
#ifdef VORS_MAXREG
__global__ void __maxnreg__(VORS_MAXREG) k_align(const AlignParams P) {
#else
__global__ void __launch_bounds__(kBlock, kMinCtasPerSm) k_align(const AlignParams P) {
#endif
    LmShared& S = lm_shared();
    const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int team = P.team;
    const int team_id = blockIdx.x / team, rank = blockIdx.x - team_id * team;
    const int n_teams = gridDim.x / team;
    TeamScratch* scratch = team > 1 ? P.scratch + team_id : nullptr;
    unsigned epoch = 0;      // passes this team has synchronised on so far (same in every CTA of the team)
    // this warp's TMA ring: slot of the next stage to consume and the parity its full barrier completes with
    uint32_t ring_slot = 0, ring_parity = 0;
    const uint32_t ring_base = smem_u32(&S.ring[warp][0]), bar_base = smem_u32(&S.full_bar[warp][0]);
    const uint64_t l2_policy = l2_evict_first_policy();
    if (tid == 0) {
        for (int w = 0; w < kWarps; ++w)
            for (int st = 0; st < kStages; ++st) mbar_init(smem_u32(&S.full_bar[w][st]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < kWarps * 21) (&S.hout[0][0])[tid] = 0.0;
#if VORS_TIMING
    if (tid < 96) (&S.dbg[0][0])[tid] = 0;
#endif
    for (int i = tid; i < kConsumers * kHsmStride; i += kBlock) (&S.hsm[0][0])[i] = 0.0f;
    __syncthreads();

    // Jobs: one per team when the batch fits the device; a larger batch is handed out dynamically (team == 1 only), so that a
    // CTA that finishes a cheap alignment early takes the next one instead of idling behind the slowest: the spread of the
    // per-alignment times (border candidates, boundary-band work) otherwise costs ~10 % at the end of every launch.
    const bool dynamic_jobs = P.job_counter != nullptr && team == 1 && P.n_jobs > n_teams;
    __shared__ int s_next_job;
    for (int job_idx = team_id;;) {
        if (job_idx >= P.n_jobs) break;
        const AlignJob& job = P.jobs[job_idx];
        const bool writer = (rank == 0);
        if (writer && tid == 0) P.results[job_idx].t_begin_ns = global_timer_ns();
        if (tid == 0) {
            S.out_model = P.init[job_idx];
            S.failed = 0;
            S.n_passes = 0;
            S.trace_len = 0;
            S.point_passes = 0ull;
        }
        __syncthreads();

        const int lvl_first = __shfl_sync(0xffffffffu, job.lvl_first, 0), lvl_last = __shfl_sync(0xffffffffu, job.lvl_last, 0);
        for (int lvl = lvl_first; lvl >= lvl_last; --lvl) {
            const LevelJob& lj = job.lv[lvl];
            const int n = __shfl_sync(0xffffffffu, *lj.n_ptr, 0);
            const uint32_t* __restrict__ pts = lj.pts;
            const int n_stages = kTiled ? __shfl_sync(0xffffffffu, lj.n_tiles, 0) : (n + kStageCand - 1) / kStageCand;
            const int TW = team * kWarps;
            if (tid < kWarps * 21) (&S.hout[0][0])[tid] = 0.0;
            if (tid >= 32 && tid < 32 + 21) S.h_total[tid - 32] = lj.h_total[tid - 32];
            if (tid == 0) {
                const Intrinsics k = lj.k;
                const int wm2i = lj.cols - 2, hm2i = lj.rows - 2;
                LevelConst c;
                c.cx = k.cx;
                c.cy = k.cy;
                c.hu = 0.5f * float(wm2i);
                c.hv = 0.5f * float(hm2i);
                c.lo_u = c.hu - kBandPx;
                c.lo_v = c.hv - kBandPx;
                c.hi_u = c.hu + kBandPx;
                c.hi_v = c.hv + kBandPx;
                c.wm2 = float(wm2i);
                c.hm2 = float(hm2i);
                c.zero_u = lj.zero_u;
                c.zero_v = lj.zero_v;
                c.rows = uint32_t(lj.rows);
                c.n = n;
                c.tex_ku = lj.tex_ku;
                c.tex_kv = lj.tex_kv;
                c.tex = job.tex;
                c.tiles_y = lj.tiles_y;
                c.inv_tiles_y = 1.0f / float(lj.tiles_y > 0 ? lj.tiles_y : 1);
                choose_floor_magic(c, lj.img);
                c.img = lj.img;
                c.pts = pts;
                c.k = k;
                c.inv_fx = 1.0 / double(k.fx);
                c.inv_fy = 1.0 / double(k.fy);
                c.k01 = -double(k.s) / (double(k.fx) * double(k.fy));
                c.huber_delta = P.huber_delta;
                {
                    const double fu = k.fx, fv = k.fy;
                    const double ga[6] = {fu, fv, -1.0, -c.inv_fy, c.inv_fx, fv * c.inv_fx};
                    const double gb[6] = {0.0, 0.0, 0.0, -fv, fu, -(fu * c.inv_fy)};
                    for (int t = 0; t < 6; ++t) {
                        c.ga[t] = ga[t];
                        c.gb[t] = gb[t];
                    }
                }
                s_lc = c;
                S.cand_model = S.out_model;
                S.init_phase = 1;  // also: the level's far bitmap and H_outside start empty
            }
            if (warp == 0) {
                __syncwarp();
                if (lane < 12) S.M[lane] = warp_matrix_entry(lane, S.cand_model, s_lc);
            }
            __syncthreads();

            for (;;) {
#if VORS_TIMING
                const long long t_pass0 = clock64();
#endif
                {
                    // ---- consumer warps: four candidates per lane per stage
                    float M[12];
#pragma unroll
                    for (int c = 0; c < 12; ++c) M[c] = S.M[c];
                    PassConst lc;
                    lc.cx = s_lc.cx; lc.cy = s_lc.cy; lc.hu = s_lc.hu; lc.hv = s_lc.hv; lc.lo_u = s_lc.lo_u; lc.lo_v = s_lc.lo_v;
                    lc.magic_u = s_lc.magic_u; lc.magic_v = s_lc.magic_v; lc.rows = s_lc.rows; lc.img_biased = s_lc.img_biased;
                    lc.zero_u = s_lc.zero_u; lc.zero_v = s_lc.zero_v;
                    lc.tex_ku = s_lc.tex_ku; lc.tex_kv = s_lc.tex_kv; lc.tex = s_lc.tex;
                    const Intrinsics k = s_lc.k;
                    constexpr int kNumS = kHuber ? 27 : kNumMoments;  // per-thread sums besides the energy
                    const float huber_delta = s_lc.huber_delta;
                    Acc<kHuber> acc;
                    acc.e = 0.0f;
#pragma unroll
                    for (int c = 0; c < kNumS; ++c) acc.s[c] = 0.0f;
                    Defer df;
                    df.near = lj.defer;
                    df.far = lj.defer_far;
                    df.n_bad = 0;
                    df.n_near = 0;
                    df.stash = &S.stash[warp][0][0];
                    df.n_words = 0;
                    df.wlist = S.near_words[warp];
                    df.any_flip = 0;
                    df.first_pass = __shfl_sync(0xffffffffu, S.init_phase, 0);
                    float* hs = S.hsm[tid];
                    const int gw = rank * kWarps + warp;
                    int n_slots = 0;
                    // software pipeline over candidates: front(q+1) is issued before back(q).  The pipeline is primed
                    // with a candidate that reads nothing and contributes exact zeros.
                    Front fa, fb;
                    fb.pk = 0u; fb.gr = 0u; fb.a = 0.0f; fb.b = 0.0f; fb.rho = 0.0f; fb.fa = 0.0f; fb.fb = 0.0f;
#if VORS_TEX
                    fb.t00 = fb.t10 = fb.t01 = fb.t11 = 0.0f;
#else
                    fb.t00 = fb.t10 = fb.t01 = fb.t11 = 0u;
#endif
                    Front f0 = fb, f1 = fb, f2 = fb;  // tiled records: three pipeline slots
                    const int n_words = n_stages * (kTiled ? kTileWordsBm : kStageWordsBm);  // 32-slot words of the level (padding included)
                    if constexpr (kTiled) {
                        // ---- tiled dense records: every warp streams whole tiles (32 rows x 12 columns = one 3840-byte
                        // stage, one bulk copy) through its ring; a lane owns one row of the tile, a word one column
                        const int tiles_y = s_lc.tiles_y;
                        if (lane == 0) {
                            uint32_t slot = ring_slot;
                            for (int j = 0, c = gw; j < kStages - 1 && c < n_stages; ++j, c += TW) {
                                const uint32_t bar = bar_base + slot * 8u;
                                mbar_expect_tx(bar, kTileBytes);
                                bulk_g2s(ring_base + slot * kRingBytes, pts + size_t(c) * kTileWords, kTileBytes, bar, l2_policy);
                                slot = (slot + 1 == kStages) ? 0u : slot + 1;
                            }
                        }
                        if (df.first_pass) {  // the level's far bitmap starts empty: clear the words of this warp's tiles
                            for (int c = gw; c < n_stages; c += TW)
                                if (lane < kTileWordsBm) df.far[kTileWordsBm * c + lane] = 0u;
                            __syncwarp();
                        }
                        const bool far_lane = lane < kTileWordsBm && !df.first_pass;
                        const unsigned* far_ptr = df.far + (kTileWordsBm * gw + lane);
                        const int far_step = kTileWordsBm * TW;
                        unsigned old_next = (far_lane && gw < n_stages) ? __ldcg(far_ptr) : 0u;
                        // tile (tx, ty) of stage c = tx * tiles_y + ty, advanced by TW stages per iteration
                        int tx = gw / tiles_y, ty = gw - tx * tiles_y;
                        const int dtx = TW / tiles_y, dty = TW - dtx * tiles_y;
                        const float lane_f = float(lane);
                        const int rows_lvl = int(s_lc.rows);
                        // dead lanes (see TileLane): the middle of the inside range, and texture offsets that turn it into the zero page
                        const float mid_u = floorf(lc.hu), mid_v = floorf(lc.hv);
                        const float dead_U = ((mid_u + 0.5f) - lc.cx) * 1048576.0f, dead_V = ((mid_v + 0.5f) - lc.cy) * 1048576.0f;
                        const float dead_ku = lc.tex_ku + (mid_u - lc.zero_u), dead_kv = lc.tex_kv + (mid_v - lc.zero_v);
                        TileLane tl;
                        // thread-dependent on paper (threadIdx.y is 0 in every thread of this 1-D block, which the compiler cannot
                        // know), so that these values live in ordinary registers instead of uniform ones
                        const int zero_y = threadIdx.y;
                        tl.M0v = S.M[0 + zero_y];
                        tl.M4v = S.M[4 + zero_y];
                        tl.M8v = S.M[8 + zero_y];
                        for (int c = gw; c < n_stages; c += TW) {
                            const int c_ahead = c + (kStages - 1) * TW;
                            if (lane == 0 && c_ahead < n_stages) {
                                const uint32_t slot = (ring_slot + kStages - 1 >= kStages) ? ring_slot - 1 : ring_slot + kStages - 1;
                                const uint32_t bar = bar_base + slot * 8u;
                                mbar_expect_tx(bar, kTileBytes);
                                bulk_g2s(ring_base + slot * kRingBytes, pts + size_t(c_ahead) * kTileWords, kTileBytes, bar, l2_policy);
                            }
                            // per-lane constants of the tile while the bytes land
                            tl.a0 = float(kTileCols * tx + zero_y) - lc.cx;  // camera.rs:135-140 starts from these rounded differences too
                            tl.b = (float(kTileRows * ty) + lane_f) - lc.cy;  // (integers: the sum is exact)
                            tl.Ub = fmaf(M[1], tl.b, fmaf(tl.M0v, tl.a0, M[2]));
                            tl.Vb = fmaf(M[5], tl.b, fmaf(tl.M4v, tl.a0, M[6]));
                            tl.Wb = fmaf(M[9], tl.b, fmaf(tl.M8v, tl.a0, M[10]));
                            tl.ku = lc.tex_ku;
                            tl.kv = lc.tex_kv;
                            const int rows_live = min(kTileRows, rows_lvl - kTileRows * ty);  // rows of this tile inside the image
                            if (lane >= rows_live) {  // dead lanes (see TileLane)
                                tl.Ub = dead_U;
                                tl.Vb = dead_V;
                                tl.Wb = 1048576.0f;
                                tl.ku = dead_ku;
                                tl.kv = dead_kv;
                            }
                            mbar_wait(bar_base + ring_slot * 8u, ring_parity);  // TMA bytes have landed
                            const unsigned old_words = old_next;
                            const unsigned old_nz = __ballot_sync(0xffffffffu, old_words != 0u);
                            far_ptr += far_step;
                            if (far_lane && c + TW < n_stages) old_next = __ldcg(far_ptr);
                            const float* sp = &S.ring[warp][ring_slot * kRingWords];
                            // software pipeline, three slots: word J's front (its texture gather) is issued two words before its
                            // back, so two gathers per lane are in flight while a third word is being reduced.  The body is
                            // unrolled over half a tile (six words = two rounds of the three slots) and run twice: twelve words
                            // unrolled do not fit the instruction cache.
#define VORS_TSTEP(J, RHO, GR, TMPL, FNEW, FOLD)                                                                            \
    tile_front<J>(tl, RHO, GR, TMPL, M, w0, jh, old_words, old_nz_h, lc, k, df, hs, lane, FNEW);                            \
    tile_back(FOLD, acc);
#pragma unroll 1
                            for (int h = 0; h < 2; ++h) {
                                const int jh = kTileHalfCols * h, w0 = kTileWordsBm * c + jh;
                                const unsigned old_nz_h = old_nz >> jh;
                                // the lane's six inverse depths, gradients and template values of this half tile
                                const float2* r2 = reinterpret_cast<const float2*>(sp + (h * kTileRows + lane) * kTileHalfCols);
                                const uint2* g2 = reinterpret_cast<const uint2*>(sp + kTileSlots + (h * kTileRows + lane) * kTileHalfCols);
                                const uint32_t* t1 = reinterpret_cast<const uint32_t*>(sp + 2 * kTileSlots) + (h * kTileRows + lane) * (kTileHalfCols / 2);
                                const float2 ra = r2[0], rb = r2[1], rc = r2[2];
                                const uint2 ga = g2[0], gb = g2[1], gc = g2[2];
                                const uint32_t ta = t1[0], tb = t1[1], tc = t1[2];
                                VORS_TSTEP(0, ra.x, ga.x, half_lo(ta), f0, f1)
                                VORS_TSTEP(1, ra.y, ga.y, half_hi(ta), f1, f2)
                                VORS_TSTEP(2, rb.x, gb.x, half_lo(tb), f2, f0)
                                VORS_TSTEP(3, rb.y, gb.y, half_hi(tb), f0, f1)
                                VORS_TSTEP(4, rc.x, gc.x, half_lo(tc), f1, f2)
                                VORS_TSTEP(5, rc.y, gc.y, half_hi(tc), f2, f0)
                                // the second half starts six columns further
                                tl.a0 += float(kTileHalfCols);
                                tl.Ub = fmaf(tl.M0v, float(kTileHalfCols), tl.Ub);
                                tl.Vb = fmaf(tl.M4v, float(kTileHalfCols), tl.Vb);
                                tl.Wb = fmaf(tl.M8v, float(kTileHalfCols), tl.Wb);
                            }
#undef VORS_TSTEP
                            __syncwarp();  // all lanes are done with this slot: the next iteration may refill it
                            ring_slot = (ring_slot + 1 == kStages) ? 0u : ring_slot + 1;
                            ring_parity ^= (ring_slot == 0) ? 1u : 0u;
                            n_slots += kTileCols * rows_live;
                            tx += dtx;
                            ty += dty;
                            if (ty >= tiles_y) {
                                ty -= tiles_y;
                                ++tx;
                            }
                        }
                    } else if (n_words <= kSmallWordsPerWarp * TW) {
                        // ---- small level: a 256-candidate stage per warp would leave most warps idle and the rest with
                        // eight dependent word-steps, so the words are dealt round-robin to all warps and read straight from
                        // global memory (three coalesced loads per word); same front / back arithmetic, no pipelining.
                        for (int wi = gw; wi < n_words; wi += TW) {
                            const uint32_t* rec = pts + size_t(wi >> 1) * (3 * kChunk) + 32 * (wi & 1) + lane;
                            const uint32_t pk = __ldg(rec), gr = __ldg(rec + 2 * kChunk);
                            const float rho = __uint_as_float(__ldg(rec + kChunk));
                            unsigned old = 0u;
                            if (df.first_pass) {
                                if (lane == 0) df.far[wi] = 0u;
                                __syncwarp();
                            } else {
                                old = __ldcg(df.far + wi);
                            }
                            const FrontA xa = front_a(pk, rho, gr, M, lc);
                            front_b<kSkew, kHuber>(xa, wi, 0, old, old != 0u ? 1u : 0u, lc, k, df, hs, lane, fa);
                            back<kSkew, kHuber>(fa, k, huber_delta, acc);
                            n_slots += 32;
                        }
                    } else {
                        // every warp streams its own stages (gw, gw + TW, ...) through its ring: lane 0 issues one bulk copy
                        // per stage, kStages - 1 stages ahead of the one being consumed (SASS UBLKCP + SYNCS)
                        if (lane == 0) {
                            uint32_t slot = ring_slot;
                            for (int j = 0, c = gw; j < kStages - 1 && c < n_stages; ++j, c += TW) {
                                const uint32_t bar = bar_base + slot * 8u;
                                mbar_expect_tx(bar, kStageBytes);
                                bulk_g2s(ring_base + slot * kRingBytes, pts + size_t(c) * kStageWords, kStageBytes, bar, l2_policy);
                                slot = (slot + 1 == kStages) ? 0u : slot + 1;
                            }
                        }
                        if (df.first_pass) {  // the level's far bitmap starts empty: clear the words of this warp's stages
                            for (int c = gw; c < n_stages; c += TW)
                                if (lane < kStageWordsBm) df.far[kStageWordsBm * c + lane] = 0u;
                            __syncwarp();  // orders these stores before lane 0's later stores to the same words
                        }
                        // the stage's far-bitmap words of the previous pass (one per lane 0..kStageWordsBm-1), fetched one iteration ahead
                        const bool far_lane = lane < kStageWordsBm && !df.first_pass;
                        const unsigned* far_ptr = df.far + (kStageWordsBm * gw + lane);  // this lane's word of the stage being fetched
                        const int far_step = kStageWordsBm * TW;
                        unsigned old_next = (far_lane && gw < n_stages) ? __ldcg(far_ptr) : 0u;
                        for (int c = gw; c < n_stages; c += TW) {
                            // refill the slot freed by the previous iteration (every lane consumed its loads before that
                            // iteration's trailing __syncwarp) with the stage kStages - 1 ahead
                            const int c_ahead = c + (kStages - 1) * TW;
                            if (lane == 0 && c_ahead < n_stages) {
                                const uint32_t slot = (ring_slot + kStages - 1 >= kStages) ? ring_slot - 1 : ring_slot + kStages - 1;
                                const uint32_t bar = bar_base + slot * 8u;
                                mbar_expect_tx(bar, kStageBytes);
                                bulk_g2s(ring_base + slot * kRingBytes, pts + size_t(c_ahead) * kStageWords, kStageBytes, bar, l2_policy);
                            }
                            mbar_wait(bar_base + ring_slot * 8u, ring_parity);  // TMA bytes have landed
                            const unsigned old_words = old_next;
                            const unsigned old_nz = __ballot_sync(0xffffffffu, old_words != 0u);
                            far_ptr += far_step;
                            if (far_lane && c + TW < n_stages) old_next = __ldcg(far_ptr);
                            const float* sp = &S.ring[warp][ring_slot * kRingWords] + lane;
                            // word j of the stage = 32 consecutive candidates (chunk j / 2, half j % 2), one per lane
#define VORS_STEP(CH, HALF, FNEW, FOLD)                                                               \
    {                                                                                                 \
        const float* q = sp + (CH) * 3 * kChunk + (HALF) * 32;                                        \
        const FrontA xa = front_a(__float_as_uint(q[0]), q[kChunk], __float_as_uint(q[2 * kChunk]), M, lc); \
        front_b<kSkew, kHuber>(xa, kStageWordsBm * c + 2 * (CH) + (HALF), 2 * (CH) + (HALF), old_words, old_nz, lc, k, df, hs, lane, FNEW); \
        back<kSkew, kHuber>(FOLD, k, huber_delta, acc);                                               \
    }
#pragma unroll(kChunkUnroll)
                            for (int ch = 0; ch < kStageChunks; ++ch) {
                                VORS_STEP(ch, 0, fa, fb)
                                VORS_STEP(ch, 1, fb, fa)
                            }
#undef VORS_STEP
                            __syncwarp();  // all lanes are done with this slot: the next iteration may refill it
                            ring_slot = (ring_slot + 1 == kStages) ? 0u : ring_slot + 1;
                            ring_parity ^= (ring_slot == 0) ? 1u : 0u;
                            n_slots += kStageCand;
                        }
                    }
                    if constexpr (kTiled) {  // the last two words of the last tile (slots 4 % 3 and 5 % 3) are still in flight
                        tile_back(f1, acc);
                        tile_back(f2, acc);
                    } else {
                        back<kSkew, kHuber>(fb, k, huber_delta, acc);
                    }
#if VORS_TIMING
                    const long long t_hot = clock64();
#endif
                    int n_bad = df.n_bad;
                    if (df.n_near > 0) {  // warp-uniform
                        int n_fix = 0;
                        Acc<kHuber> tmp = acc;  // only this copy has its address taken: `acc` itself stays in registers in the hot loop
                        deferred_pass<kSkew, kHuber, kTiled>(warp, lane, gw, TW, n_stages,
                                                             (!kTiled && n_words <= kSmallWordsPerWarp * TW) ? n_words : 0, lj.defer, lj.defer_far, df.n_near, df.wlist, df.n_words,
                                                             reinterpret_cast<uint32_t*>(&S.ring[warp][0]), hs, &tmp, &n_fix, &df.any_flip);
                        acc = tmp;
#pragma unroll
                        for (int d = 16; d > 0; d >>= 1) n_fix += __shfl_xor_sync(0xffffffffu, n_fix, d);
                        n_bad -= n_fix;
                    }
                    if (df.any_flip) {  // warp-uniform: fold this warp's per-thread +-J J^T sums into its f64 slot, re-zero them
#pragma unroll
                        for (int c = 0; c < 21; ++c) {
                            float v = hs[c];
                            hs[c] = 0.0f;
#pragma unroll
                            for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
                            if (lane == 0) S.hout[warp][c] += double(v);
                        }
                    }
#if VORS_TIMING
                    const long long t_def = clock64();
                    if (lane == 0 && lvl < 8) {
                        atomicMax((unsigned long long*)&S.dbgw[lvl][0], (unsigned long long)(t_hot - t_pass0));
                        atomicMax((unsigned long long*)&S.dbgw[lvl][1], (unsigned long long)(t_def - t_hot));
                        atomicAdd((unsigned long long*)&S.dbgw[lvl][2], (unsigned long long)(t_def - t_hot));
                        atomicAdd((unsigned long long*)&S.dbgw[lvl][3], (unsigned long long)(df.n_near));
                    }
#endif
                    // ---- reduce: warp shuffle, then per-CTA f64 sums in fixed order
                    float vals[2 + kNumS];
                    vals[0] = acc.e;
                    vals[1] = lane == 0 ? float(n_slots - n_bad) : 0.0f;
#pragma unroll
                    for (int c = 0; c < kNumS; ++c) vals[2 + c] = acc.s[c];
#pragma unroll
                    for (int c = 0; c < 2 + kNumS; ++c) {
                        float v = vals[c];
#pragma unroll
                        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
                        vals[c] = v;
                    }
                    if (lane == 0) {
#pragma unroll
                        for (int c = 0; c < 2 + kNumS; ++c) S.warp_part[warp][c] = vals[c];
                    }
                }
#if VORS_TIMING
                const long long t_pass1 = clock64();
#endif
                __syncthreads();
#if VORS_TIMING
                const long long t_pass2 = clock64();
#endif
                // ---- reduce + finish + decide + step: one warp alone (redundantly identical in every CTA of a team), the
                // other warps wait at the barrier below.  The last warp: the scheduler favours higher warp ids, so the serial
                // chain is not starved by the co-resident CTA's hot loop.  Per-CTA f64 sums in fixed order -> deterministic.
                if (warp == kSerialWarp) {
#if VORS_TIMING
                    const long long ts0 = clock64();
#endif
                    for (int v = lane; v < kNumRaw; v += 32) {
                        double s = 0.0;
                        if (v < (kHuber ? 29 : kRawH)) {
#pragma unroll
                            for (int w = 0; w < kWarps; ++w) s += double(S.warp_part[w][v]);
                        } else {
#pragma unroll
                            for (int w = 0; w < kWarps; ++w) {
                                s += S.hout[w][v - kRawH];  // cumulative over the level's passes
                            }
                        }
                        if (team > 1)
                            scratch->part[epoch & 1][rank][v] = s;
                        else
                            S.raw[v] = s;
                    }
                    if (team > 1) {  // peer partials through global memory, one counter barrier per pass
                        __threadfence();
                        __syncwarp();
                        ++epoch;
                        if (lane == 0) {
                            atomicAdd(&scratch->counter, 1u);
                            const unsigned target = epoch * unsigned(team);
                            while (ld_acquire_u32(&scratch->counter) < target) __nanosleep(32);
                        }
                        __syncwarp();
                        for (int v = lane; v < kNumRaw; v += 32) {
                            double s = 0.0;
                            const double* pp = &scratch->part[(epoch - 1) & 1][0][v];
                            for (int r = 0; r < team; ++r) s += __ldcg(pp + r * 40);
                            S.raw[v] = s;
                        }
                    }
                    __syncwarp();
#if VORS_TIMING
                    const long long ts1 = clock64();
#endif
#if VORS_OLD_SERIAL
                    if (lane < kNumAcc) S.tot[lane] = finish_entry<kSkew, kHuber>(lane, S.raw, s_lc, S.h_total);
                    __syncwarp();
                    if (lane == 0) {
                        S.point_passes += (unsigned long long)n;
                        if (job.pass_only) {
                            S.n_passes += 1;
                            S.cont = 0;
                        } else {
                            lm_decide(S, P, job, job_idx, lvl, writer);
                        }
                    }
                    __syncwarp();
                    if (S.cont && lane < 12) S.M[lane] = warp_matrix_entry(lane, S.cand_model, s_lc);
#else
                    serial_round<kSkew, kHuber>(P, job_idx, lvl, writer, lane, n, job.pass_only);
#endif
#if VORS_TIMING
                    __syncwarp();
                    if (lane == 0 && lvl < 8) {
                        const long long ts4 = clock64();
                        S.dbgs[lvl][0] += ts1 - ts0; S.dbgs[lvl][1] += 0; S.dbgs[lvl][2] += ts4 - ts1; S.dbgs[lvl][3] += 0;
                    }
#endif
                }
                __syncthreads();
#if VORS_TIMING
                if (tid == 0 && lvl < 8) {
                    const long long t_pass3 = clock64();
                    S.dbg[lvl][0] += t_pass1 - t_pass0;  // warp 0: its share of the pass
                    S.dbg[lvl][1] += t_pass2 - t_pass1;  // warp 0: waiting for the other warps
                    S.dbg[lvl][2] += t_pass3 - t_pass2;  // reduction + serial finish / decide / step
                    S.dbg[lvl][3] += 1;
                }
#endif
                if (!__shfl_sync(0xffffffffu, S.cont, 0)) break;
            }

            if (__shfl_sync(0xffffffffu, job.pass_only, 0)) {
                if (writer && tid == 0) {
                    AlignResult& R = P.results[job_idx];
                    R.pass_n_inside = int(S.tot[1]);
                    R.pass_energy = float(S.tot[0]) / float(int(S.tot[1]));
                    for (int c = 0; c < 6; ++c) R.pass_g[c] = float(S.tot[2 + c]);
                    for (int c = 0; c < 21; ++c) R.pass_H[c] = float(S.tot[8 + c]);
                }
                break;
            }
            if (writer && tid == 0) {
                AlignResult& R = P.results[job_idx];
                R.n_iters[lvl] = S.iter;
                R.energy[lvl] = S.keptE;
                R.n_points[lvl] = n;
            }
            const int failed = __shfl_sync(0xffffffffu, S.failed, 0);
            if (tid == 0 && !failed) S.out_model = S.kept_model;  // inverse_compositional.rs:193
            __syncthreads();
            if (failed) break;  // inverse_compositional.rs:195-199
        }

        // ---- optical flow of the coarsest level's candidates under lm_model (inverse_compositional.rs:210-221)
        float flow = 0.0f;
        if (job.flow_level >= 0 && !job.pass_only) {
            const LevelJob& lj = job.lv[job.flow_level];
            const int n = *lj.n_ptr;
            if (tid == 0) warp_matrix(S.out_model, lj.k, S.M);
            __syncthreads();
            if (rank == 0) {
                float s = 0.0f;
                const int n_slots_flow = kTiled ? lj.n_tiles * kTileSlots : n;
                for (int i = tid; i < n_slots_flow; i += kConsumers) {
                    const SlotRec sr = load_slot<kTiled>(lj.pts, lj.tiles_y, i);
                    const float rho = sr.rho, x = sr.x, y = sr.y;
                    if (kTiled && !(rho == rho && rho != 0.0f)) continue;  // no depth / outside the image
                    const float U = fmaf(S.M[0], x, fmaf(S.M[1], y, fmaf(S.M[3], rho, S.M[2])));
                    const float V = fmaf(S.M[4], x, fmaf(S.M[5], y, fmaf(S.M[7], rho, S.M[6])));
                    const float W = fmaf(S.M[8], x, fmaf(S.M[9], y, fmaf(S.M[11], rho, S.M[10])));
                    s += fabsf(x - U / W) + fabsf(y - V / W);
                }
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
                if (lane == 0) S.warp_part[warp][0] = s;
            }
            __syncthreads();
            if (rank == 0 && tid == 0) {
                double t = 0.0;
                for (int w = 0; w < kWarps; ++w) t += double(S.warp_part[w][0]);
                flow = float(t) / float(n);  // 0/0 = NaN with no candidates, like the reference
            }
        }
#if VORS_TIMING
        if (tid == 0 && (job_idx == 0 || job_idx == 147)) {
            for (int l = 0; l < 8; ++l)
                if (S.dbg[l][3])
                    printf("job %d lvl %d passes %lld: pass %lld wait %lld serial %lld cycles/pass | max hot %lld max post %lld mean post %lld near/pass %lld | reduce %lld finish %lld decide %lld M %lld\n", job_idx, l, S.dbg[l][3],
                           S.dbg[l][0] / S.dbg[l][3], S.dbg[l][1] / S.dbg[l][3], S.dbg[l][2] / S.dbg[l][3], S.dbgw[l][0], S.dbgw[l][1], S.dbgw[l][2] / S.dbg[l][3] / kWarps, S.dbgw[l][3] / S.dbg[l][3],
                           S.dbgs[l][0] / S.dbg[l][3], S.dbgs[l][1] / S.dbg[l][3], S.dbgs[l][2] / S.dbg[l][3], S.dbgs[l][3] / S.dbg[l][3]);
        }
        if (tid == 0) for (int l = 0; l < 8; ++l) S.dbg[l][0] = S.dbg[l][1] = S.dbg[l][2] = S.dbg[l][3] = S.dbgw[l][0] = S.dbgw[l][1] = S.dbgw[l][2] = S.dbgw[l][3] = S.dbgs[l][0] = S.dbgs[l][1] = S.dbgs[l][2] = S.dbgs[l][3] = 0;
#endif
        if (writer && tid == 0) {
            AlignResult& R = P.results[job_idx];
            R.model = S.out_model;
            R.status = S.failed ? VORS_OPTIMIZATION_FAILED : VORS_OK;
            R.optical_flow = flow;
            R.n_passes = S.n_passes;
            R.trace_len = S.trace_len < kTraceCap ? S.trace_len : kTraceCap;
            R.point_passes = S.point_passes;
            R.t_end_ns = global_timer_ns();
        }
        if (dynamic_jobs) {
            if (tid == 0) s_next_job = n_teams + int(atomicAdd(P.job_counter, 1u));
            __syncthreads();
            job_idx = s_next_job;
        } else {
            job_idx += n_teams;
        }
        __syncthreads();
    }
}

}  // namespace

template <typename F>
static cudaError_t for_each_align_kernel(F f) {
    cudaError_t e = f((const void*)k_align<false, false, false>);
    if (e == cudaSuccess) e = f((const void*)k_align<true, false, false>);
    if (e == cudaSuccess) e = f((const void*)k_align<false, true, false>);
    if (e == cudaSuccess) e = f((const void*)k_align<true, true, false>);
    if (e == cudaSuccess) e = f((const void*)k_align<false, false, true>);
    return e;
}

static cudaError_t align_prepare() {
    return for_each_align_kernel(
        [](const void* fn) { return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(LmShared))); });
}

cudaError_t align_query(AlignLaunchInfo* info) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    int sms = 0, per_sm = 1 << 30;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    e = align_prepare();
    if (e != cudaSuccess) return e;
    e = for_each_align_kernel([&per_sm](const void* fn) {  // co-residency bound that holds for every variant
        int n = 0;
        const cudaError_t r = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, kBlock, sizeof(LmShared));
        per_sm = n < per_sm ? n : per_sm;
        return r;
    });
    if (e != cudaSuccess) return e;
    info->block = kBlock;
    info->sm_count = sms;
    info->max_resident_ctas = sms * per_sm;
    return cudaSuccess;
}

cudaError_t launch_align(Launcher& L, const AlignParams& p, int n_teams) {
    const int grid = n_teams * p.team;
    ++L.launches;
    cudaError_t e = align_prepare();
    if (e != cudaSuccess) return e;
    const bool huber = p.huber_delta > 0.0f;
    if (p.tiled && (huber || p.has_skew)) return cudaErrorInvalidValue;  // tiled records exist for zero skew, plain L2 only
    const void* fn = p.tiled ? (const void*)k_align<false, false, true>
                   : huber   ? (p.has_skew ? (const void*)k_align<true, true, false> : (const void*)k_align<false, true, false>)
                             : (p.has_skew ? (const void*)k_align<true, false, false> : (const void*)k_align<false, false, false>);
    void* args[] = {(void*)&p};
    if (p.team > 1)  // co-residency of a team's CTAs is required by the counter barrier: cooperative launch checks it
        return cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kBlock), args, sizeof(LmShared), L.stream);
    return cudaLaunchKernel(fn, dim3(grid), dim3(kBlock), args, sizeof(LmShared), L.stream);
}

}  // namespace vors
