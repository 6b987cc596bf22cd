// align_kernel.cu — the persistent direct-alignment kernel, sm_100a.
//
// One launch runs COMPLETE alignments: for every job (keyframe, frame, prior) a team of `team` CTAs
// walks the pyramid coarse to fine and, per level, runs the reference's Levenberg-Marquardt loop
// (src/math/optimizer.rs:57-70 + src/core/track/lm_optimizer.rs:113-192) entirely on the device:
//
//   pass      warp-specialised.  A PRODUCER warp stages the level's three candidate streams (12 B per
//             candidate) into shared memory with TMA bulk copies (cp.async.bulk, SASS UBLKCP) through a
//             full/empty mbarrier ring of kStages x 64 candidates per consumer warp, so HBM latency is
//             covered by ~74 KB in flight per SM instead of by occupancy.  Eight CONSUMER warps evaluate
//             two candidates per lane per stage, branch-free: warp with the folded 3x4 matrix of lie.cuh
//             (lm_optimizer.rs:213-219), the reference's conservative inside rule, f32 bilinear sample of
//             u8 texels (lm_optimizer.rs:227-251), residual against the template value, Jacobian recomputed
//             in registers (inverse_compositional.rs:313-341), and sum r^2, n_inside, g = sum J r.
//             H = sum J J^T over the inside set is formed as H_total - H_outside: H_total is precomputed per
//             keyframe level (k_h_total), so only candidates that fall outside accumulate J J^T, under a
//             warp-uniform branch (eval_energy + compute_eval_data, lm_optimizer.rs:68-107, fused);
//   reduce    warp shuffles -> shared memory -> f64 per-CTA partials -> (team > 1) peer partials
//             through global memory with one counter barrier per pass; fixed order => deterministic;
//   decide    thread 0 of every CTA redundantly replays accept / reject / stop (lm_optimizer.rs:
//             140-192), damps, solves the 6x6 system by Cholesky, applies se3::exp and the first
//             order renormalisation (lm_optimizer.rs:123-136, 198-209) and publishes the next
//             candidate model's warp matrix; no host round trip anywhere in the loop.
//
// team == 1 is the throughput configuration (one alignment per CTA, two CTAs per SM so one CTA's
// serial solve overlaps the other's pass); team > 1 trades efficiency for latency on few streams.
// No tensor cores: the work is ~130 scalar f32 instructions per 12-byte candidate, bounded by
// instruction issue, not by a dense contraction.
#include <cooperative_groups.h>

#include "vors_device.cuh"

namespace vors {

namespace {

constexpr int kWarps = 9;                     // consumer warps per CTA (+ 1 producer warp = 10 warps)
constexpr int kConsumers = kWarps * 32;
constexpr int kBlock = kConsumers + 32;       // + one producer warp
constexpr int kMinCtasPerSm = 2;              // register cap 96: 20 warps per SM
constexpr int kStageChunks = 2;               // chunk-blocked records per ring stage
constexpr int kStageCand = kStageChunks * kChunk;  // 128 candidates = 4 per lane per stage (== kPtAlign)
constexpr int kStages = 3;                    // TMA ring depth per consumer warp
constexpr int kStageWords = kStageChunks * 3 * kChunk;  // per chunk: pk[64] | idepth[64] | grad[64]
constexpr uint32_t kStageBytes = kStageWords * 4;
constexpr int kHsmStride = 28;
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
    float warp_part[kWarps][16];   // per-warp sums of the pass accumulators (E, n, 11 moments or 6 gradient entries)
    double hout[kWarps][21];       // per-warp sum of J J^T over the candidates that fell outside
    double raw[kNumRaw];           // CTA / team totals of the raw accumulators
    double tot[32];                // finished pass: sum r^2, n_inside, g[6], H[21]
    alignas(8) unsigned long long full_bar[kWarps][kStages];
    alignas(8) unsigned long long empty_bar[kWarps][kStages];
    alignas(128) float ring[kWarps][kStages * kStageWords];
    // per-thread J J^T accumulators of the hot loop's outside candidates: 21 floats at a 28-word stride (16-byte
    // vector accesses of a quarter warp then hit 8 distinct bank groups); zero between passes
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
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {  // non-blocking
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done)
                 : "r"(bar), "r"(parity)
                 : "memory");
    return done != 0;
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

// int -> float conversions: I2F on the XU pipe (measured: the ALU pipe is the busier one in this kernel, so the
// 2-ALU-op magic-number conversion is only used where it fuses with work that is needed anyway).
__device__ __forceinline__ float u2f(uint32_t v) { return float(v); }
__device__ __forceinline__ float s16_2f(uint32_t v16) { return float(int(short(v16))); }
// 1/x: MUFU.RCP seed + one Newton step (2 FMAs) = correctly rounded to within 1 ulp without the slow-path range
// checks of an IEEE division; x = 0 / inf / NaN give inf / NaN, which the inside test rejects like the reference.
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return fmaf(r, fmaf(-x, r, 1.0f), r);
}

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// index of (a, b), a <= b, in the packed upper triangle
__device__ __host__ constexpr int tri(int a, int b) { return a * 6 - a * (a - 1) / 2 + (b - a); }

// Raw per-thread accumulators of one pass.  Zero skew (every intrinsics set the reference ships): instead of the
// six Jacobian entries, eleven moments of (p, q) = (gu r, gv r) from which g = sum J r is assembled once per pass
// in f64 (J is linear in (gu, gv) with coefficients polynomial in a = x - cu, b = y - cv and idepth,
// inverse_compositional.rs:326-340) - 16 instead of 29 instructions per candidate.  With skew the plain g[6].
struct Acc {
    float e;
    float s[11];
};
enum { kSrp, kSrq, kSrt, kSabp, kSbbq, kSq, kSaap, kSp, kSabq, kSbp, kSaq };

__device__ __forceinline__ void accumulate_moments(Acc& acc, float gu, float gv, float a, float b, float rho, float r) {
    const float p = gu * r, q = gv * r;
    const float ap = a * p, bq = b * q;
    acc.s[kSrp] = fmaf(rho, p, acc.s[kSrp]);
    acc.s[kSrq] = fmaf(rho, q, acc.s[kSrq]);
    acc.s[kSrt] = fmaf(rho, ap + bq, acc.s[kSrt]);
    acc.s[kSaap] = fmaf(a, ap, acc.s[kSaap]);
    acc.s[kSbbq] = fmaf(b, bq, acc.s[kSbbq]);
    acc.s[kSabp] = fmaf(b, ap, acc.s[kSabp]);
    acc.s[kSabq] = fmaf(a, bq, acc.s[kSabq]);
    acc.s[kSp] += p;
    acc.s[kSq] += q;
    acc.s[kSbp] = fmaf(b, p, acc.s[kSbp]);
    acc.s[kSaq] = fmaf(a, q, acc.s[kSaq]);
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
    float su, sv;        // 2 / (W-2), 2 / (H-2): maps the inside range [0, W-2) x [0, H-2) to (-1, 1)^2
    float lim_lo, lim_hi;  // 1 -+ band: inside for sure below lim_lo, outside for sure above lim_hi
    float wm2, hm2;
    float zero_u, zero_v;
    float magic_u, magic_v;     // floor constants Cu, Cv (see kMagicBits)
    uint32_t rows;
    int n;
    const uint8_t* img_biased;  // img - ((bits(Cu) * rows + bits(Cv)) mod 2^32)
    const uint8_t* img;
    const uint32_t* pts;        // the level's chunk-blocked candidates (deferred pass)
    Intrinsics k;
};

// Chooses Cu so that (bits(Cu) * rows + bits(Cv)) mod 2^32 plus any texel offset of the slab (< 2^27) cannot wrap.
__device__ __forceinline__ void choose_floor_magic(LevelConst& c, const uint8_t* img) {
    uint32_t du = 0u;
    const uint32_t bv = kMagicBits;
    uint32_t c32 = (kMagicBits + du) * c.rows + bv;
    if (c32 >= 0xF8000000u) {  // shift the constant past the wrap-around: adds du * rows >= 2^27
        du = (0x08000000u + c.rows - 1u) / c.rows;
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
    uint32_t t00, t10, t01, t11;
};

// The kernel's shared state: the LM / ring block in dynamic shared memory, the per-level constants in static.
extern __shared__ __align__(128) unsigned char smem_raw[];
__device__ __forceinline__ LmShared& lm_shared() { return *reinterpret_cast<LmShared*>(smem_raw); }
__shared__ LevelConst s_lc;

// Deferred candidates.  The hot loop only evaluates candidates that are inside the frame for sure (fast warp, margin of
// kBandPx); every other slot - padding, candidates near an inside-test boundary, candidates that fall outside - is
// redirected to the zero page there (exact zero contributions) and flagged in a per-level bitmap in global memory
// (one bit per candidate slot, laid out like the ring stages: word 4*stage + j, bit = lane, slot 2*lane + (j&1) + 64*(j>>1)).
// After the hot loop each warp revisits the flagged slots of its own stages: the reference's own arithmetic
// (lm_optimizer.rs:213-231) decides membership; inside candidates are evaluated in full, outside ones contribute
// J J^T to H_outside.  Keeping all of this (and its function calls) out of the hot loop keeps that loop call-free.
template <bool kSkew>
__device__ __noinline__ void deferred_pass(int warp, int lane, int first_stage, int stage_stride, int n_stages, uint32_t* __restrict__ bitmap,
                                           Acc* acc_io, int* n_fix) {
    LmShared& S = lm_shared();
    const LevelConst& lc = s_lc;
    const uint32_t* __restrict__ pts = lc.pts;
    const Intrinsics k = lc.k;
    const uint8_t* __restrict__ img = lc.img;
    const int rows = int(lc.rows);
    Acc acc = *acc_io;
    int fixed = 0;
    float h[21];
#pragma unroll
    for (int c = 0; c < 21; ++c) h[c] = 0.0f;
    for (int c0 = first_stage; c0 < n_stages; c0 += 32 * stage_stride) {
        const int c = c0 + lane * stage_stride;
        if (c >= n_stages) continue;
        uint4* wp = reinterpret_cast<uint4*>(bitmap) + c;
        const uint4 w4 = *wp;
        if ((w4.x | w4.y | w4.z | w4.w) == 0u) continue;
        *wp = make_uint4(0u, 0u, 0u, 0u);
        const uint32_t words[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
            uint32_t bits = words[j];
            while (bits) {
                const int l = __ffs(bits) - 1;
                bits &= bits - 1u;
                const int i = c * kStageCand + 2 * l + (j & 1) + kChunk * (j >> 1);
                if (i >= lc.n) continue;  // padding
                const uint32_t pk = __ldg(pts + pt_word(i, 0)), gr = __ldg(pts + pt_word(i, 2));
                const float rho = __uint_as_float(__ldg(pts + pt_word(i, 1)));
                const float x = float(rec_x(pk)), y = float(rec_y(pk));
                const float gu = rec_gx(gr), gv = rec_gy(gr);
                const float2 uv = warp_exact(S.cand_model, k, x, y, rho);
                // 0 <= floor(u) < W-2  <=>  0 <= u < W-2 (W-2 is an integer); NaN compares false -> outside
                const bool inside = (uv.x >= 0.0f) && (uv.x < lc.wm2) && (uv.y >= 0.0f) && (uv.y < lc.hm2);
                const float ca = x - lc.cx, cb = y - lc.cy;
                if (inside) {
                    const float fu = floorf(uv.x), fv = floorf(uv.y);
                    const float a = uv.x - fu, b = uv.y - fv;
                    const uint8_t* p = img + (size_t(int(fu)) * size_t(rows) + size_t(int(fv)));
                    const float t00 = float(__ldg(p)), t10 = float(__ldg(p + 1)), t01 = float(__ldg(p + rows)), t11 = float(__ldg(p + rows + 1));
                    const float val = (1.0f - b) * (1.0f - a) * t00 + b * (1.0f - a) * t10 + (1.0f - b) * a * t01 + b * a * t11;
                    const float r = val - float(rec_tmpl(pk));
                    acc.e = fmaf(r, r, acc.e);
                    ++fixed;
                    if (kSkew) {
                        float J[6];
                        jacobian_centred<true>(gu, gv, ca, cb, rho, k, J);
#pragma unroll
                        for (int q = 0; q < 6; ++q) acc.s[q] = fmaf(J[q], r, acc.s[q]);
                    } else {
                        accumulate_moments(acc, gu, gv, ca, cb, rho, r);
                    }
                } else {
                    float J[6];
                    jacobian_centred<kSkew>(gu, gv, ca, cb, rho, k, J);
#pragma unroll
                    for (int q = 0; q < 6; ++q)
#pragma unroll
                        for (int d = q; d < 6; ++d) h[tri(q, d)] = fmaf(J[q], J[d], h[tri(q, d)]);
                }
            }
        }
    }
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 21; ++c) {
        float v = h[c];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        if (lane == 0) S.hout[warp][c] += double(v);
    }
    *acc_io = acc;
    *n_fix = fixed;
}

// Per-level constants the common path keeps in registers.
struct PassConst {
    float cx, cy, su, sv, lim_lo, magic_u, magic_v, zero_u, zero_v;
    uint32_t rows;
    const uint8_t* img_biased;
};

// Warp-uniform bookkeeping of the slots of a pass that the hot loop did not evaluate.
struct Defer {
    uint32_t* bitmap;  // this level's bitmap (global)
    int n_bad;         // slots of this pass redirected to the zero page so far
    int any_far;       // some candidate of this warp went into hsm
};

// J J^T of this lane's candidate (zero for lanes with `far` false) added to the thread's shared-memory accumulators.
template <bool kSkew>
__device__ __forceinline__ void add_outside(bool far, uint32_t gr, float a, float b, float rho, const Intrinsics& k, float* hs) {
    float J[6];
    jacobian_centred<kSkew>(rec_gx(gr), rec_gy(gr), a, b, rho, k, J);
#pragma unroll
    for (int c = 0; c < 6; ++c) J[c] = far ? J[c] : 0.0f;  // (padding slots carry a NaN inverse depth)
    float4* h4 = reinterpret_cast<float4*>(hs);
    float h[24];
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        const float4 t = h4[q];
        h[4 * q] = t.x; h[4 * q + 1] = t.y; h[4 * q + 2] = t.z; h[4 * q + 3] = t.w;
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) {
#pragma unroll
        for (int d = c; d < 6; ++d) h[tri(c, d)] = fmaf(J[c], J[d], h[tri(c, d)]);
    }
#pragma unroll
    for (int q = 0; q < 6; ++q) h4[q] = make_float4(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]);
}

// front: unpack, warp, inside test, texel gathers; branch-free unless some lane of the warp is not inside for sure.
// `word` is the bitmap word of this call's 32 slots.
template <bool kSkew>
__device__ __forceinline__ void front(uint32_t pk, float rho, uint32_t gr, int word, const float (&M)[12], const PassConst& lc,
                                      const Intrinsics& k, Defer& df, float* hs, int lane, Front& f) {
    // int -> float through the 2^23 magic number: one LOP3 (ALU pipe) + one FADD (FMA pipe) instead of mask + I2F
    const float x = __uint_as_float((pk & 0xFFFu) | 0x4B000000u) - 8388608.0f;
    const float y = float((pk >> 12) & 0xFFFu);
    const float a = x - lc.cx, b = y - lc.cy;  // camera.rs:135-140 starts from these rounded differences too
    const float U = fmaf(M[0], a, fmaf(M[1], b, fmaf(M[3], rho, M[2])));
    const float V = fmaf(M[4], a, fmaf(M[5], b, fmaf(M[7], rho, M[6])));
    const float W = fmaf(M[8], a, fmaf(M[9], b, fmaf(M[11], rho, M[10])));
    const float iw = rcp_approx(W);
    float u = fmaf(U, iw, lc.cx), v = fmaf(V, iw, lc.cy);
    // lm_optimizer.rs:231: inside iff 0 <= floor(u) < W-2 and 0 <= floor(v) < H-2; here: inside with a margin
    const float m = fmaxf(fabsf(fmaf(u, lc.su, -1.0f)), fabsf(fmaf(v, lc.sv, -1.0f)));
    const bool ok = m < lc.lim_lo;  // false for NaN
    const unsigned not_ok = __ballot_sync(0xffffffffu, !ok);
    if (not_ok) {  // warp-uniform, rare
        const bool far = m > s_lc.lim_hi;  // outside for sure (false for NaN): only J J^T is needed, formed here
        const unsigned near = __ballot_sync(0xffffffffu, !ok && !far);
        if (near && lane == 0) df.bitmap[word] = near;  // band / NaN / padding: see deferred_pass
        df.n_bad += __popc(not_ok);
        if (__any_sync(0xffffffffu, far)) {
            add_outside<kSkew>(far, gr, a, b, rho, k, hs);
            df.any_far = 1;
        }
        if (!ok) {
            u = lc.zero_u;
            v = lc.zero_v;
            pk = 0u;   // template 0: r = 0 - 0
            rho = 0.0f;
        }
    }
    // floor and fraction without F2I / I2F (see kMagicBits)
    const float tu = __fadd_rd(u, lc.magic_u), tv = __fadd_rd(v, lc.magic_v);
    f.fa = u - (tu - lc.magic_u);
    f.fb = v - (tv - lc.magic_v);
    const uint8_t* p = lc.img_biased + (__float_as_uint(tu) * lc.rows + __float_as_uint(tv));
    f.t00 = __ldg(p);
    f.t10 = __ldg(p + 1);
    f.t01 = __ldg(p + lc.rows);
    f.t11 = __ldg(p + lc.rows + 1);
    f.pk = pk;
    f.gr = gr;
    f.a = a;
    f.b = b;
    f.rho = rho;
}

// back: bilinear sample, residual, accumulate.
template <bool kSkew>
__device__ __forceinline__ void back(const Front& f, const Intrinsics& k, Acc& acc) {
    const float gu = rec_gx(f.gr), gv = rec_gy(f.gr);
    const float a = f.fa, b = f.fb;
    // bilinear blend exactly as lm_optimizer.rs:241-246 writes it (a along x, b along y)
    const float val = (1.0f - b) * (1.0f - a) * u2f(f.t00) + b * (1.0f - a) * u2f(f.t10) + (1.0f - b) * a * u2f(f.t01) + b * a * u2f(f.t11);
    const float r = val - u2f(f.pk >> 24);
    acc.e = fmaf(r, r, acc.e);
    if (kSkew) {
        float J[6];
        jacobian_centred<true>(gu, gv, f.a, f.b, f.rho, k, J);
#pragma unroll
        for (int c = 0; c < 6; ++c) acc.s[c] = fmaf(J[c], r, acc.s[c]);
    } else {
        accumulate_moments(acc, gu, gv, f.a, f.b, f.rho, r);
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
        for (int c = 0; c < 6; ++c) S.keptg[c] = float(tot[2 + c]);
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
    for (int r = 0; r < 6; ++r)
        for (int c = 0; c < 6; ++c) A[r * 6 + c] = S.keptH[r <= c ? tri(r, c) : tri(c, r)];
    for (int c = 0; c < 6; ++c) A[c * 6 + c] *= 1.0f + S.lam;
    for (int c = 0; c < 6; ++c) b[c] = S.keptg[c];
    if (!cholesky6_solve(A, b)) {
        S.failed = 1;
        S.cont = 0;
        return;
    }
    const Pose delta = se3_exp(b);
    S.cand_model = pose_renormalize(pose_mul(S.kept_model, pose_inverse(delta)));
    warp_matrix(S.cand_model, job.lv[lvl].k, S.M, true);
    S.cont = 1;
}

// Assemble sum r^2, n_inside, g[6], H[21] from the raw totals (one thread, f64).
template <bool kSkew>
__device__ __forceinline__ void finish_pass(LmShared& S, const Intrinsics& k, const double* __restrict__ h_total) {
    const double* raw = S.raw;
    double* tot = S.tot;
    tot[0] = raw[0];
    tot[1] = raw[1];
    if (kSkew) {
        for (int c = 0; c < 6; ++c) tot[2 + c] = raw[2 + c];
    } else {
        const double* m = raw + 2;
        const double fu = k.fx, fv = k.fy;
        tot[2] = fu * m[kSrp];
        tot[3] = fv * m[kSrq];
        tot[4] = -m[kSrt];
        tot[5] = -(m[kSabp] + m[kSbbq]) / fv - fv * m[kSq];
        tot[6] = (m[kSaap] + m[kSabq]) / fu + fu * m[kSp];
        tot[7] = (fv / fu) * m[kSaq] - (fu / fv) * m[kSbp];
    }
    // H over the inside set = H_total (all candidates, per keyframe level) - H_outside; an empty inside set must give an
    // exactly zero H (the reference then fails its Cholesky, lm_optimizer.rs:131-133)
    for (int c = 0; c < 21; ++c) tot[8 + c] = raw[1] > 0.0 ? h_total[c] - raw[13 + c] : 0.0;
}

template <bool kSkew>
__global__ void __launch_bounds__(kBlock, kMinCtasPerSm) k_align(const AlignParams P) {
    LmShared& S = lm_shared();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool is_producer = warp == kWarps;
    const int team = P.team;
    const int team_id = blockIdx.x / team, rank = blockIdx.x - team_id * team;
    const int n_teams = gridDim.x / team;
    TeamScratch* scratch = team > 1 ? P.scratch + team_id : nullptr;
    unsigned epoch = 0;      // passes this team has synchronised on so far (same in every CTA of the team)
    uint32_t ring_count = 0; // stages this warp has consumed (consumer) / this lane has filled (producer lane w)
    const uint64_t l2_policy = l2_evict_first_policy();
    if (tid == 0) {
        for (int w = 0; w < kWarps; ++w)
            for (int st = 0; st < kStages; ++st) {
                mbar_init(smem_u32(&S.full_bar[w][st]), 1);
                mbar_init(smem_u32(&S.empty_bar[w][st]), 1);
            }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < kWarps * 21) (&S.hout[0][0])[tid] = 0.0;
    for (int i = tid; i < kConsumers * kHsmStride; i += kBlock) (&S.hsm[0][0])[i] = 0.0f;
    __syncthreads();

    for (int job_idx = team_id; job_idx < P.n_jobs; job_idx += n_teams) {
        const AlignJob& job = P.jobs[job_idx];
        const bool writer = (rank == 0);
        if (tid == 0) {
            S.out_model = P.init[job_idx];
            S.failed = 0;
            S.n_passes = 0;
            S.trace_len = 0;
            S.point_passes = 0ull;
        }
        __syncthreads();

        for (int lvl = job.lvl_first; lvl >= job.lvl_last; --lvl) {
            const LevelJob& lj = job.lv[lvl];
            const int n = *lj.n_ptr;
            const uint32_t* __restrict__ pts = lj.pts;
            const int n_stages = (n + kStageCand - 1) / kStageCand;
            const int TW = team * kWarps;
            if (tid == 0) {
                const Intrinsics k = lj.k;
                const int wm2i = lj.cols - 2, hm2i = lj.rows - 2;
                LevelConst c;
                c.cx = k.cx;
                c.cy = k.cy;
                c.su = 2.0f / float(wm2i);
                c.sv = 2.0f / float(hm2i);
                const float eps = 2.0f * kBandPx / float(min(wm2i, hm2i));
                c.lim_lo = 1.0f - eps;
                c.lim_hi = 1.0f + eps;
                c.wm2 = float(wm2i);
                c.hm2 = float(hm2i);
                c.zero_u = lj.zero_u;
                c.zero_v = lj.zero_v;
                c.rows = uint32_t(lj.rows);
                c.n = n;
                choose_floor_magic(c, lj.img);
                c.img = lj.img;
                c.pts = pts;
                c.k = k;
                s_lc = c;
                S.cand_model = S.out_model;
                S.init_phase = 1;
                warp_matrix(S.cand_model, k, S.M, true);
            }
            __syncthreads();

            for (;;) {
                if (is_producer) {
                    // ---- producer warp: lane w feeds consumer warp w's ring (stages gw, gw + TW, ...) with one bulk copy
                    // per stage; lanes poll their consumer's empty barrier without blocking each other
                    int c = rank * kWarps + lane;
                    bool active = lane < kWarps && c < n_stages;
                    while (__any_sync(0xffffffffu, active)) {
                        bool did = false;
                        if (active) {
                            const uint32_t stage = ring_count % kStages, use = ring_count / kStages;
                            if (mbar_test(smem_u32(&S.empty_bar[lane][stage]), (use & 1u) ^ 1u)) {  // stage drained
                                const uint32_t bar = smem_u32(&S.full_bar[lane][stage]);
                                mbar_expect_tx(bar, kStageBytes);
                                bulk_g2s(smem_u32(&S.ring[lane][stage * kStageWords]), pts + size_t(c) * kStageWords, kStageBytes, bar,
                                         l2_policy);
                                ++ring_count;
                                c += TW;
                                active = c < n_stages;
                                did = true;
                            }
                        }
                        if (!__any_sync(0xffffffffu, did)) __nanosleep(64);
                    }
                } else {
                    // ---- consumer warps: four candidates per lane per stage
                    float M[12];
#pragma unroll
                    for (int c = 0; c < 12; ++c) M[c] = S.M[c];
                    PassConst lc;
                    lc.cx = s_lc.cx; lc.cy = s_lc.cy; lc.su = s_lc.su; lc.sv = s_lc.sv; lc.lim_lo = s_lc.lim_lo;
                    lc.magic_u = s_lc.magic_u; lc.magic_v = s_lc.magic_v; lc.rows = s_lc.rows; lc.img_biased = s_lc.img_biased;
                    lc.zero_u = s_lc.zero_u; lc.zero_v = s_lc.zero_v;
                    const Intrinsics k = s_lc.k;
                    Acc acc;
                    acc.e = 0.0f;
#pragma unroll
                    for (int c = 0; c < 11; ++c) acc.s[c] = 0.0f;
                    Defer df;
                    df.bitmap = lj.defer;
                    df.n_bad = 0;
                    df.any_far = 0;
                    float* hs = S.hsm[tid];
                    const int gw = rank * kWarps + warp;
                    int n_slots = 0;
                    // software pipeline over candidates: front(q+1) is issued before back(q).  The pipeline is primed
                    // with a candidate that reads nothing and contributes exact zeros.
                    Front fa, fb;
                    fb.pk = 0u; fb.gr = 0u; fb.a = 0.0f; fb.b = 0.0f; fb.rho = 0.0f; fb.fa = 0.0f; fb.fb = 0.0f;
                    fb.t00 = fb.t10 = fb.t01 = fb.t11 = 0u;
                    for (int c = gw; c < n_stages; c += TW) {
                        const uint32_t stage = ring_count % kStages, use = ring_count / kStages;
                        mbar_wait(smem_u32(&S.full_bar[warp][stage]), use & 1u);  // TMA bytes have landed
                        const float* sp = &S.ring[warp][stage * kStageWords] + 2 * lane;
                        // lane owns candidates 2*lane, 2*lane+1 of both chunks of the stage
                        {
                            const uint2 pk = *reinterpret_cast<const uint2*>(sp);
                            const float2 rh = *reinterpret_cast<const float2*>(sp + kChunk);
                            const uint2 gr = *reinterpret_cast<const uint2*>(sp + 2 * kChunk);
                            front<kSkew>(pk.x, rh.x, gr.x, 4 * c, M, lc, k, df, hs, lane, fa);
                            back<kSkew>(fb, k, acc);
                            front<kSkew>(pk.y, rh.y, gr.y, 4 * c + 1, M, lc, k, df, hs, lane, fb);
                            back<kSkew>(fa, k, acc);
                        }
                        {
                            const uint2 pk = *reinterpret_cast<const uint2*>(sp + 3 * kChunk);
                            const float2 rh = *reinterpret_cast<const float2*>(sp + 4 * kChunk);
                            const uint2 gr = *reinterpret_cast<const uint2*>(sp + 5 * kChunk);
                            __syncwarp();
                            if (lane == 0) mbar_arrive(smem_u32(&S.empty_bar[warp][stage]));  // hand the stage back
                            front<kSkew>(pk.x, rh.x, gr.x, 4 * c + 2, M, lc, k, df, hs, lane, fa);
                            back<kSkew>(fb, k, acc);
                            front<kSkew>(pk.y, rh.y, gr.y, 4 * c + 3, M, lc, k, df, hs, lane, fb);
                            back<kSkew>(fa, k, acc);
                        }
                        ++ring_count;
                        n_slots += kStageCand;
                    }
                    back<kSkew>(fb, k, acc);
                    int n_bad = df.n_bad;
                    if (df.any_far) {  // warp-uniform: fold this warp's per-thread J J^T sums into its f64 slot, re-zero them
#pragma unroll
                        for (int c = 0; c < 21; ++c) {
                            float v = hs[c];
                            hs[c] = 0.0f;
#pragma unroll
                            for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
                            if (lane == 0) S.hout[warp][c] += double(v);
                        }
                    }
                    if (n_bad > 0) {  // warp-uniform
                        int n_fix = 0;
                        Acc tmp = acc;  // only this copy has its address taken: `acc` itself stays in registers in the hot loop
                        deferred_pass<kSkew>(warp, lane, gw, TW, n_stages, lj.defer, &tmp, &n_fix);
                        acc = tmp;
#pragma unroll
                        for (int d = 16; d > 0; d >>= 1) n_fix += __shfl_xor_sync(0xffffffffu, n_fix, d);
                        n_bad -= n_fix;
                    }
                    // ---- reduce: warp shuffle, then per-CTA f64 sums in fixed order
                    float vals[13];
                    vals[0] = acc.e;
                    vals[1] = lane == 0 ? float(n_slots - n_bad) : 0.0f;
#pragma unroll
                    for (int c = 0; c < 11; ++c) vals[2 + c] = acc.s[c];
#pragma unroll
                    for (int c = 0; c < 13; ++c) {
                        float v = vals[c];
#pragma unroll
                        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
                        vals[c] = v;
                    }
                    if (lane == 0) {
#pragma unroll
                        for (int c = 0; c < 13; ++c) S.warp_part[warp][c] = vals[c];
                    }
                }
                __syncthreads();
                if (tid < kNumRaw) {
                    double s = 0.0;
                    if (tid < 13) {
#pragma unroll
                        for (int w = 0; w < kWarps; ++w) s += double(S.warp_part[w][tid]);
                    } else {
#pragma unroll
                        for (int w = 0; w < kWarps; ++w) {
                            s += S.hout[w][tid - 13];
                            S.hout[w][tid - 13] = 0.0;
                        }
                    }
                    if (team > 1) {
                        scratch->part[epoch & 1][rank][tid] = s;
                        __threadfence();
                    } else {
                        S.raw[tid] = s;
                    }
                }
                if (team > 1) {
                    __syncthreads();
                    ++epoch;
                    if (tid == 0) {
                        atomicAdd(&scratch->counter, 1u);
                        const unsigned target = epoch * unsigned(team);
                        while (ld_acquire_u32(&scratch->counter) < target) __nanosleep(32);
                    }
                    __syncthreads();
                    if (tid < kNumRaw) {
                        double s = 0.0;
                        const double* pp = &scratch->part[(epoch - 1) & 1][0][tid];
                        for (int r = 0; r < team; ++r) s += __ldcg(pp + r * 40);
                        S.raw[tid] = s;
                    }
                }
                __syncthreads();

                // ---- finish + decide + step (serial, redundantly identical in every CTA of the team)
                if (tid == 0) {
                    finish_pass<kSkew>(S, s_lc.k, lj.h_total);
                    S.point_passes += (unsigned long long)n;
                    if (job.pass_only) {
                        S.n_passes += 1;
                        S.cont = 0;
                    } else {
                        lm_decide(S, P, job, job_idx, lvl, writer);
                    }
                }
                __syncthreads();
                if (!S.cont) break;
            }

            if (job.pass_only) {
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
            const int failed = S.failed;
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
            if (rank == 0 && !is_producer) {
                float s = 0.0f;
                for (int i = tid; i < n; i += kConsumers) {
                    const uint32_t p = lj.pts[pt_word(i, 0)];
                    const float rho = __uint_as_float(lj.pts[pt_word(i, 1)]);
                    const float x = float(rec_x(p)), y = float(rec_y(p));
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
        if (writer && tid == 0) {
            AlignResult& R = P.results[job_idx];
            R.model = S.out_model;
            R.status = S.failed ? VORS_OPTIMIZATION_FAILED : VORS_OK;
            R.optical_flow = flow;
            R.n_passes = S.n_passes;
            R.trace_len = S.trace_len < kTraceCap ? S.trace_len : kTraceCap;
            R.point_passes = S.point_passes;
        }
        __syncthreads();
    }
}

}  // namespace

static cudaError_t align_prepare() {
    cudaError_t e = cudaFuncSetAttribute(k_align<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(LmShared)));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_align<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(LmShared)));
    return e;
}

cudaError_t align_query(AlignLaunchInfo* info) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    int sms = 0, per_sm = 0;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    e = align_prepare();
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_align<true>, kBlock, sizeof(LmShared));
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
    const void* fn = p.has_skew ? (const void*)k_align<true> : (const void*)k_align<false>;
    if (p.team > 1) {
        // co-residency of a team's CTAs is required by the counter barrier: cooperative launch checks it
        void* args[] = {(void*)&p};
        return cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kBlock), args, sizeof(LmShared), L.stream);
    }
    if (p.has_skew)
        k_align<true><<<grid, kBlock, sizeof(LmShared), L.stream>>>(p);
    else
        k_align<false><<<grid, kBlock, sizeof(LmShared), L.stream>>>(p);
    return cudaGetLastError();
}

}  // namespace vors
