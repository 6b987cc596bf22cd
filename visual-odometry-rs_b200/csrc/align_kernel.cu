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

constexpr int kWarps = 7;                     // consumer warps per CTA (7 + 1 producer = 8 warps: 2 CTAs/SM at up to 128 registers)
constexpr int kConsumers = kWarps * 32;
constexpr int kBlock = kConsumers + 32;       // + one producer warp
constexpr int kStages = 6;                    // TMA ring depth per consumer warp
constexpr int kStageWords = 3 * kChunk;       // pk | idepth | grad, kChunk 4-byte words each (one chunk-blocked record)
constexpr uint32_t kStageBytes = kStageWords * 4;

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
    float warp_part[kWarps][32];
    double tot[32];
    alignas(8) unsigned long long full_bar[kWarps][kStages];
    alignas(8) unsigned long long empty_bar[kWarps][kStages];
    alignas(128) float ring[kWarps][kStages * kStageWords];
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

struct Acc {
    float e;
    int n;
    float g[6];
    float h[21];  // J J^T of the candidates that fell OUTSIDE (subtracted from H_total)
};

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

// A candidate between its gather issue (front) and its arithmetic (back).  The consumer loop runs the front of
// candidate q+1 before the back of candidate q, so texel latency (L1 miss -> L2) hides behind ~170 instructions.
struct Front {
    uint32_t pk, gr;
    float x, y, rho, a, b;
    uint32_t t00, t10, t01, t11;
    bool live, inside;
};

// front: warp, inside test, texel gathers (branch-free on the common path; `live` is false for the padding lanes
// of a partial last chunk; outside / padding candidates read texel (0,0) and later contribute exact zeros).
__device__ __forceinline__ void front(bool live, uint32_t pk, float rho, uint32_t gr, const float (&M)[12], const Intrinsics& k,
                                      const Pose* __restrict__ model, const uint8_t* __restrict__ img, int rows, int wm2i, int hm2i,
                                      float wm2, float hm2, Front& f) {
    // padding lanes of a partial last chunk hold (pk, rho, grad) = 0 (k_compact_scan): finite, and J = 0 exactly
    const float x = u2f(pk & 0xFFFu), y = u2f((pk >> 12) & 0xFFFu);
    const float U = fmaf(M[0], x, fmaf(M[1], y, fmaf(M[3], rho, M[2])));
    const float V = fmaf(M[4], x, fmaf(M[5], y, fmaf(M[7], rho, M[6])));
    const float W = fmaf(M[8], x, fmaf(M[9], y, fmaf(M[11], rho, M[10])));
    const float iw = rcp_approx(W);
    float u = U * iw, v = V * iw;
    // within kBandPx of an inside-test boundary the reference's own arithmetic decides (rare)
    const float band = fminf(fminf(fabsf(u), fabsf(u - wm2)), fminf(fabsf(v), fabsf(v - hm2)));
    if (band < kBandPx) {
        const float2 uv = warp_exact(*model, k, x, y, rho);
        u = uv.x;
        v = uv.y;
    }
    // lm_optimizer.rs:231: inside iff 0 <= floor(u) < W-2 and 0 <= floor(v) < H-2; NaN coordinates are outside
    // (float->int of NaN is 0, so NaN needs its own test; +-inf saturate and fail the range test).
    const int iu = __float2int_rd(u), iv = __float2int_rd(v);
    const bool inside = live && (unsigned(iu) < unsigned(wm2i)) && (unsigned(iv) < unsigned(hm2i)) && ((u + v) == (u + v));
    const uint8_t* p = img + (inside ? unsigned(iu * rows + iv) : 0u);  // unsigned offset: no sign extension
    f.t00 = __ldg(p);
    f.t10 = __ldg(p + 1);
    f.t01 = __ldg(p + unsigned(rows));
    f.t11 = __ldg(p + unsigned(rows) + 1);
    f.a = u - u2f(uint32_t(iu) & 0xFFFu);
    f.b = v - u2f(uint32_t(iv) & 0xFFFu);
    f.pk = pk;
    f.gr = gr;
    f.x = x;
    f.y = y;
    f.rho = rho;
    f.live = live;
    f.inside = inside;
}

// back: Jacobian, bilinear sample, residual, accumulate.
template <bool kSkew>
__device__ __forceinline__ void back(const Front& f, const Intrinsics& k, Acc& acc) {
    const float gu = s16_2f(f.gr & 0xFFFFu), gv = s16_2f(f.gr >> 16);
    float J[6];
    jacobian_at<kSkew>(gu, gv, f.x, f.y, f.rho, k, J);
    const float a = f.a, b = f.b;
    // bilinear blend exactly as lm_optimizer.rs:241-246 writes it (a along x, b along y)
    const float val = (1.0f - b) * (1.0f - a) * u2f(f.t00) + b * (1.0f - a) * u2f(f.t10) + (1.0f - b) * a * u2f(f.t01) + b * a * u2f(f.t11);
    const float r = f.inside ? val - u2f(f.pk >> 24) : 0.0f;
    acc.e = fmaf(r, r, acc.e);
    acc.n += f.inside ? 1 : 0;
#pragma unroll
    for (int c = 0; c < 6; ++c) acc.g[c] = fmaf(J[c], r, acc.g[c]);
    // H = H_total - sum over outside candidates of J J^T: only warps that own an outside candidate pay for it
    const bool outside = f.live && !f.inside;
    if (__any_sync(0xffffffffu, outside)) {
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            const float jc = outside ? J[c] : 0.0f;
#pragma unroll
            for (int d = c; d < 6; ++d) acc.h[tri(c, d)] = fmaf(jc, J[d], acc.h[tri(c, d)]);
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
    warp_matrix(S.cand_model, job.lv[lvl].k, S.M);
    S.cont = 1;
}

template <bool kSkew>
__global__ void __launch_bounds__(kBlock, 2) k_align(const AlignParams P) {
    __shared__ LmShared S;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool is_producer = warp == kWarps;
    const int team = P.team;
    const int team_id = blockIdx.x / team, rank = blockIdx.x - team_id * team;
    const int n_teams = gridDim.x / team;
    TeamScratch* scratch = team > 1 ? P.scratch + team_id : nullptr;
    unsigned epoch = 0;      // passes this team has synchronised on so far (same in every CTA of the team)
    uint32_t ring_count = 0; // chunks this warp has consumed (consumer) / this lane has filled (producer lane w)
    const uint64_t l2_policy = l2_evict_first_policy();
    if (tid == 0) {
        for (int w = 0; w < kWarps; ++w)
            for (int st = 0; st < kStages; ++st) {
                mbar_init(smem_u32(&S.full_bar[w][st]), 1);
                mbar_init(smem_u32(&S.empty_bar[w][st]), 1);
            }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
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
            const int rows = lj.rows;
            const int wm2i = lj.cols - 2, hm2i = rows - 2;
            const float wm2 = float(wm2i), hm2 = float(hm2i);
            const Intrinsics k = lj.k;
            const uint32_t* __restrict__ pts = lj.pts;
            const uint8_t* __restrict__ img = lj.img;
            const int n_chunks = (n + kChunk - 1) / kChunk;
            const int TW = team * kWarps;
            if (tid == 0) {
                S.cand_model = S.out_model;
                S.init_phase = 1;
                warp_matrix(S.cand_model, k, S.M);
            }
            __syncthreads();

            for (;;) {
                float vals[kNumAcc];
                if (is_producer) {
                    // ---- producer warp: lane w feeds consumer warp w's ring (chunks gw, gw + TW, ...) with one 768-byte
                    // bulk copy per stage; lanes poll their consumer's empty barrier without blocking each other
                    {
                        int c = rank * kWarps + lane;
                        bool active = lane < kWarps && c < n_chunks;
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
                                    active = c < n_chunks;
                                    did = true;
                                }
                            }
                            if (!__any_sync(0xffffffffu, did)) __nanosleep(64);
                        }
                    }
#pragma unroll
                    for (int c = 0; c < kNumAcc; ++c) vals[c] = 0.0f;
                } else {
                    // ---- consumer warps: two candidates per lane per stage
                    float M[12];
#pragma unroll
                    for (int c = 0; c < 12; ++c) M[c] = S.M[c];
                    Acc acc;
                    acc.e = 0.0f;
                    acc.n = 0;
#pragma unroll
                    for (int c = 0; c < 6; ++c) acc.g[c] = 0.0f;
#pragma unroll
                    for (int c = 0; c < 21; ++c) acc.h[c] = 0.0f;
                    const int gw = rank * kWarps + warp;
                    // software pipeline over candidates: front(q+1) is issued before back(q)
                    Front fx, fy;
                    fy.pk = 0u; fy.gr = 0u; fy.x = 0.0f; fy.y = 0.0f; fy.rho = 1.0f; fy.a = 0.0f; fy.b = 0.0f;
                    fy.t00 = fy.t10 = fy.t01 = fy.t11 = 0u;
                    fy.live = false; fy.inside = false;
                    for (int c = gw; c < n_chunks; c += TW) {
                        const uint32_t stage = ring_count % kStages, use = ring_count / kStages;
                        mbar_wait(smem_u32(&S.full_bar[warp][stage]), use & 1u);  // TMA bytes have landed
                        const float* sp = &S.ring[warp][stage * kStageWords];
                        const uint32_t pk0 = __float_as_uint(sp[lane]), pk1 = __float_as_uint(sp[lane + 32]);
                        const float rho0 = sp[kChunk + lane], rho1 = sp[kChunk + lane + 32];
                        const uint32_t gr0 = __float_as_uint(sp[2 * kChunk + lane]), gr1 = __float_as_uint(sp[2 * kChunk + lane + 32]);
                        __syncwarp();
                        if (lane == 0) mbar_arrive(smem_u32(&S.empty_bar[warp][stage]));  // hand the stage back
                        ++ring_count;
                        const int i0 = c * kChunk + lane;
                        front(i0 < n, pk0, rho0, gr0, M, k, &S.cand_model, img, rows, wm2i, hm2i, wm2, hm2, fx);
                        back<kSkew>(fy, k, acc);
                        front(i0 + 32 < n, pk1, rho1, gr1, M, k, &S.cand_model, img, rows, wm2i, hm2i, wm2, hm2, fy);
                        back<kSkew>(fx, k, acc);
                    }
                    back<kSkew>(fy, k, acc);
                    vals[0] = acc.e;
                    vals[1] = float(acc.n);
#pragma unroll
                    for (int c = 0; c < 6; ++c) vals[2 + c] = acc.g[c];
#pragma unroll
                    for (int c = 0; c < 21; ++c) vals[8 + c] = acc.h[c];
                    // ---- reduce: warp shuffle, then per-CTA f64 sums in fixed order
#pragma unroll
                    for (int c = 0; c < kNumAcc; ++c) {
                        float v = vals[c];
#pragma unroll
                        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
                        vals[c] = v;
                    }
                    if (lane == 0) {
#pragma unroll
                        for (int c = 0; c < kNumAcc; ++c) S.warp_part[warp][c] = vals[c];
                    }
                }
                __syncthreads();
                if (tid < kNumAcc) {
                    double s = 0.0;
#pragma unroll
                    for (int w = 0; w < kWarps; ++w) s += double(S.warp_part[w][tid]);
                    if (team > 1) {
                        scratch->part[epoch & 1][rank][tid] = s;
                        __threadfence();
                    } else {
                        S.tot[tid] = s;
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
                    if (tid < kNumAcc) {
                        double s = 0.0;
                        const double* pp = &scratch->part[(epoch - 1) & 1][0][tid];
                        for (int r = 0; r < team; ++r) s += __ldcg(pp + r * 32);
                        S.tot[tid] = s;
                    }
                }
                __syncthreads();
                // H over the inside set = H_total (all candidates, per keyframe level) - H_outside; an empty inside
                // set must give an exactly zero H (the reference then fails its Cholesky, lm_optimizer.rs:131-133)
                if (tid >= 8 && tid < kNumAcc) S.tot[tid] = S.tot[1] > 0.0 ? lj.h_total[tid - 8] - S.tot[tid] : 0.0;
                __syncthreads();

                // ---- decide + step (serial, redundantly identical in every CTA of the team)
                if (tid == 0) {
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
                    const float x = float(p & 0xFFFu), y = float((p >> 12) & 0xFFFu);
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

cudaError_t align_query(AlignLaunchInfo* info) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    int sms = 0, per_sm = 0;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_align<true>, kBlock, 0);
    if (e != cudaSuccess) return e;
    info->block = kBlock;
    info->sm_count = sms;
    info->max_resident_ctas = sms * per_sm;
    return cudaSuccess;
}

cudaError_t launch_align(Launcher& L, const AlignParams& p, int n_teams) {
    const int grid = n_teams * p.team;
    ++L.launches;
    const void* fn = p.has_skew ? (const void*)k_align<true> : (const void*)k_align<false>;
    if (p.team > 1) {
        // co-residency of a team's CTAs is required by the counter barrier: cooperative launch checks it
        void* args[] = {(void*)&p};
        return cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kBlock), args, 0, L.stream);
    }
    if (p.has_skew)
        k_align<true><<<grid, kBlock, 0, L.stream>>>(p);
    else
        k_align<false><<<grid, kBlock, 0, L.stream>>>(p);
    return cudaGetLastError();
}

}  // namespace vors
