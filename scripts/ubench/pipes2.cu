// Second pipe microbenchmark: does FFMA2 (packed f32x2) free issue slots when mixed with ALU work? (sm_100a)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define REP8(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)
#define FFMA(i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(c1), "f"(c2));
#define FFMB(i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(g[i]) : "f"(c1), "f"(c2));
#define FFMA2(i) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(d[i]) : "l"(dc));
#define FFMA2S(i) asm volatile("{.reg .b64 t; mov.b64 t, {%1,%1}; fma.rn.f32x2 %0, %0, t, t;}" : "+l"(d[i]) : "f"(c1));
#define LOP(i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(iseed), "r"(u[(i + 1) & 7]));
#define LOPI(i) asm volatile("lop3.b32 %0, %0, 0x00ffff00, 0x4b000000, 0xea;" : "+r"(u[i]));
#define H2F(i) { unsigned short h = (unsigned short)v[i]; float t; asm volatile("cvt.f32.f16 %0, %1;" : "=f"(t) : "h"(h)); v[i] = __float_as_uint(t); }
#define I2FP(i) { float t; asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(t) : "r"(v[i])); v[i] = __float_as_uint(t); }
#define IMAD(i) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[i]) : "r"(iseed), "r"(u[(i + 1) & 7]));
#define IMADW(i) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(d[i]) : "r"(u[i]), "r"(v[i]));
#define LDS(i) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(h[i]) : "r"(sa + i * 128) : "memory");
#define FSETSEL(i) f[i] = (f[i] < c1) ? g[i] : c2;

template <int KIND>
__global__ void __launch_bounds__(256) bench(float* out, int iters, float seed, uint32_t iseed) {
    __shared__ float sm[2048];
    float f[8], g[8], h[8];
    u64 d[8];
    uint32_t u[8], v[8];
    for (int i = threadIdx.x; i < 2048; i += 256) sm[i] = seed;
    __syncthreads();
    const uint32_t sa = uint32_t(__cvta_generic_to_shared(sm)) + (threadIdx.x & 31) * 4;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        f[i] = seed + i + threadIdx.x;
        g[i] = seed * i; h[i] = 0.f;
        u[i] = iseed + i * 77u + threadIdx.x;
        v[i] = iseed * 3 + i;
        asm volatile("mov.b64 %0, {%1,%2};" : "=l"(d[i]) : "f"(f[i]), "f"(f[i] + 1.0f));
    }
    u64 dc;
    asm volatile("mov.b64 %0, {%1,%2};" : "=l"(dc) : "f"(seed * 0.5f), "f"(seed * 0.25f));
    const float c1 = seed * 1.0001f, c2 = seed * 0.3f;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (KIND == 0) { REP8(FFMA) REP8(FFMB) REP8(LOP) }            // 2 FFMA + 1 LOP3
            else if (KIND == 1) { REP8(FFMA2) REP8(LOP) }                  // 1 FFMA2 + 1 LOP3
            else if (KIND == 2) { REP8(FFMA) REP8(FFMB) REP8(FFMA) REP8(FFMB) REP8(LOP) }  // 4 FFMA + 1 LOP3
            else if (KIND == 3) { REP8(FFMA2) REP8(FFMA2) REP8(LOP) }      // 2 FFMA2 + 1 LOP3
            else if (KIND == 4) { REP8(FFMA2S) REP8(LOP) }                 // FFMA2 scalar-operand + LOP3
            else if (KIND == 5) { REP8(H2F) REP8(LOP) }                    // HADD2.F32 + LOP3
            else if (KIND == 6) { REP8(H2F) REP8(FFMB) }                   // HADD2.F32 + FFMA
            else if (KIND == 7) { REP8(I2FP) REP8(FFMB) }                  // I2FP + FFMA
            else if (KIND == 8) { REP8(IMAD) }                             // IMAD
            else if (KIND == 9) { REP8(IMADW) }                            // IMAD.WIDE.U32
            else if (KIND == 10) { REP8(IMADW) REP8(LOP) }                 // IMAD.WIDE + LOP3
            else if (KIND == 11) { REP8(FFMA2) REP8(LDS) }                 // FFMA2 + LDS
            else if (KIND == 12) { REP8(FFMA) REP8(FFMA2) }                // FFMA + FFMA2
            else if (KIND == 13) { REP8(LOPI) REP8(FFMB) REP8(FFMA) }      // LOP3 imm-form + 2 FFMA
            else if (KIND == 14) { REP8(FFMA2) REP8(FFMA2) REP8(FFMA2) REP8(LOP) REP8(LDS) }  // 3 FFMA2 + LOP3 + LDS
            else if (KIND == 15) { REP8(FFMA) REP8(FFMB) REP8(FFMA) REP8(FFMB) REP8(FFMA) REP8(FFMB) REP8(LOP) REP8(LDS) }  // 6 FFMA + LOP3 + LDS
            else if (KIND == 16) { REP8(IMAD) REP8(FFMB) }                 // IMAD + FFMA
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float lo, hi;
        asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(d[i]));
        s += f[i] + g[i] + h[i] + lo + hi + __uint_as_float(u[i]) + __uint_as_float(v[i]);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND>
void run(const char* name, int groups_instr, float* out, int sms, double clk_ghz) {
    const int iters = 1000, grid = sms * 8;
    bench<KIND><<<grid, 256>>>(out, 10, 1.0f, 12345u);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a);
    bench<KIND><<<grid, 256>>>(out, iters, 1.0f, 12345u);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double groups = double(grid) * 8 * iters * 32.0;  // warp-level groups (one group = one of each listed instruction)
    const double cycles = ms * 1e-3 * clk_ghz * 1e9;
    const double cyc_per_group_smsp = cycles / (groups / sms / 4.0);
    printf("%-40s %7.3f ms  %5.2f cyc/group/SMSP  (%d instr/group -> IPC/SMSP %.2f)\n", name, ms, cyc_per_group_smsp, groups_instr, groups_instr / cyc_per_group_smsp);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double ghz = clk_khz * 1e-6;
    printf("%s, %d SMs, %.3f GHz nominal\n", p.name, p.multiProcessorCount, ghz);
    float* out;
    cudaMalloc(&out, size_t(p.multiProcessorCount) * 8 * 256 * 4);
    const int s = p.multiProcessorCount;
    run<0>("2 FFMA + LOP3", 3, out, s, ghz);
    run<1>("FFMA2 + LOP3", 2, out, s, ghz);
    run<2>("4 FFMA + LOP3", 5, out, s, ghz);
    run<3>("2 FFMA2 + LOP3", 3, out, s, ghz);
    run<4>("FFMA2(scalar operands) + LOP3", 2, out, s, ghz);
    run<5>("HADD2.F32 + LOP3", 2, out, s, ghz);
    run<6>("HADD2.F32 + FFMA", 2, out, s, ghz);
    run<7>("I2FP + FFMA", 2, out, s, ghz);
    run<8>("IMAD", 1, out, s, ghz);
    run<9>("IMAD.WIDE.U32", 1, out, s, ghz);
    run<10>("IMAD.WIDE.U32 + LOP3", 2, out, s, ghz);
    run<11>("FFMA2 + LDS", 2, out, s, ghz);
    run<12>("FFMA + FFMA2", 2, out, s, ghz);
    run<13>("LOP3(imm) + 2 FFMA", 3, out, s, ghz);
    run<14>("3 FFMA2 + LOP3 + LDS", 5, out, s, ghz);
    run<15>("6 FFMA + LOP3 + LDS", 8, out, s, ghz);
    run<16>("IMAD + FFMA", 2, out, s, ghz);
    return 0;
}
