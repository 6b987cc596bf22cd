// Pipe-throughput microbenchmark for the instructions k_align is built from (sm_100a).
// Prints warp-instructions per clock per SM for each instruction class, alone and mixed.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipes pipes.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;

#define REP8(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)

template <int KIND>
__global__ void __launch_bounds__(256) bench(float* out, int iters, float seed, uint32_t iseed) {
    float f[8];
    u64 d[8];
    uint32_t u[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        f[i] = seed + i + threadIdx.x;
        u[i] = iseed + i * 77u + threadIdx.x;
        asm volatile("mov.b64 %0, {%1,%2};" : "=l"(d[i]) : "f"(f[i]), "f"(f[i] + 1.0f));
    }
    u64 dc;
    asm volatile("mov.b64 %0, {%1,%2};" : "=l"(dc) : "f"(seed * 0.5f), "f"(seed * 0.25f));
    const float c1 = seed * 1.0001f, c2 = seed * 0.3f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (KIND == 0) {  // FFMA 3-reg
#define X(i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(c1), "f"(c2));
                REP8(X)
#undef X
            } else if (KIND == 1) {  // FFMA2 3 pair regs
#define X(i) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(d[i]) : "l"(dc));
                REP8(X)
#undef X
            } else if (KIND == 2) {  // FFMA + LOP3 interleaved 1:1
#define X(i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(c1), "f"(c2)); \
             asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(iseed), "r"(u[(i + 1) & 7]));
                REP8(X)
#undef X
            } else if (KIND == 3) {  // FFMA2 + LOP3 interleaved 1:1
#define X(i) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(d[i]) : "l"(dc)); \
             asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(iseed), "r"(u[(i + 1) & 7]));
                REP8(X)
#undef X
            } else if (KIND == 4) {  // LOP3 alone
#define X(i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(iseed), "r"(u[(i + 1) & 7]));
                REP8(X)
#undef X
            } else if (KIND == 5) {  // cvt.rn.f32.u32 (I2FP?)
#define X(i) asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(f[i]) : "r"(u[i])); \
             asm volatile("mov.b32 %0, %1;" : "=r"(u[i]) : "f"(f[i]));
                REP8(X)
#undef X
            } else if (KIND == 6) {  // mad.wide.u32
#define X(i) asm volatile("{.reg .b32 lo, hi; mov.b64 {lo,hi}, %0; mad.wide.u32 %0, lo, %1, %0;}" : "+l"(d[i]) : "r"(iseed));
                REP8(X)
#undef X
            } else if (KIND == 7) {  // FFMA2 + 2x LOP3 (1:2)
#define X(i) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(d[i]) : "l"(dc)); \
             asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(iseed), "r"(u[(i + 1) & 7])); \
             asm volatile("lop3.b32 %0, %0, %1, %2, 0x69;" : "+r"(u[i]) : "r"(iseed), "r"(u[(i + 3) & 7]));
                REP8(X)
#undef X
            } else if (KIND == 8) {  // MUFU.RCP
#define X(i) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
                REP8(X)
#undef X
            } else if (KIND == 9) {  // FADD2 with immediate + rm
#define X(i) asm volatile("add.rm.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(dc));
                REP8(X)
#undef X
            } else if (KIND == 10) {  // FMNMX
#define X(i) asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(f[(i + 1) & 7]));
                REP8(X)
#undef X
            } else if (KIND == 11) {  // FFMA2 with scalar-broadcast operand: fold via mov.b64 {c,c}
#define X(i) { u64 b; asm volatile("mov.b64 %0, {%1,%1};" : "=l"(b) : "f"(c1)); asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(d[i]) : "l"(b)); }
                REP8(X)
#undef X
            } else if (KIND == 12) {  // cvt f32<-u8 of a byte (I2F.U8, XU?)
#define X(i) { uint32_t b = u[i] & 0xffu; asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(f[i]) : "r"(b)); u[i] += __float_as_uint(f[i]); }
                REP8(X)
#undef X
            } else if (KIND == 13) {  // FFMA2 : FFMA : LOP3 = 1:1:1
#define X(i) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(d[i]) : "l"(dc)); \
             asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(c1), "f"(c2)); \
             asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(iseed), "r"(u[(i + 1) & 7]));
                REP8(X)
#undef X
            } else if (KIND == 14) {  // F2I floor
#define X(i) { int v; asm volatile("cvt.rmi.s32.f32 %0, %1;" : "=r"(v) : "f"(f[i])); f[i] = __int_as_float(v); }
                REP8(X)
#undef X
            } else if (KIND == 15) {  // HADD2.F32 (cvt.f32.f16)
#define X(i) { unsigned short h = (unsigned short)u[i]; asm volatile("cvt.f32.f16 %0, %1;" : "=f"(f[i]) : "h"(h)); u[i] = __float_as_uint(f[i]); }
                REP8(X)
#undef X
            } else if (KIND == 16) {  // FSETP + FSEL
#define X(i) f[i] = (f[i] < c1) ? f[(i + 1) & 7] : c2;
                REP8(X)
#undef X
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float lo, hi;
        asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(d[i]));
        s += f[i] + lo + hi + __uint_as_float(u[i]);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND>
void run(const char* name, int ops_per_x, float* out, int sms, double clk_ghz) {
    const int iters = 2000, grid = sms * 8;
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
    const double warp_instr = double(grid) * 8 /*warps*/ * iters * 64.0 * ops_per_x;
    const double cycles = ms * 1e-3 * clk_ghz * 1e9;
    printf("%-44s %8.3f ms  %6.2f warp-instr/clk/SM  (%s)\n", name, ms, warp_instr / cycles / sms, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double ghz = clk_khz * 1e-6;
    printf("%s, %d SMs, %.3f GHz nominal (rates assume the GPU runs at this clock)\n", p.name, p.multiProcessorCount, ghz);
    float* out;
    cudaMalloc(&out, size_t(p.multiProcessorCount) * 8 * 256 * 4);
    const int s = p.multiProcessorCount;
    run<0>("FFMA (3-reg)", 1, out, s, ghz);
    run<1>("FFMA2 (pairs)", 1, out, s, ghz);
    run<11>("FFMA2 (scalar-broadcast operand)", 1, out, s, ghz);
    run<4>("LOP3", 1, out, s, ghz);
    run<2>("FFMA + LOP3 (1:1), total instr", 2, out, s, ghz);
    run<3>("FFMA2 + LOP3 (1:1), total instr", 2, out, s, ghz);
    run<7>("FFMA2 + 2 LOP3 (1:2), total instr", 3, out, s, ghz);
    run<13>("FFMA2 + FFMA + LOP3 (1:1:1), total instr", 3, out, s, ghz);
    run<5>("cvt.f32.u32 (I2FP) [+mov]", 1, out, s, ghz);
    run<12>("cvt.f32.u32 of byte [+lop+iadd]", 1, out, s, ghz);
    run<6>("mad.wide.u32 (IMAD.WIDE)", 1, out, s, ghz);
    run<8>("MUFU.RCP", 1, out, s, ghz);
    run<9>("FADD2.RM", 1, out, s, ghz);
    run<10>("FMNMX", 1, out, s, ghz);
    run<14>("F2I.FLOOR", 1, out, s, ghz);
    run<15>("HADD2.F32 (cvt.f32.f16)", 1, out, s, ghz);
    run<16>("FSETP+FSEL pair, per pair", 1, out, s, ghz);
    return 0;
}
