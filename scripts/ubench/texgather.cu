// Microbenchmark for the round-2 plan: fetch the 2x2 bilinear footprint of a warped candidate with ONE
// texture gather (tld4 on a CUDA array) instead of four byte loads with 64-bit address arithmetic (what k_align does today).
// Both variants run the same synthetic "alignment pass": every lane warps one candidate per step with a smooth displacement,
// samples the image bilinearly, and accumulates the residual; the filler FFMAs stand in for the rest of the loop (warp,
// moments) so that the sample's instructions compete for issue slots like they do in k_align.  NOT part of the library.
//
//   build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/ubench/texgather scripts/ubench/texgather.cu
//   run:   scripts/ubench/texgather            -> ms per pass and candidates/s for ldg4 and tld4, max |difference| of the sums
//
// Layout notes that the measurement has to confirm:
//  * the pyramid is column-major (y fastest): the CUDA array is created with width = rows, height = cols, so texture x = image y;
//  * tld4 returns the footprint the bilinear filter would use at (x, y): texel indices floor(x - 0.5), floor(y - 0.5) and +1.  The
//    kernel passes fl + 1.0 (fl = its own floor of the warped coordinate), i.e. the corner shared by the four texels, which is half
//    a texel away from every footprint boundary: the 8-bit fixed-point coordinate conversion of the texture unit cannot select a
//    different footprint than the kernel's floor;
//  * component order of tld4 (CUDA "tex2Dgather"): .x = (i, j+1), .y = (i+1, j+1), .z = (i+1, j), .w = (i, j) in texture (x, y);
//  * border address mode returns 0 outside: the zero page of the linear layout for free.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x)                                                                                     \
    do {                                                                                          \
        cudaError_t e_ = (x);                                                                     \
        if (e_ != cudaSuccess) {                                                                  \
            fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_));    \
            exit(1);                                                                              \
        }                                                                                         \
    } while (0)

constexpr int kRows = 480, kCols = 640, kImages = 296, kWarps = 10, kFiller = 40;

struct Params {
    const uint8_t* linear;            // kImages column-major images
    const cudaTextureObject_t* tex;   // kImages texture objects over gather-enabled arrays
    float* out;                       // one sum per CTA
    float shift_u, shift_v;           // displacement of the synthetic warp (pixels)
    int passes;
};

// candidate (x, y) of `image` -> warped coordinates: a smooth, depth-like displacement so that neighbouring lanes read
// neighbouring texels (as in dense alignment) but not the identical address pattern
__device__ __forceinline__ void warp_point(int x, int y, const Params& p, float& u, float& v) {
    const float a = float(x) - 0.5f * kCols, b = float(y) - 0.5f * kRows;
    const float w = 1.0f + 1e-4f * a - 5e-5f * b;
    float iw;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(iw) : "f"(w));
    u = fmaf(a + p.shift_u, iw, 0.5f * kCols);
    v = fmaf(b + p.shift_v, iw, 0.5f * kRows);
}

template <bool kGather>
__global__ void __launch_bounds__(kWarps * 32, 2) k_pass(const Params p) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, img = blockIdx.x % kImages;
    const uint8_t* base = p.linear + size_t(img) * kRows * kCols;
    const cudaTextureObject_t tex = p.tex[img];
    float acc = 0.0f, f[4] = {1.0f, 2.0f, 3.0f, 4.0f};
    const int n_words = kRows * kCols / 32;
    for (int pass = 0; pass < p.passes; ++pass) {
        for (int wi = warp; wi < n_words; wi += kWarps) {
            const int i = wi * 32 + lane, x = i / kRows, y = i - x * kRows;
            float u, v;
            warp_point(x, y, p, u, v);
            const bool ok = (fabsf(u - 0.5f * (kCols - 2)) < 0.5f * (kCols - 2) - 0.01f) & (fabsf(v - 0.5f * (kRows - 2)) < 0.5f * (kRows - 2) - 0.01f);
            const float fu = floorf(u), fv = floorf(v);
            float t00, t10, t01, t11;  // (x, y), (x, y+1), (x+1, y), (x+1, y+1)
            if (kGather) {
                // texture x = image y (rows are the fast axis); outside candidates read the border (zeros)
                const float ty = ok ? fu + 1.0f : -8.0f, tx = ok ? fv + 1.0f : -8.0f;
                float4 q;
                asm volatile("tld4.r.2d.v4.f32.f32 {%0, %1, %2, %3}, [%4, {%5, %6}];"
                             : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w)
                             : "l"(tex), "f"(tx), "f"(ty));
                // texture (i, j) = (image y, image x): .w = (y, x), .z = (y+1, x), .x = (y, x+1), .y = (y+1, x+1)
                t00 = q.w; t10 = q.z; t01 = q.x; t11 = q.y;
            } else {
                const uint8_t* ptr = base + (ok ? size_t(int(fu)) * kRows + size_t(int(fv)) : 0);
                t00 = float(__ldg(ptr)); t10 = float(__ldg(ptr + 1)); t01 = float(__ldg(ptr + kRows)); t11 = float(__ldg(ptr + kRows + 1));
                if (!ok) t00 = t10 = t01 = t11 = 0.0f;
            }
            const float a = u - fu, b = v - fv;
            const float top = fmaf(a, t01 - t00, t00), bot = fmaf(a, t11 - t10, t10);
            const float val = fmaf(b, bot - top, top);
            acc += ok ? val : 0.0f;
#pragma unroll
            for (int k = 0; k < kFiller; ++k) f[k & 3] = fmaf(f[k & 3], 1.0000001f, val);
        }
    }
    acc += 1e-30f * (f[0] + f[1] + f[2] + f[3]);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    __shared__ float part[kWarps];
    if (lane == 0) part[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.0f;
        for (int w = 0; w < kWarps; ++w) s += part[w];
        p.out[blockIdx.x] = s;
    }
}

int main() {
    const size_t I = size_t(kRows) * kCols;
    std::vector<uint8_t> host(I * kImages);
    uint32_t rng = 12345u;
    for (size_t i = 0; i < host.size(); ++i) {
        rng = rng * 1664525u + 1013904223u;
        const size_t k = i % I;
        host[i] = uint8_t(128 + 60 * sinf(0.05f * float(k / kRows)) * cosf(0.07f * float(k % kRows)) + float((rng >> 24) & 7));
    }
    uint8_t* d_linear;
    CK(cudaMalloc(&d_linear, host.size()));
    CK(cudaMemcpy(d_linear, host.data(), host.size(), cudaMemcpyHostToDevice));

    std::vector<cudaTextureObject_t> tex(kImages);
    std::vector<cudaArray_t> arrays(kImages);
    const cudaChannelFormatDesc fmt = cudaCreateChannelDesc<uint8_t>();
    for (int i = 0; i < kImages; ++i) {
        CK(cudaMallocArray(&arrays[i], &fmt, kRows, kCols, cudaArrayTextureGather));  // width = rows (fast axis), height = cols
        CK(cudaMemcpy2DToArray(arrays[i], 0, 0, host.data() + size_t(i) * I, kRows, kRows, kCols, cudaMemcpyHostToDevice));
        cudaResourceDesc res{};
        res.resType = cudaResourceTypeArray;
        res.res.array.array = arrays[i];
        cudaTextureDesc td{};
        td.addressMode[0] = td.addressMode[1] = cudaAddressModeBorder;
        td.filterMode = cudaFilterModePoint;
        td.readMode = cudaReadModeNormalizedFloat;  // texel / 255: the u8 -> f32 conversion moves into the texture unit
        td.normalizedCoords = 0;
        CK(cudaCreateTextureObject(&tex[i], &res, &td, nullptr));
    }
    cudaTextureObject_t* d_tex;
    CK(cudaMalloc(&d_tex, sizeof(cudaTextureObject_t) * kImages));
    CK(cudaMemcpy(d_tex, tex.data(), sizeof(cudaTextureObject_t) * kImages, cudaMemcpyHostToDevice));
    float* d_out;
    CK(cudaMalloc(&d_out, sizeof(float) * kImages * 2));

    Params p{d_linear, d_tex, d_out, 3.3f, -2.7f, 4};
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    std::vector<float> out_l(kImages), out_t(kImages);
    for (int variant = 0; variant < 2; ++variant) {
        p.out = d_out + variant * kImages;
        for (int rep = 0; rep < 3; ++rep) {  // last repetition is the one reported
            CK(cudaEventRecord(e0));
            if (variant == 0)
                k_pass<false><<<kImages, kWarps * 32>>>(p);
            else
                k_pass<true><<<kImages, kWarps * 32>>>(p);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
        }
        CK(cudaGetLastError());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double cand = double(I) * kImages * p.passes;
        printf("%s: %.3f ms for %d passes over %d images = %.1f G candidates/s\n", variant ? "tld4" : "ldg4", ms, p.passes, kImages, cand / ms * 1e-6);
        CK(cudaMemcpy((variant ? out_t : out_l).data(), p.out, sizeof(float) * kImages, cudaMemcpyDeviceToHost));
    }
    // tld4 returns texel / 255: compare 255 * sum with the byte-load sum (relative difference ~1e-7 expected)
    double worst = 0.0;
    for (int i = 0; i < kImages; ++i) {
        const double a = out_l[i], b = 255.0 * out_t[i];
        const double rel = fabs(a - b) / (fabs(a) + 1e-30);
        if (rel > worst) worst = rel;
    }
    printf("max relative difference of the per-image sums (ldg4 vs 255 * tld4): %.3e\n", worst);
    return 0;
}
