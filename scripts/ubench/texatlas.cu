// Microbenchmark for the round-2 texture path: ONE gather- and surface-enabled 2D CUDA array (an atlas) holds the frames of
// all streams; texels are f16 (0..255 exactly representable: the read returns the exact value, "fp16 texture samples
// accumulated to fp32") or u8 with normalised-float reads (texel / 255); a kernel fills the atlas through surface stores;
// the synthetic alignment pass of texgather.cu samples it with one tld4 per candidate.  Compared with four byte loads.
// NOT part of the library.
//   build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/ubench/texatlas scripts/ubench/texatlas.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
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
constexpr int kPerRow = 64;                        // images side by side along texture x (2D gather textures: <= 32768 x 32768)
constexpr int kCellW = kRows + 2, kCellH = kCols + 2;  // + two never-written (zero) texels: the zero page

struct Params {
    const uint8_t* linear;
    cudaTextureObject_t tex;
    float* out;
    float shift_u, shift_v;
    int passes;
};

__device__ __forceinline__ void warp_point(int x, int y, const Params& p, float& u, float& v) {
    const float a = float(x) - 0.5f * kCols, b = float(y) - 0.5f * kRows;
    const float w = 1.0f + 1e-4f * a - 5e-5f * b;
    float iw;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(iw) : "f"(w));
    u = fmaf(a + p.shift_u, iw, 0.5f * kCols);
    v = fmaf(b + p.shift_v, iw, 0.5f * kRows);
}

// kMode 0: four byte loads; 1: tld4 on the atlas
template <int kMode>
__global__ void __launch_bounds__(kWarps * 32, 2) k_pass(const Params p, float scale) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, img = blockIdx.x % kImages;
    const uint8_t* base = p.linear + size_t(img) * kRows * kCols;
    const float ox = float((img % kPerRow) * kCellW), oy = float((img / kPerRow) * kCellH);
    float acc = 0.0f, f[4] = {1.0f, 2.0f, 3.0f, 4.0f};
    const int n_words = kRows * kCols / 32;
    for (int pass = 0; pass < p.passes; ++pass) {
        for (int wi = warp; wi < n_words; wi += kWarps) {
            const int i = wi * 32 + lane, x = i / kRows, y = i - x * kRows;
            float u, v;
            warp_point(x, y, p, u, v);
            const bool ok = (fabsf(u - 0.5f * (kCols - 2)) < 0.5f * (kCols - 2) - 0.01f) & (fabsf(v - 0.5f * (kRows - 2)) < 0.5f * (kRows - 2) - 0.01f);
            // outside candidates are pointed at the cell's zero corner (image coordinates (kCols, kRows))
            const float uu = ok ? u : float(kCols), vv = ok ? v : float(kRows);
            const float fu = floorf(uu), fv = floorf(vv);
            float t00, t10, t01, t11;
            if (kMode == 1) {
                const float ty = fu + 1.0f + oy, tx = fv + 1.0f + ox;
                float4 q;
                asm volatile("tld4.r.2d.v4.f32.f32 {%0, %1, %2, %3}, [%4, {%5, %6}];"
                             : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w)
                             : "l"(p.tex), "f"(tx), "f"(ty));
                t00 = q.w; t10 = q.z; t01 = q.x; t11 = q.y;
            } else {
                const uint8_t* ptr = base + (ok ? size_t(int(fu)) * kRows + size_t(int(fv)) : 0);
                t00 = float(__ldg(ptr)); t10 = float(__ldg(ptr + 1)); t01 = float(__ldg(ptr + kRows)); t11 = float(__ldg(ptr + kRows + 1));
                if (!ok) t00 = t10 = t01 = t11 = 0.0f;
            }
            const float a = uu - fu, b = vv - fv;
            const float top = fmaf(a, t01 - t00, t00), bot = fmaf(a, t11 - t10, t10);
            const float val = scale * fmaf(b, bot - top, top);
            acc += val;
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

// fills the atlas from the linear column-major images through surface stores (what the pyramid kernel will do)
template <bool kHalf>
__global__ void k_fill(const uint8_t* __restrict__ linear, cudaSurfaceObject_t surf) {
    const int img = blockIdx.y;
    const uint8_t* src = linear + size_t(img) * kRows * kCols;
    const int ox = (img % kPerRow) * kCellW, oy = (img / kPerRow) * kCellH;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kRows * kCols; i += gridDim.x * blockDim.x) {
        const int x = i / kRows, y = i - x * kRows;
        if (kHalf)
            surf2Dwrite<unsigned short>(__half_as_ushort(__float2half_rn(float(src[i]))), surf, (ox + y) * 2, oy + x);
        else
            surf2Dwrite<unsigned char>(src[i], surf, ox + y, oy + x);
    }
}

__global__ void k_to_half(const uint8_t* __restrict__ in, __half* __restrict__ out, size_t n) {
    for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) out[i] = __float2half_rn(float(in[i]));
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
    float* d_out;
    CK(cudaMalloc(&d_out, sizeof(float) * kImages));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int W = kPerRow * kCellW, H = ((kImages + kPerRow - 1) / kPerRow) * kCellH;
    std::vector<float> ref(kImages), got(kImages);
    for (int variant = 0; variant < 3; ++variant) {  // 0 ldg4, 1 atlas u8 normalised, 2 atlas f16
        Params p{d_linear, 0, d_out, 3.3f, -2.7f, 4};
        cudaArray_t arr = nullptr;
        cudaSurfaceObject_t surf = 0;
        float fill_ms = 0.0f;
        if (variant) {
            const cudaChannelFormatDesc fmt = variant == 1 ? cudaCreateChannelDesc<uint8_t>() : cudaCreateChannelDescHalf();
            bool surface_ok = true;
            if (cudaMallocArray(&arr, &fmt, W, H, cudaArrayTextureGather | cudaArraySurfaceLoadStore) != cudaSuccess) {
                (void)cudaGetLastError();
                surface_ok = false;
                printf("   gather + surface flags together: rejected for %d x %d; gather-only array, filled by cudaMemcpy2DToArray\n", W, H);
                CK(cudaMallocArray(&arr, &fmt, W, H, cudaArrayTextureGather));
            } else {
                printf("   gather + surface flags together: accepted for %d x %d\n", W, H);
            }
            // zero the whole array (gaps = zero page) with a copy from a zeroed linear buffer
            void* z;
            const size_t pitch = size_t(W) * (variant == 1 ? 1 : 2);
            CK(cudaMalloc(&z, pitch * H));
            CK(cudaMemset(z, 0, pitch * H));
            CK(cudaMemcpy2DToArray(arr, 0, 0, z, pitch, pitch, H, cudaMemcpyDeviceToDevice));
            CK(cudaFree(z));
            cudaResourceDesc res{};
            res.resType = cudaResourceTypeArray;
            res.res.array.array = arr;
            if (surface_ok) CK(cudaCreateSurfaceObject(&surf, &res));
            cudaTextureDesc td{};
            td.addressMode[0] = td.addressMode[1] = cudaAddressModeBorder;
            td.filterMode = cudaFilterModePoint;
            td.readMode = variant == 1 ? cudaReadModeNormalizedFloat : cudaReadModeElementType;
            td.normalizedCoords = 0;
            CK(cudaCreateTextureObject(&p.tex, &res, &td, nullptr));
            for (int rep = 0; rep < 2; ++rep) {
                CK(cudaEventRecord(e0));
                if (surface_ok) {
                    if (variant == 1) k_fill<false><<<dim3(64, kImages), 256>>>(d_linear, surf);
                    else k_fill<true><<<dim3(64, kImages), 256>>>(d_linear, surf);
                } else {
                    // one 2D copy per image (column-major image = kCols rows of kRows texels); f16: converted into a linear staging buffer first
                    static __half* stage = nullptr;
                    if (variant == 2 && !stage) CK(cudaMalloc(&stage, I * kImages * sizeof(__half)));
                    if (variant == 2) k_to_half<<<1024, 256>>>(d_linear, stage, I * kImages);
                    for (int i = 0; i < kImages; ++i) {
                        const size_t esz = variant == 1 ? 1 : 2;
                        const void* src = variant == 1 ? (const void*)(d_linear + size_t(i) * I) : (const void*)(stage + size_t(i) * I);
                        CK(cudaMemcpy2DToArrayAsync(arr, size_t((i % kPerRow) * kCellW) * esz, size_t((i / kPerRow) * kCellH), src, kRows * esz, kRows * esz, kCols,
                                                    cudaMemcpyDeviceToDevice, 0));
                    }
                }
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
            }
            CK(cudaGetLastError());
            CK(cudaEventElapsedTime(&fill_ms, e0, e1));
        }
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaEventRecord(e0));
            if (variant == 0) k_pass<0><<<kImages, kWarps * 32>>>(p, 1.0f);
            else k_pass<1><<<kImages, kWarps * 32>>>(p, variant == 1 ? 255.0f : 1.0f);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
        }
        CK(cudaGetLastError());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double cand = double(I) * kImages * p.passes;
        const char* names[] = {"ldg4 (linear u8)", "tld4 atlas u8 normalised", "tld4 atlas f16"};
        printf("%-26s %.3f ms for %d passes over %d images = %.1f G candidates/s", names[variant], ms, p.passes, kImages, cand / ms * 1e-6);
        if (variant) printf("   (surface fill of %d frames: %.3f ms = %.1f GB/s read)", kImages, fill_ms, double(I) * kImages / fill_ms * 1e-6);
        printf("\n");
        CK(cudaMemcpy((variant ? got : ref).data(), d_out, sizeof(float) * kImages, cudaMemcpyDeviceToHost));
        if (variant) {
            double worst = 0.0;
            int exact = 0;
            for (int i = 0; i < kImages; ++i) {
                const double rel = fabs(double(ref[i]) - double(got[i])) / (fabs(double(ref[i])) + 1e-30);
                if (rel > worst) worst = rel;
                exact += ref[i] == got[i];
            }
            printf("   vs ldg4: max relative difference of the per-image sums %.3e, %d of %d sums bit-identical\n", worst, exact, kImages);
            CK(cudaDestroyTextureObject(p.tex));
            if (surf) CK(cudaDestroySurfaceObject(surf));
            CK(cudaFreeArray(arr));
        }
    }
    return 0;
}
