// engine.cu — host side of libvors_b200.so: the batched tracker engine and the C ABI.
//
// The host code mirrors the reference's Tracker state machine (src/core/track/inverse_compositional.rs:
// 74-248) for n independent streams: it owns the per-stream HBM slabs, feeds frames to the device,
// launches the precompute kernels and the persistent align kernel, and keeps the poses / timestamps.
// Everything numerical runs on the device; there is no CPU fallback (no device -> VORS_E_CUDA).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "vors_device.cuh"

namespace vors {

static thread_local std::string g_last_error;

static int fail(int code, const char* what, const char* detail = nullptr) {
    g_last_error = what;
    if (detail) {
        g_last_error += ": ";
        g_last_error += detail;
    }
    return code;
}

#define CU_TRY(expr)                                                                 \
    do {                                                                             \
        cudaError_t _e = (expr);                                                     \
        if (_e != cudaSuccess) {                                                     \
            (void)cudaGetLastError();                                                \
            return fail(_e == cudaErrorMemoryAllocation ? VORS_E_NOMEM : VORS_E_CUDA, #expr, cudaGetErrorString(_e)); \
        }                                                                            \
    } while (0)

// Makes `dev` current for the enclosing scope and restores the caller's device afterwards (the library must not
// change the calling thread's current device as a side effect).
struct DeviceScope {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceScope(int dev) {
        int cur = -1;
        err = cudaGetDevice(&cur);
        if (err == cudaSuccess && cur != dev) {
            err = cudaSetDevice(dev);
            if (err == cudaSuccess) prev = cur;
        }
    }
    ~DeviceScope() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};
#define DEVICE_SCOPE(dev)          \
    DeviceScope _device_scope(dev); \
    CU_TRY(_device_scope.err)

static Pose to_pose(const vors_pose& p) { return Pose{{p.t[0], p.t[1], p.t[2]}, {p.q[0], p.q[1], p.q[2], p.q[3]}}; }
static vors_pose from_pose(const Pose& p) {
    vors_pose o;
    o.t[0] = p.t.x; o.t[1] = p.t.y; o.t[2] = p.t.z;
    o.q[0] = p.q.i; o.q[1] = p.q.j; o.q[2] = p.q.k; o.q[3] = p.q.w;
    return o;
}

// multires::limited_sequence + halve shapes (multires.rs:38-88).
static int pyramid_shapes(uint32_t rows, uint32_t cols, uint32_t max_levels, uint32_t* out_rows, uint32_t* out_cols) {
    int n = 0;
    uint32_t r = rows, c = cols;
    for (;;) {
        if (out_rows) out_rows[n] = r;
        if (out_cols) out_cols[n] = c;
        ++n;
        if (!(uint32_t(n) < max_levels)) break;
        if (r / 2 == 0 || c / 2 == 0) break;
        if (n >= 32) break;
        r /= 2;
        c /= 2;
    }
    return n;
}

struct StreamState {  // inverse_compositional.rs:52-60 `State`
    Pose kf_pose = pose_identity();
    Pose cur_pose = pose_identity();
    double kf_depth_ts = 0, kf_img_ts = 0, cur_depth_ts = 0, cur_img_ts = 0;
};

class Engine {
   public:
    vors_config cfg{};
    int n = 0, rows = 0, cols = 0, layout = VORS_COL_MAJOR;
    int device = 0;
    Geom g{};
    Intrinsics intr[kMaxLevels]{};
    Launcher L{};
    Launcher LC{};                   // copy stream: H2D + transposes of the next chunk overlap the align kernel
    cudaEvent_t ev_up[2]{};
    AlignLaunchInfo info{};
    std::vector<StreamState> st;
    bool tracing = false;

    // device slabs ([n] x per-stream extent)
    uint8_t* d_pyr = nullptr;        // frame pyramid of the frame being tracked, pix_stride per stream
    uint8_t* d_stage8 = nullptr;     // its row-major staging, rows*cols per stream
    // second set: the frames announced for the NEXT track call are uploaded (and their pyramids built) here on the copy
    // stream while the current frames are being aligned; the sets swap when that call arrives
    uint8_t* d_pyr_alt = nullptr;
    uint8_t* d_stage8_alt = nullptr;
    AlignJob* d_jobs_alt = nullptr;  // job descriptors pointing into d_pyr_alt
    bool pending = false;            // d_pyr_alt holds prefetched frames
    const uint8_t* pending_dev = nullptr;  // ... announced as a device buffer (else: the host pointers in pending_ptrs)
    std::vector<const uint8_t*> pending_ptrs;
    cudaEvent_t ev_pref{};
    uint16_t* d_stage16 = nullptr;   // row-major depth staging (launch order)
    uint16_t* d_depth = nullptr;     // column-major depth, rows*cols per stream
    uint32_t* d_grad = nullptr;      // gradient pairs of all levels
    uint16_t* d_g2 = nullptr;        // squared gradient norms
    uint8_t* d_mask = nullptr;       // coarse-to-fine masks of all levels
    float* d_idepth = nullptr;       // idepth pyramid maps (NaN = unknown)
    float* d_weight = nullptr;
    uint32_t* d_defer = nullptr;     // align kernel's deferred-slot bitmaps (near, far), 2 x pt_total / 32 words per stream
    uint32_t* d_pts = nullptr;       // chunk-blocked candidates, 3 * pt_total words per stream
    int* d_blk_count = nullptr;
    int* d_n_points = nullptr;       // [n][kMaxLevels]
    double* d_h_total = nullptr;     // [n][kMaxLevels][kHStride]
    int* d_items = nullptr;
    AlignJob* d_jobs = nullptr;
    Pose* d_init = nullptr;
    AlignResult* d_results = nullptr;
    TeamScratch* d_scratch = nullptr;
    vors_trace_rec* d_trace = nullptr;
    uint16_t* d_gradmag = nullptr;   // DSO mode: gradient magnitude of level 0 (one stream at a time)
    uint8_t* d_dso_ws = nullptr;     // DSO mode: selector workspace
    int* h_dso_flags = nullptr;      // pinned
    float* d_tmp = nullptr;          // small scratch (jacobian export etc.)
    size_t tmp_bytes = 0;

    // pinned host mirrors
    Pose* h_init = nullptr;
    AlignResult* h_results = nullptr;
    int* h_items = nullptr;
    int* h_n_points = nullptr;
    AlignJob* h_jobs = nullptr;

    cudaEvent_t ev[5]{};
    float last_ms[4] = {0, 0, 0, 0};
    unsigned long long last_launches = 0, last_point_passes = 0;
    int scratch_teams = 0;

    ~Engine() { destroy(); }

    void destroy() {
        DeviceScope scope(device);
        if (L.stream) cudaStreamSynchronize(L.stream);
        void* dev_ptrs[] = {d_pyr, d_stage8, d_pyr_alt, d_stage8_alt, d_jobs_alt, d_stage16, d_depth, d_grad, d_g2, d_mask, d_idepth, d_weight, d_pts, d_defer,
                            d_blk_count, d_n_points, d_h_total, d_items, d_jobs, d_init, d_results, d_scratch, d_trace, d_tmp, d_gradmag, d_dso_ws};
        for (void* p : dev_ptrs)
            if (p) cudaFree(p);
        void* host_ptrs[] = {h_init, h_results, h_items, h_n_points, h_jobs, h_dso_flags};
        for (void* p : host_ptrs)
            if (p) cudaFreeHost(p);
        for (auto& e : ev)
            if (e) cudaEventDestroy(e);
        if (ev_pref) cudaEventDestroy(ev_pref);
        ev_pref = nullptr;
        for (auto& e : ev_up)
            if (e) cudaEventDestroy(e);
        if (LC.stream) cudaStreamDestroy(LC.stream);
        LC.stream = nullptr;
        if (L.stream) cudaStreamDestroy(L.stream);
        L.stream = nullptr;
        d_pyr = nullptr;
    }

    int setup(const vors_config* c, uint32_t n_, uint32_t rows_, uint32_t cols_, int layout_, bool tracker_rules = true) {
        if (!c || n_ == 0 || rows_ == 0 || cols_ == 0) return fail(VORS_E_INVALID, "null config or empty batch / image");
        if (layout_ != VORS_COL_MAJOR && layout_ != VORS_ROW_MAJOR) return fail(VORS_E_INVALID, "unknown layout");
        if (c->nb_levels == 0 || c->nb_levels > kMaxLevels) return fail(VORS_E_INVALID, "nb_levels must be in 1..VORS_MAX_LEVELS");
        if (rows_ > 4096 || cols_ > 4096) return fail(VORS_E_INVALID, "images larger than 4096x4096 are not supported");
        if (c->candidate_mode > VORS_CANDIDATES_DSO) return fail(VORS_E_INVALID, "unsupported candidate_mode");
        if (c->candidates_diff_threshold > 65535u) return fail(VORS_E_INVALID, "candidates_diff_threshold exceeds u16");
        uint32_t lr[32], lc[32];
        const int levels = pyramid_shapes(rows_, cols_, c->nb_levels, lr, lc);
        // the reference indexes pyramid[nb_levels-1] (inverse_compositional.rs:183-185) and computes
        // `width - 2` on usize (lm_optimizer.rs:231): both panic on degenerate sizes -> invalid argument here
        if (uint32_t(levels) < c->nb_levels) return fail(VORS_E_INVALID, "image too small: pyramid shorter than nb_levels");
        if (tracker_rules && (lr[levels - 1] < 3 || lc[levels - 1] < 3)) return fail(VORS_E_INVALID, "coarsest level smaller than 3x3");

        cfg = *c;
        n = int(n_); rows = int(rows_); cols = int(cols_); layout = layout_;
        int count = 0;
        if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
            (void)cudaGetLastError();
            return fail(VORS_E_CUDA, "no CUDA device available (libvors_b200 has no CPU fallback)");
        }
        if (cfg.idepth_fusion > VORS_FUSION_STATISTICALLY_SIMILAR) return fail(VORS_E_INVALID, "unknown idepth_fusion");
        if (cfg.gradient_operator > VORS_GRADIENT_SCHARR) return fail(VORS_E_INVALID, "unknown gradient_operator");
        if (cfg.device >= 0) {
            if (cfg.device >= count) return fail(VORS_E_INVALID, "device ordinal out of range");
            device = cfg.device;
        } else {
            CU_TRY(cudaGetDevice(&device));
        }
        DEVICE_SCOPE(device);
        cudaDeviceProp prop;
        CU_TRY(cudaGetDeviceProperties(&prop, device));
        if (prop.major != 10) return fail(VORS_E_CUDA, "libvors_b200 is built for sm_100a only; device is not compute capability 10.x");

        g.L = levels;
        int off = 0, boff = 0, poff = 0;
        Intrinsics k{cfg.fx, cfg.fy, cfg.cx, cfg.cy, cfg.skew};
        for (int l = 0; l < levels; ++l) {
            g.rows[l] = int(lr[l]);
            g.cols[l] = int(lc[l]);
            g.off[l] = off;
            g.blk_off[l] = boff;
            g.pt_off[l] = poff;
            poff += (g.rows[l] * g.cols[l] + kPtAlign - 1) / kPtAlign * kPtAlign;
            off += g.rows[l] * g.cols[l];
            boff += (g.rows[l] * g.cols[l] + kCompactBlock - 1) / kCompactBlock;
            intr[l] = k;  // camera.rs:106-108 `multi_res`
            k = half_res(k);
        }
        for (int l = levels; l <= kMaxLevels; ++l) g.blk_off[l] = boff;
        g.pix_total = off;
        g.pix_stride = (off + g.rows[0] + 2 + 15) / 16 * 16;  // + zero page (vors_device.cuh)
        g.pt_total = poff;
        g.blk_total = boff;

        CU_TRY(cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking));
        CU_TRY(cudaStreamCreateWithFlags(&LC.stream, cudaStreamNonBlocking));
        for (auto& e : ev) CU_TRY(cudaEventCreate(&e));
        for (auto& e : ev_up) CU_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&ev_pref, cudaEventDisableTiming));
        CU_TRY(align_query(&info));
        if (info.max_resident_ctas < 1) return fail(VORS_E_CUDA, "align kernel cannot be resident on this device");

        const size_t N = size_t(n), P = size_t(g.pix_stride), I = size_t(rows) * cols, PT = size_t(g.pt_total);
        // + slack past the last slab (zero page reads of the last stream stay inside the allocation with room to spare)
        CU_TRY(cudaMalloc(&d_pyr, N * P + (size_t(1) << 20)));
        CU_TRY(cudaMemsetAsync(d_pyr, 0, N * P + (size_t(1) << 20), L.stream));  // the zero page of every slab stays zero: no kernel writes it
        CU_TRY(cudaMalloc(&d_stage8, N * I));
        CU_TRY(cudaMalloc(&d_pyr_alt, N * P + (size_t(1) << 20)));
        CU_TRY(cudaMemsetAsync(d_pyr_alt, 0, N * P + (size_t(1) << 20), L.stream));
        CU_TRY(cudaMalloc(&d_stage8_alt, N * I));
        CU_TRY(cudaMalloc(&d_jobs_alt, N * sizeof(AlignJob)));
        CU_TRY(cudaMalloc(&d_stage16, N * I * 2));
        CU_TRY(cudaMalloc(&d_depth, N * I * 2));
        CU_TRY(cudaMalloc(&d_grad, N * P * 4));
        CU_TRY(cudaMalloc(&d_g2, N * P * 2));
        CU_TRY(cudaMalloc(&d_mask, N * P));
        CU_TRY(cudaMalloc(&d_idepth, N * P * 4));
        CU_TRY(cudaMalloc(&d_weight, N * P * 4));
        CU_TRY(cudaMalloc(&d_pts, N * PT * 12));
        CU_TRY(cudaMalloc(&d_defer, N * PT / 4));
        CU_TRY(cudaMemsetAsync(d_defer, 0, N * PT / 4, L.stream));  // the align kernel leaves it all-zero after every pass
        // a partial last chunk is staged whole: keep its padding initialised
        CU_TRY(cudaMemsetAsync(d_pts, 0, N * PT * 12, L.stream));
        CU_TRY(cudaMalloc(&d_blk_count, N * size_t(g.blk_total) * 4));
        CU_TRY(cudaMalloc(&d_n_points, N * kMaxLevels * 4));
        CU_TRY(cudaMalloc(&d_h_total, N * kMaxLevels * kHStride * sizeof(double)));
        CU_TRY(cudaMemsetAsync(d_h_total, 0, N * kMaxLevels * kHStride * sizeof(double), L.stream));
        CU_TRY(cudaMalloc(&d_items, N * 4));
        CU_TRY(cudaMalloc(&d_jobs, N * sizeof(AlignJob)));
        CU_TRY(cudaMalloc(&d_init, N * sizeof(Pose)));
        CU_TRY(cudaMalloc(&d_results, N * sizeof(AlignResult)));
        scratch_teams = info.max_resident_ctas;
        CU_TRY(cudaMalloc(&d_scratch, size_t(scratch_teams) * sizeof(TeamScratch)));
        CU_TRY(cudaMemsetAsync(d_n_points, 0, N * kMaxLevels * 4, L.stream));
        CU_TRY(cudaMemsetAsync(d_results, 0, N * sizeof(AlignResult), L.stream));
        CU_TRY(cudaMallocHost(&h_dso_flags, 64));
        if (cfg.candidate_mode == VORS_CANDIDATES_DSO) {
            CU_TRY(cudaMalloc(&d_gradmag, I * 2));
            CU_TRY(cudaMalloc(&d_dso_ws, dso_workspace_bytes(rows, cols)));
        }
        CU_TRY(cudaMallocHost(&h_init, N * sizeof(Pose)));
        CU_TRY(cudaMallocHost(&h_results, N * sizeof(AlignResult)));
        CU_TRY(cudaMallocHost(&h_items, N * 4));
        CU_TRY(cudaMallocHost(&h_n_points, N * kMaxLevels * 4));
        CU_TRY(cudaMallocHost(&h_jobs, N * sizeof(AlignJob)));
        std::memset(h_n_points, 0, N * kMaxLevels * 4);
        st.assign(N, StreamState{});

        // static part of the job descriptors: full alignment of stream i against its keyframe
        std::swap(d_pyr, d_pyr_alt);
        for (int i = 0; i < n; ++i) fill_job(h_jobs[i], i, levels - 1, 0, levels - 1, 0);
        CU_TRY(cudaMemcpyAsync(d_jobs_alt, h_jobs, N * sizeof(AlignJob), cudaMemcpyHostToDevice, L.stream));
        CU_TRY(cudaStreamSynchronize(L.stream));
        std::swap(d_pyr, d_pyr_alt);
        for (int i = 0; i < n; ++i) fill_job(h_jobs[i], i, levels - 1, 0, levels - 1, 0);
        CU_TRY(cudaMemcpyAsync(d_jobs, h_jobs, N * sizeof(AlignJob), cudaMemcpyHostToDevice, L.stream));
        CU_TRY(cudaStreamSynchronize(L.stream));
        return VORS_OK;
    }

    void fill_job(AlignJob& j, int stream, int lvl_first, int lvl_last, int flow_level, int pass_only) const {
        std::memset(&j, 0, sizeof(j));
        const size_t base = size_t(stream) * g.pix_stride;
        const size_t pbase = size_t(stream) * g.pt_total;
        for (int l = 0; l < g.L; ++l) {
            LevelJob& lj = j.lv[l];
            lj.pts = d_pts + 3 * (pbase + g.pt_off[l]);
            lj.defer = d_defer + (pbase + g.pt_off[l]) / 32;
            lj.defer_far = d_defer + (size_t(n) * g.pt_total + pbase + g.pt_off[l]) / 32;
            lj.img = d_pyr + base + g.off[l];
            lj.n_ptr = d_n_points + stream * kMaxLevels + l;
            lj.h_total = d_h_total + (size_t(stream) * kMaxLevels + l) * kHStride;
            lj.rows = g.rows[l];
            lj.cols = g.cols[l];
            const int zoff = g.pix_total - g.off[l];  // from this level's image to the slab's zero page
            lj.zero_u = float(zoff / g.rows[l]);
            lj.zero_v = float(zoff % g.rows[l]);
            lj.k = intr[l];
        }
        j.lvl_first = lvl_first;
        j.lvl_last = lvl_last;
        j.flow_level = flow_level;
        j.pass_only = pass_only;
    }

    // ---- uploads ------------------------------------------------------------------------------
    // Frames of streams [start, start + m) into level 0 of their frame pyramids (column-major), on launcher X's stream.
    int upload_images_host_range(const uint8_t* const* img, int start, int m, Launcher& X, uint8_t* pyr_slab = nullptr,
                                 uint8_t* stage_slab = nullptr) {
        const size_t I = size_t(rows) * cols;
        bool contiguous = true;
        for (int i = 1; i < m && contiguous; ++i) contiguous = (img[start + i] == img[start] + size_t(i) * I);
        uint8_t* pyr0 = (pyr_slab ? pyr_slab : d_pyr) + size_t(start) * g.pix_stride;
        if (layout == VORS_COL_MAJOR) {
            if (contiguous) {
                CU_TRY(cudaMemcpy2DAsync(pyr0, size_t(g.pix_stride), img[start], I, I, size_t(m), cudaMemcpyHostToDevice, X.stream));
            } else {
                for (int i = 0; i < m; ++i)
                    CU_TRY(cudaMemcpyAsync(pyr0 + size_t(i) * g.pix_stride, img[start + i], I, cudaMemcpyHostToDevice, X.stream));
            }
        } else {
            uint8_t* stage = (stage_slab ? stage_slab : d_stage8) + size_t(start) * I;
            if (contiguous) {
                CU_TRY(cudaMemcpyAsync(stage, img[start], I * size_t(m), cudaMemcpyHostToDevice, X.stream));
            } else {
                for (int i = 0; i < m; ++i)
                    CU_TRY(cudaMemcpyAsync(stage + size_t(i) * I, img[start + i], I, cudaMemcpyHostToDevice, X.stream));
            }
            launch_transpose_u8(X, stage, pyr0, size_t(g.pix_stride), nullptr, m, rows, cols);
        }
        return VORS_OK;
    }
    int upload_images_host(const uint8_t* const* img) { return upload_images_host_range(img, 0, n, L); }

    int upload_images_device(const uint8_t* img_dev, Launcher* X = nullptr, uint8_t* pyr_slab = nullptr) {  // column-major, n*rows*cols contiguous
        const size_t I = size_t(rows) * cols;
        CU_TRY(cudaMemcpy2DAsync(pyr_slab ? pyr_slab : d_pyr, size_t(g.pix_stride), img_dev, I, I, size_t(n), cudaMemcpyDeviceToDevice,
                                 (X ? *X : L).stream));
        return VORS_OK;
    }

    // Depth maps of the m streams in h_items into the column-major depth slab.
    int upload_depth_host(const uint16_t* const* depth, int m) {
        const size_t I = size_t(rows) * cols;
        for (int j = 0; j < m; ++j) {
            const int s = h_items[j];
            if (!depth || !depth[s]) return fail(VORS_E_INVALID, "depth map required for a keyframe but pointer is null");
            if (layout == VORS_COL_MAJOR)
                CU_TRY(cudaMemcpyAsync(d_depth + size_t(s) * I, depth[s], I * 2, cudaMemcpyHostToDevice, L.stream));
            else
                CU_TRY(cudaMemcpyAsync(d_stage16 + size_t(j) * I, depth[s], I * 2, cudaMemcpyHostToDevice, L.stream));
        }
        if (layout == VORS_ROW_MAJOR) launch_transpose_u16(L, d_stage16, d_depth, I, d_items, m, rows, cols);
        return VORS_OK;
    }

    // ---- keyframe precompute (inverse_compositional.rs:105-161) for the m streams in d_items --------
    int precompute(const uint16_t* depth_slab, int m) {
        const int dense = cfg.candidate_mode == VORS_CANDIDATES_DENSE;
        launch_gradients(L, g, cfg.gradient_operator == VORS_GRADIENT_SCHARR ? 1 : 0, d_pyr, d_grad,
                         (dense || cfg.candidate_mode == VORS_CANDIDATES_DSO) ? nullptr : d_g2, d_items, m);
        const bool dso = cfg.candidate_mode == VORS_CANDIDATES_DSO;
        if (dso) {
            // BASELINE config 3: DSO selection on the level-0 gradient magnitude (examples/candidates_dso.rs:40-60:
            // sqrt(squared_norm_direct) as u16, nb_iterations_left = 2), one stream at a time (host-driven recursion)
            for (int j = 0; j < m; ++j) {
                const int s = h_items[j];
                launch_sqnorm_direct(L, d_pyr + size_t(s) * g.pix_stride, rows, cols, 1, d_gradmag);
                const int rc = dso_select_device(L, d_gradmag, rows, cols, int(cfg.dso_nb_target ? cfg.dso_nb_target : 2000), 2,
                                                 0x9E3779B97F4A7C15ull, d_mask + size_t(s) * g.pix_stride, d_dso_ws, h_dso_flags,
                                                 nullptr, nullptr);
                if (rc != VORS_OK) return fail(rc, "DSO candidate selection failed");
            }
        } else if (!dense && g.L > 1) {
            launch_c2f(L, g, uint16_t(cfg.candidates_diff_threshold), d_g2, d_mask, d_items, m);
        }
        // a 1-level coarse-to-fine pyramid selects every pixel (coarse_to_fine.rs:19-21)
        launch_idepth(L, g, depth_slab, size_t(rows) * cols, d_mask, dense || (g.L == 1 && !dso), cfg.depth_scale, cfg.idepth_variance,
                      cfg.idepth_fusion == VORS_FUSION_STATISTICALLY_SIMILAR ? 1 : 0, d_idepth,
                      d_weight, d_items, m);
        launch_compact(L, g, d_idepth, d_pyr, d_grad, d_blk_count, d_n_points, d_pts, d_items, m);
        launch_h_total(L, g, intr, d_pts, d_n_points, d_h_total, d_items, m);
        CU_TRY(cudaGetLastError());
        CU_TRY(cudaMemcpyAsync(h_n_points, d_n_points, size_t(n) * kMaxLevels * 4, cudaMemcpyDeviceToHost, L.stream));
        return VORS_OK;
    }

    int set_items_all() {
        for (int i = 0; i < n; ++i) h_items[i] = i;
        CU_TRY(cudaMemcpyAsync(d_items, h_items, size_t(n) * 4, cudaMemcpyHostToDevice, L.stream));
        return VORS_OK;
    }

    // n x Config::init (inverse_compositional.rs:74-100)
    int init_keyframes(const double* depth_ts, const uint16_t* const* depth, const uint16_t* depth_dev, const double* img_ts,
                       const uint8_t* const* img, const uint8_t* img_dev) {
        DEVICE_SCOPE(device);
        int rc;
        if (img_dev) {
            if ((rc = upload_images_device(img_dev)) != VORS_OK) return rc;
        } else {
            if (!img) return fail(VORS_E_INVALID, "null image pointer array");
            for (int i = 0; i < n; ++i)
                if (!img[i]) return fail(VORS_E_INVALID, "null image pointer");
            if ((rc = upload_images_host(img)) != VORS_OK) return rc;
        }
        launch_pyramid(L, g, d_pyr, nullptr, n);
        if ((rc = set_items_all()) != VORS_OK) return rc;
        const uint16_t* depth_slab = depth_dev;
        if (!depth_dev) {
            if ((rc = upload_depth_host(depth, n)) != VORS_OK) return rc;
            depth_slab = d_depth;
        }
        if ((rc = precompute(depth_slab, n)) != VORS_OK) return rc;
        CU_TRY(cudaStreamSynchronize(L.stream));
        for (int i = 0; i < n; ++i) {
            StreamState s;
            s.kf_depth_ts = s.cur_depth_ts = depth_ts ? depth_ts[i] : 0.0;
            s.kf_img_ts = s.cur_img_ts = img_ts ? img_ts[i] : 0.0;
            st[size_t(i)] = s;
        }
        return VORS_OK;
    }

    int ensure_trace() {
        if (tracing && !d_trace) CU_TRY(cudaMalloc(&d_trace, size_t(n) * kTraceCap * sizeof(vors_trace_rec)));
        return VORS_OK;
    }

    AlignParams make_params(int n_jobs, int team, int start = 0) const {
        AlignParams p{};
        p.jobs = d_jobs + start;
        p.init = d_init + start;
        p.results = d_results + start;
        p.trace = tracing ? d_trace + size_t(start) * kTraceCap : nullptr;
        p.scratch = d_scratch;
        p.n_jobs = n_jobs;
        p.team = team;
        p.lm_coef_init = cfg.lm_coef_init;
        p.lm_coef_reject_mult = cfg.lm_coef_reject_mult;
        p.lm_coef_accept_mult = cfg.lm_coef_accept_mult;
        p.energy_delta_stop = cfg.energy_delta_stop;
        p.max_iters = int(cfg.max_iters);
        p.fixed_iters = int(cfg.fixed_iters);
        p.has_skew = cfg.skew != 0.0f ? 1 : 0;
        p.huber_delta = cfg.huber_delta > 0.0f ? cfg.huber_delta : 0.0f;
        return p;
    }

    // CTAs per alignment: 1 when the batch alone fills the device, otherwise spread each alignment over enough CTAs for
    // ~8 level-0 candidates per thread, and keep small keyframes (< 10 k candidates) on one CTA (measured,
    // scripts/latency_breakdown.py: the per-pass team barrier costs more than such a keyframe's whole hot loop; a 1080p
    // coarse-to-fine keyframe wants ~8 CTAs, a dense 640x480 one 64+), bounded by what can be co-resident.
    void choose_team(int n_jobs, int max_points, int* team, int* n_teams) const {
        const int cap = info.max_resident_ctas;
        int t = 1;
        if (cfg.team_size) {
            t = int(cfg.team_size);
        } else if (n_jobs < cap) {
            const int by_points = max_points < 10000 ? 1 : (max_points + info.block * 8 - 1) / (info.block * 8);
            t = std::max(1, std::min(std::min(cap / n_jobs, by_points), 64));
        }
        t = std::max(1, std::min(t, std::min(kMaxTeam, cap)));
        *team = t;
        *n_teams = std::max(1, std::min(n_jobs, cap / t));
    }

    int run_align(int n_jobs, int max_points, int start = 0) {
        int team, n_teams;
        choose_team(n_jobs, max_points, &team, &n_teams);
        int rc;
        if ((rc = ensure_trace()) != VORS_OK) return rc;
        if (team > 1) CU_TRY(cudaMemsetAsync(d_scratch, 0, size_t(n_teams) * sizeof(TeamScratch), L.stream));
        const AlignParams p = make_params(n_jobs, team, start);
        CU_TRY(launch_align(L, p, n_teams));
        return VORS_OK;
    }

    // n x Tracker::track (inverse_compositional.rs:170-240)
    // `next_img` (optional, host frames): the frames the NEXT call will be given; their upload and pyramid build then
    // overlap this call's alignment.
    int track(const double* depth_ts, const uint16_t* const* depth, const uint16_t* depth_dev, const double* img_ts,
              const uint8_t* const* img, const uint8_t* img_dev, int* status, vors_track_stats* stats,
              const uint8_t* const* next_img = nullptr, const uint8_t* next_img_dev = nullptr) {
        DEVICE_SCOPE(device);
        const unsigned long long launches0 = L.launches + LC.launches;
        int rc;
        CU_TRY(cudaEventRecord(ev[0], L.stream));
        // :177 lm_model = current_frame_pose^-1 * keyframe_pose
        for (int i = 0; i < n; ++i) h_init[i] = pose_mul(pose_inverse(st[size_t(i)].cur_pose), st[size_t(i)].kf_pose);
        CU_TRY(cudaMemcpyAsync(d_init, h_init, size_t(n) * sizeof(Pose), cudaMemcpyHostToDevice, L.stream));
        int max_points = 0;
        for (int i = 0; i < n; ++i) max_points = std::max(max_points, h_n_points[i * kMaxLevels]);
        // Host frames of a large batch are processed as two half batches: the H2D copy (+ transpose) of the second half
        // runs on the copy stream while the first half is being aligned (each half still fills the device: the align
        // kernel spreads every alignment over cap / (n/2) CTAs).  SURVEY §8f rank 2.
        // were these frames announced by the previous call?  Then they (and their pyramids) are already on the device.
        bool prefetched = pending && (img_dev ? pending_dev == img_dev : (img && !pending_dev));
        for (int i = 0; prefetched && !img_dev && i < n; ++i) prefetched = (img[i] == pending_ptrs[size_t(i)]);
        pending = false;
        pending_dev = nullptr;
        if (prefetched) {
            std::swap(d_pyr, d_pyr_alt);
            std::swap(d_stage8, d_stage8_alt);
            std::swap(d_jobs, d_jobs_alt);
            CU_TRY(cudaStreamWaitEvent(L.stream, ev_pref, 0));
        }
        const int n_chunks = (!prefetched && !img_dev && n >= 64 && cfg.team_size == 0) ? 2 : 1;
        if (prefetched) {
        } else if (img_dev) {
            if ((rc = upload_images_device(img_dev)) != VORS_OK) return rc;
        } else {
            if (!img) return fail(VORS_E_INVALID, "null image pointer array");
            for (int i = 0; i < n; ++i)
                if (!img[i]) return fail(VORS_E_INVALID, "null image pointer");
            if (n_chunks == 1) {
                if ((rc = upload_images_host(img)) != VORS_OK) return rc;
            } else {
                CU_TRY(cudaEventRecord(ev_up[0], L.stream));       // the copy stream must not overtake earlier work on L
                CU_TRY(cudaStreamWaitEvent(LC.stream, ev_up[0], 0));
                for (int c = 0; c < n_chunks; ++c) {
                    const int start = c * (n / 2), m = c == 0 ? n / 2 : n - n / 2;
                    if ((rc = upload_images_host_range(img, start, m, LC)) != VORS_OK) return rc;
                    CU_TRY(cudaEventRecord(ev_up[c], LC.stream));
                }
            }
        }
        CU_TRY(cudaEventRecord(ev[1], L.stream));
        if (n_chunks == 1) {
            if (!prefetched) launch_pyramid(L, g, d_pyr, nullptr, n);  // :178
            CU_TRY(cudaEventRecord(ev[2], L.stream));
            if ((rc = run_align(n, max_points)) != VORS_OK) return rc;  // :181-201
        } else {
            CU_TRY(cudaEventRecord(ev[2], L.stream));
            for (int c = 0; c < n_chunks; ++c) {
                const int start = c * (n / 2), m = c == 0 ? n / 2 : n - n / 2;
                CU_TRY(cudaStreamWaitEvent(L.stream, ev_up[c], 0));
                launch_pyramid(L, g, d_pyr + size_t(start) * g.pix_stride, nullptr, m);
                if ((rc = run_align(m, max_points, start)) != VORS_OK) return rc;
            }
        }
        CU_TRY(cudaEventRecord(ev[3], L.stream));
        CU_TRY(cudaMemcpyAsync(h_results, d_results, size_t(n) * sizeof(AlignResult), cudaMemcpyDeviceToHost, L.stream));
        if (next_img) {
            // the other set is idle (everything earlier calls did with it has completed): stage the next frames there on the
            // copy stream while the align kernel runs
            for (int i = 0; i < n; ++i)
                if (!next_img[i]) return fail(VORS_E_INVALID, "null next-image pointer");
            if ((rc = upload_images_host_range(next_img, 0, n, LC, d_pyr_alt, d_stage8_alt)) != VORS_OK) return rc;
            launch_pyramid(LC, g, d_pyr_alt, nullptr, n);
            CU_TRY(cudaEventRecord(ev_pref, LC.stream));
            pending_ptrs.assign(next_img, next_img + n);
            pending = true;
        } else if (next_img_dev) {
            if ((rc = upload_images_device(next_img_dev, &LC, d_pyr_alt)) != VORS_OK) return rc;
            launch_pyramid(LC, g, d_pyr_alt, nullptr, n);
            CU_TRY(cudaEventRecord(ev_pref, LC.stream));
            pending_dev = next_img_dev;
            pending = true;
        }
        CU_TRY(cudaStreamSynchronize(L.stream));

        int m = 0;
        bool any_failed = false;
        unsigned long long pp = 0;
        for (int i = 0; i < n; ++i) {
            StreamState& s = st[size_t(i)];
            const AlignResult& r = h_results[i];
            s.cur_depth_ts = depth_ts ? depth_ts[i] : 0.0;  // :203-204
            s.cur_img_ts = img_ts ? img_ts[i] : 0.0;
            const bool ok = r.status == VORS_OK;
            if (ok) s.cur_pose = pose_mul(s.kf_pose, pose_inverse(r.model));  // :206-208
            any_failed |= !ok;
            const bool change = r.optical_flow >= cfg.keyframe_flow_threshold;  // :224 (NaN -> false)
            if (change) h_items[m++] = i;
            pp += r.point_passes;
            if (status) status[i] = r.status;
            if (stats) {
                vors_track_stats& o = stats[i];
                std::memset(&o, 0, sizeof(o));
                o.status = r.status;
                o.keyframe_changed = change ? 1 : 0;
                o.optical_flow = r.optical_flow;
                for (int l = 0; l < g.L; ++l) {
                    o.n_iters[l] = r.n_iters[l];
                    o.energy[l] = r.energy[l];
                    o.n_points[l] = r.n_points[l];
                }
                o.n_passes = r.n_passes;
            }
        }
        last_point_passes = pp;
        // :227-239 keyframe switch: rebuild from the frame's pyramid (still on the device) and its depth map
        if (m > 0) {
            CU_TRY(cudaMemcpyAsync(d_items, h_items, size_t(m) * 4, cudaMemcpyHostToDevice, L.stream));
            const uint16_t* depth_slab = depth_dev;
            if (!depth_dev) {
                if ((rc = upload_depth_host(depth, m)) != VORS_OK) return rc;
                depth_slab = d_depth;
            }
            if ((rc = precompute(depth_slab, m)) != VORS_OK) return rc;
            for (int j = 0; j < m; ++j) {
                StreamState& s = st[size_t(h_items[j])];
                s.kf_depth_ts = s.cur_depth_ts;
                s.kf_img_ts = s.cur_img_ts;
                s.kf_pose = s.cur_pose;
            }
        }
        CU_TRY(cudaEventRecord(ev[4], L.stream));
        CU_TRY(cudaStreamSynchronize(L.stream));
        for (int k = 0; k < 4; ++k) CU_TRY(cudaEventElapsedTime(&last_ms[k], ev[k], ev[k + 1]));
        last_launches = (L.launches + LC.launches) - launches0;
        return any_failed ? VORS_OPTIMIZATION_FAILED : VORS_OK;
    }

    int need_tmp(size_t bytes) {
        if (bytes > tmp_bytes) {
            if (d_tmp) cudaFree(d_tmp);
            d_tmp = nullptr;
            tmp_bytes = 0;
            CU_TRY(cudaMalloc(&d_tmp, bytes));
            tmp_bytes = bytes;
        }
        return VORS_OK;
    }

    int copy_trace(uint32_t stream, vors_trace_rec* out, int cap, int* len) const {
        if (stream >= uint32_t(n)) return fail(VORS_E_INVALID, "stream index out of range");
        if (!tracing || !d_trace) return fail(VORS_E_INVALID, "tracing was not enabled before the last track call");
        const int have = h_results[stream].trace_len;
        const int k = std::max(0, std::min(have, cap));
        if (k > 0 && out) {
            DEVICE_SCOPE(device);
            CU_TRY(cudaMemcpy(out, d_trace + size_t(stream) * kTraceCap, size_t(k) * sizeof(vors_trace_rec), cudaMemcpyDeviceToHost));
        }
        if (len) *len = k;
        return VORS_OK;
    }
};

static int new_engine(const vors_config* cfg, uint32_t n, uint32_t rows, uint32_t cols, int layout, Engine** out) {
    Engine* e = new (std::nothrow) Engine();
    if (!e) return fail(VORS_E_NOMEM, "out of host memory");
    const int rc = e->setup(cfg, n, rows, cols, layout);
    if (rc != VORS_OK) {
        const std::string keep = g_last_error;
        delete e;
        g_last_error = keep;
        return rc;
    }
    *out = e;
    return VORS_OK;
}

}  // namespace vors

using namespace vors;

struct vors_tracker {
    Engine* e;
};
struct vors_batch {
    Engine* e;
};
struct vors_keyframe {
    Engine* e;
};

extern "C" {

void vors_config_default(vors_config* c) {
    if (!c) return;
    std::memset(c, 0, sizeof(*c));
    c->nb_levels = 6;                  // src/bin/vors_track.rs:35
    c->candidates_diff_threshold = 7;  // :36
    c->depth_scale = 5000.0f;          // src/dataset/tum_rgbd.rs:15
    c->fx = 517.306408f;               // INTRINSICS_FR1, tum_rgbd.rs:31-35
    c->fy = 516.469215f;
    c->cx = 318.643040f;
    c->cy = 255.313989f;
    c->skew = 0.0f;
    c->idepth_variance = 0.0001f;      // vors_track.rs:39
    c->candidate_mode = VORS_CANDIDATES_COARSE_TO_FINE;
    c->fixed_iters = 0;
    c->lm_coef_init = 0.1f;            // lm_optimizer.rs:115
    c->lm_coef_reject_mult = 10.0f;    // :173
    c->lm_coef_accept_mult = 0.1f;     // :186
    c->energy_delta_stop = 1.0f;       // :179
    c->max_iters = 20;                 // :157
    c->keyframe_flow_threshold = 1.0f; // inverse_compositional.rs:224
    c->device = -1;
    c->team_size = 0;
    c->dso_nb_target = 2000;
}

const char* vors_last_error(void) { return g_last_error.c_str(); }
const char* vors_version(void) { return "vors_b200 0.1 (sm_100a)"; }

int vors_device_count(void) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    int usable = 0;
    for (int d = 0; d < count; ++d) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ++usable;
    }
    return usable;
}

// ---- tracker -----------------------------------------------------------------------------------------
int vors_tracker_create(const vors_config* cfg, double depth_ts, const uint16_t* depth, double img_ts, const uint8_t* img,
                        uint32_t rows, uint32_t cols, int layout, vors_tracker** out) {
    if (!out || !depth || !img) return fail(VORS_E_INVALID, "null argument");
    *out = nullptr;
    Engine* e = nullptr;
    int rc = new_engine(cfg, 1, rows, cols, layout, &e);
    if (rc != VORS_OK) return rc;
    rc = e->init_keyframes(&depth_ts, &depth, nullptr, &img_ts, &img, nullptr);
    if (rc != VORS_OK) {
        const std::string keep = g_last_error;
        delete e;
        g_last_error = keep;
        return rc;
    }
    *out = new vors_tracker{e};
    return VORS_OK;
}

int vors_tracker_track(vors_tracker* t, double depth_ts, const uint16_t* depth, double img_ts, const uint8_t* img,
                       vors_track_stats* stats) {
    if (!t || !img) return fail(VORS_E_INVALID, "null argument");
    return t->e->track(&depth_ts, &depth, nullptr, &img_ts, &img, nullptr, nullptr, stats);
}

int vors_tracker_current_frame(const vors_tracker* t, double* depth_ts, vors_pose* pose) {
    if (!t) return fail(VORS_E_INVALID, "null tracker");
    if (depth_ts) *depth_ts = t->e->st[0].cur_depth_ts;  // inverse_compositional.rs:245: the DEPTH timestamp
    if (pose) *pose = from_pose(t->e->st[0].cur_pose);
    return VORS_OK;
}

int vors_tracker_keyframe_pose(const vors_tracker* t, vors_pose* pose) {
    if (!t || !pose) return fail(VORS_E_INVALID, "null argument");
    *pose = from_pose(t->e->st[0].kf_pose);
    return VORS_OK;
}

int vors_tracker_set_tracing(vors_tracker* t, int enabled) {
    if (!t) return fail(VORS_E_INVALID, "null tracker");
    t->e->tracing = enabled != 0;
    return VORS_OK;
}

int vors_tracker_last_trace(const vors_tracker* t, vors_trace_rec* out, int cap, int* len) {
    if (!t) return fail(VORS_E_INVALID, "null tracker");
    return t->e->copy_trace(0, out, cap, len);
}

void vors_tracker_destroy(vors_tracker* t) {
    if (!t) return;
    delete t->e;
    delete t;
}

// ---- batch -------------------------------------------------------------------------------------------
int vors_batch_create(const vors_config* cfg, uint32_t n, const double* depth_ts, const uint16_t* const* depth,
                      const double* img_ts, const uint8_t* const* img, uint32_t rows, uint32_t cols, int layout,
                      vors_batch** out) {
    if (!out || !depth || !img) return fail(VORS_E_INVALID, "null argument");
    *out = nullptr;
    Engine* e = nullptr;
    int rc = new_engine(cfg, n, rows, cols, layout, &e);
    if (rc != VORS_OK) return rc;
    rc = e->init_keyframes(depth_ts, depth, nullptr, img_ts, img, nullptr);
    if (rc != VORS_OK) {
        const std::string keep = g_last_error;
        delete e;
        g_last_error = keep;
        return rc;
    }
    *out = new vors_batch{e};
    return VORS_OK;
}

int vors_batch_track(vors_batch* b, const double* depth_ts, const uint16_t* const* depth, const double* img_ts,
                     const uint8_t* const* img, int* status, vors_track_stats* stats) {
    if (!b) return fail(VORS_E_INVALID, "null batch");
    return b->e->track(depth_ts, depth, nullptr, img_ts, img, nullptr, status, stats);
}

int vors_batch_track_next(vors_batch* b, const double* depth_ts, const uint16_t* const* depth, const double* img_ts,
                          const uint8_t* const* img, const uint8_t* const* next_img, int* status, vors_track_stats* stats) {
    if (!b) return fail(VORS_E_INVALID, "null batch");
    return b->e->track(depth_ts, depth, nullptr, img_ts, img, nullptr, status, stats, next_img);
}

int vors_batch_track_device(vors_batch* b, const double* depth_ts, const uint16_t* depth_dev, const double* img_ts,
                            const uint8_t* img_dev, int* status, vors_track_stats* stats) {
    if (!b || !depth_dev || !img_dev) return fail(VORS_E_INVALID, "null argument");
    return b->e->track(depth_ts, nullptr, depth_dev, img_ts, nullptr, img_dev, status, stats);
}

int vors_batch_track_device_next(vors_batch* b, const double* depth_ts, const uint16_t* depth_dev, const double* img_ts,
                                 const uint8_t* img_dev, const uint8_t* next_img_dev, int* status, vors_track_stats* stats) {
    if (!b || !depth_dev || !img_dev) return fail(VORS_E_INVALID, "null argument");
    return b->e->track(depth_ts, nullptr, depth_dev, img_ts, nullptr, img_dev, status, stats, nullptr, next_img_dev);
}

int vors_batch_current_frames(const vors_batch* b, double* depth_ts, vors_pose* poses) {
    if (!b) return fail(VORS_E_INVALID, "null batch");
    for (int i = 0; i < b->e->n; ++i) {
        if (depth_ts) depth_ts[i] = b->e->st[size_t(i)].cur_depth_ts;
        if (poses) poses[i] = from_pose(b->e->st[size_t(i)].cur_pose);
    }
    return VORS_OK;
}

int vors_batch_size(const vors_batch* b) { return b ? b->e->n : fail(VORS_E_INVALID, "null batch"); }

int vors_batch_last_timing(const vors_batch* b, float ms[4]) {
    if (!b || !ms) return fail(VORS_E_INVALID, "null argument");
    for (int k = 0; k < 4; ++k) ms[k] = b->e->last_ms[k];
    return VORS_OK;
}

int vors_batch_last_counters(const vors_batch* b, uint64_t* launches, uint64_t* point_passes) {
    if (!b) return fail(VORS_E_INVALID, "null batch");
    if (launches) *launches = b->e->last_launches;
    if (point_passes) *point_passes = b->e->last_point_passes;
    return VORS_OK;
}

int vors_batch_set_tracing(vors_batch* b, int enabled) {
    if (!b) return fail(VORS_E_INVALID, "null batch");
    b->e->tracing = enabled != 0;
    return VORS_OK;
}

int vors_batch_last_trace(const vors_batch* b, uint32_t stream, vors_trace_rec* out, int cap, int* len) {
    if (!b) return fail(VORS_E_INVALID, "null batch");
    return b->e->copy_trace(stream, out, cap, len);
}

void vors_batch_destroy(vors_batch* b) {
    if (!b) return;
    delete b->e;
    delete b;
}

// ---- inner seams -------------------------------------------------------------------------------------
int vors_pyramid_shapes(uint32_t rows, uint32_t cols, uint32_t max_levels, uint32_t* out_rows, uint32_t* out_cols) {
    if (rows == 0 || cols == 0) return fail(VORS_E_INVALID, "empty image");
    return pyramid_shapes(rows, cols, std::min<uint32_t>(std::max<uint32_t>(max_levels, 1u), 32u), out_rows, out_cols);
}

// A bare engine for the stand-alone image-op entry points: levels limited to what the image allows.
static int image_engine(uint32_t rows, uint32_t cols, uint32_t max_levels, Engine** out) {
    if (rows == 0 || cols == 0) return fail(VORS_E_INVALID, "empty image");
    vors_config cfg;
    vors_config_default(&cfg);
    uint32_t lr[32], lc[32];
    int levels = pyramid_shapes(rows, cols, std::max<uint32_t>(max_levels, 1u), lr, lc);
    if (levels > kMaxLevels) return fail(VORS_E_INVALID, "more than VORS_MAX_LEVELS levels requested");
    cfg.nb_levels = uint32_t(levels);
    Engine* e = new (std::nothrow) Engine();
    if (!e) return fail(VORS_E_NOMEM, "out of host memory");
    // image ops accept any size >= 1x1; only the tracker needs a >= 3x3 coarsest level
    const int rc = e->setup(&cfg, 1, rows, cols, VORS_COL_MAJOR, /*tracker_rules=*/false);
    if (rc != VORS_OK) {
        const std::string keep = g_last_error;
        delete e;
        g_last_error = keep;
        return rc;
    }
    *out = e;
    return VORS_OK;
}

int vors_mean_pyramid(const uint8_t* img, uint32_t rows, uint32_t cols, uint32_t max_levels, uint8_t* out_concat) {
    if (!img || !out_concat) return fail(VORS_E_INVALID, "null argument");
    Engine* e = nullptr;
    int rc = image_engine(rows, cols, max_levels, &e);
    if (rc != VORS_OK) return rc;
    const uint8_t* one[1] = {img};
    rc = e->upload_images_host(one);
    if (rc == VORS_OK) {
        launch_pyramid(e->L, e->g, e->d_pyr, nullptr, 1);
        cudaError_t ce = cudaMemcpyAsync(out_concat, e->d_pyr, size_t(e->g.pix_total), cudaMemcpyDeviceToHost, e->L.stream);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->L.stream);
        if (ce != cudaSuccess) rc = fail(VORS_E_CUDA, "vors_mean_pyramid", cudaGetErrorString(ce));
    }
    const int levels = e->g.L;
    delete e;
    return rc == VORS_OK ? levels : rc;
}

int vors_gradients(const uint8_t* img, uint32_t rows, uint32_t cols, uint32_t max_levels, int16_t* gx_concat, int16_t* gy_concat,
                   uint16_t* g2_concat) {
    if (!img) return fail(VORS_E_INVALID, "null argument");
    Engine* e = nullptr;
    int rc = image_engine(rows, cols, max_levels, &e);
    if (rc != VORS_OK) return rc;
    const uint8_t* one[1] = {img};
    rc = e->upload_images_host(one);
    const int levels = e->g.L;
    if (rc == VORS_OK) {
        launch_pyramid(e->L, e->g, e->d_pyr, nullptr, 1);
        launch_gradients(e->L, e->g, 0, e->d_pyr, e->d_grad, e->d_g2, nullptr, 1);
        const size_t P = size_t(e->g.pix_total);
        std::vector<uint32_t> grad(P);
        cudaError_t ce = cudaMemcpyAsync(grad.data(), e->d_grad, P * 4, cudaMemcpyDeviceToHost, e->L.stream);
        if (ce == cudaSuccess && g2_concat) ce = cudaMemcpyAsync(g2_concat, e->d_g2, P * 2, cudaMemcpyDeviceToHost, e->L.stream);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->L.stream);
        if (ce != cudaSuccess) {
            rc = fail(VORS_E_CUDA, "vors_gradients", cudaGetErrorString(ce));
        } else {
            for (size_t i = 0; i < P; ++i) {
                if (gx_concat) gx_concat[i] = int16_t(grad[i] & 0xFFFFu);
                if (gy_concat) gy_concat[i] = int16_t(grad[i] >> 16);
            }
        }
    }
    delete e;
    return rc == VORS_OK ? levels : rc;
}

int vors_candidates_coarse_to_fine(uint16_t diff_threshold, const uint16_t* g2_concat, uint32_t rows, uint32_t cols,
                                   uint32_t n_levels, uint8_t* masks_concat) {
    if (!g2_concat || !masks_concat) return fail(VORS_E_INVALID, "null argument");
    Engine* e = nullptr;
    int rc = image_engine(rows, cols, n_levels, &e);
    if (rc != VORS_OK) return rc;
    if (uint32_t(e->g.L) != n_levels) {
        delete e;
        return fail(VORS_E_INVALID, "n_levels exceeds what the image size allows");
    }
    const size_t P = size_t(e->g.pix_total);
    cudaError_t ce = cudaMemcpyAsync(e->d_g2, g2_concat, P * 2, cudaMemcpyHostToDevice, e->L.stream);
    if (ce == cudaSuccess) {
        // coarsest level: all true (coarse_to_fine.rs:19-21)
        const size_t off = size_t(e->g.off[e->g.L - 1]);
        ce = cudaMemsetAsync(e->d_mask + off, 1, P - off, e->L.stream);
    }
    if (ce == cudaSuccess) {
        launch_c2f(e->L, e->g, diff_threshold, e->d_g2, e->d_mask, nullptr, 1);
        ce = cudaMemcpyAsync(masks_concat, e->d_mask, P, cudaMemcpyDeviceToHost, e->L.stream);
    }
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->L.stream);
    if (ce != cudaSuccess) rc = fail(VORS_E_CUDA, "vors_candidates_coarse_to_fine", cudaGetErrorString(ce));
    delete e;
    return rc;
}

int vors_candidates_dso(const uint16_t* gradients, uint32_t rows, uint32_t cols, uint32_t nb_target, uint32_t nb_iterations_left,
                        uint64_t seed, uint8_t* mask_out, int* used_random_branch) {
    if (!gradients || !mask_out || nb_target == 0) return fail(VORS_E_INVALID, "null argument");
    Engine* e = nullptr;
    int rc = image_engine(rows, cols, 1, &e);
    if (rc != VORS_OK) return rc;
    const size_t px = size_t(rows) * cols;
    uint16_t* d_g = nullptr;
    uint8_t* ws = nullptr;
    int nb = 0;
    cudaError_t ce = cudaMalloc(&d_g, px * 2);
    if (ce == cudaSuccess) ce = cudaMalloc(&ws, dso_workspace_bytes(int(rows), int(cols)));
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(d_g, gradients, px * 2, cudaMemcpyHostToDevice, e->L.stream);
    if (ce != cudaSuccess) {
        rc = fail(VORS_E_CUDA, "vors_candidates_dso", cudaGetErrorString(ce));
    } else {
        rc = dso_select_device(e->L, d_g, int(rows), int(cols), int(nb_target), int(nb_iterations_left), seed, e->d_mask, ws,
                               e->h_dso_flags, used_random_branch, &nb);
        if (rc == VORS_OK) {
            ce = cudaMemcpyAsync(mask_out, e->d_mask, px, cudaMemcpyDeviceToHost, e->L.stream);
            if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->L.stream);
            if (ce != cudaSuccess) rc = fail(VORS_E_CUDA, "vors_candidates_dso", cudaGetErrorString(ce));
        } else {
            fail(rc, rc == VORS_E_INVALID ? "DSO threshold does not fit u16 (the reference panics here) or too many regions" : "CUDA error in DSO selection");
        }
    }
    cudaStreamSynchronize(e->L.stream);
    if (d_g) cudaFree(d_g);
    if (ws) cudaFree(ws);
    delete e;
    return rc == VORS_OK ? nb : rc;
}

int vors_gradient_norms_example(const uint8_t* img, uint32_t rows, uint32_t cols, uint32_t max_levels, uint16_t* g2_concat) {
    if (!img || !g2_concat) return fail(VORS_E_INVALID, "null argument");
    Engine* e = nullptr;
    int rc = image_engine(rows, cols, max_levels, &e);
    if (rc != VORS_OK) return rc;
    const uint8_t* one[1] = {img};
    rc = e->upload_images_host(one);
    const int levels = e->g.L;
    if (rc == VORS_OK) {
        launch_pyramid(e->L, e->g, e->d_pyr, nullptr, 1);
        launch_sqnorm_direct(e->L, e->d_pyr, e->g.rows[0], e->g.cols[0], 0, e->d_g2);
        for (int l = 1; l < levels; ++l)
            launch_bloc_sqnorm(e->L, e->d_pyr + e->g.off[l - 1], e->g.rows[l - 1], e->g.rows[l], e->g.cols[l], e->d_g2 + e->g.off[l]);
        cudaError_t ce = cudaMemcpyAsync(g2_concat, e->d_g2, size_t(e->g.pix_total) * 2, cudaMemcpyDeviceToHost, e->L.stream);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->L.stream);
        if (ce != cudaSuccess) rc = fail(VORS_E_CUDA, "vors_gradient_norms_example", cudaGetErrorString(ce));
    }
    delete e;
    return rc == VORS_OK ? levels : rc;
}

int vors_keyframe_create(const vors_config* cfg, const uint16_t* depth, const uint8_t* img, uint32_t rows, uint32_t cols,
                         int layout, vors_keyframe** out) {
    if (!out || !depth || !img) return fail(VORS_E_INVALID, "null argument");
    *out = nullptr;
    Engine* e = nullptr;
    int rc = new_engine(cfg, 1, rows, cols, layout, &e);
    if (rc != VORS_OK) return rc;
    const double z = 0.0;
    rc = e->init_keyframes(&z, &depth, nullptr, &z, &img, nullptr);
    if (rc != VORS_OK) {
        const std::string keep = g_last_error;
        delete e;
        g_last_error = keep;
        return rc;
    }
    *out = new vors_keyframe{e};
    return VORS_OK;
}

int vors_keyframe_levels(const vors_keyframe* kf) { return kf ? kf->e->g.L : fail(VORS_E_INVALID, "null keyframe"); }

int vors_keyframe_n_points(const vors_keyframe* kf, uint32_t level) {
    if (!kf || level >= uint32_t(kf->e->g.L)) return fail(VORS_E_INVALID, "bad keyframe / level");
    return kf->e->h_n_points[level];
}

int vors_keyframe_points(const vors_keyframe* kf, uint32_t level, uint32_t* xy, float* idepth, int16_t* grad_xy, uint8_t* tmpl) {
    if (!kf || level >= uint32_t(kf->e->g.L)) return fail(VORS_E_INVALID, "bad keyframe / level");
    Engine* e = kf->e;
    const int np = e->h_n_points[level];
    if (np == 0) return VORS_OK;
    DEVICE_SCOPE(e->device);
    const size_t cnt = size_t(np);
    const size_t words = (cnt + kChunk - 1) / kChunk * (3 * kChunk);
    std::vector<uint32_t> blk(words), pk(cnt), gr(cnt);
    CU_TRY(cudaMemcpy(blk.data(), e->d_pts + 3 * size_t(e->g.pt_off[level]), words * 4, cudaMemcpyDeviceToHost));
    for (int i = 0; i < np; ++i) {
        pk[size_t(i)] = blk[pt_word(i, 0)];
        gr[size_t(i)] = blk[pt_word(i, 2)];
        if (idepth) std::memcpy(&idepth[i], &blk[pt_word(i, 1)], 4);
    }
    for (int i = 0; i < np; ++i) {
        if (xy) {
            xy[2 * i] = rec_x(pk[size_t(i)]);
            xy[2 * i + 1] = rec_y(pk[size_t(i)]);
        }
        if (tmpl) tmpl[i] = uint8_t(rec_tmpl(pk[size_t(i)]));
        if (grad_xy) {  // half2(gx, gy): small integers, exact
            const __half2 h = *reinterpret_cast<const __half2*>(&gr[size_t(i)]);
            grad_xy[2 * i] = int16_t(__low2float(h));
            grad_xy[2 * i + 1] = int16_t(__high2float(h));
        }
    }
    return VORS_OK;
}

int vors_keyframe_jacobians(const vors_keyframe* kf, uint32_t level, float* jac6) {
    if (!kf || !jac6 || level >= uint32_t(kf->e->g.L)) return fail(VORS_E_INVALID, "bad argument");
    Engine* e = kf->e;
    const int np = e->h_n_points[level];
    if (np == 0) return VORS_OK;
    DEVICE_SCOPE(e->device);
    int rc = e->need_tmp(size_t(np) * 24);
    if (rc != VORS_OK) return rc;
    launch_jacobians(e->L, e->d_pts + 3 * size_t(e->g.pt_off[level]), np, e->intr[level], e->d_tmp);
    CU_TRY(cudaMemcpyAsync(jac6, e->d_tmp, size_t(np) * 24, cudaMemcpyDeviceToHost, e->L.stream));
    CU_TRY(cudaStreamSynchronize(e->L.stream));
    return VORS_OK;
}

int vors_keyframe_mask0(const vors_keyframe* kf, uint8_t* mask) {
    if (!kf || !mask) return fail(VORS_E_INVALID, "null argument");
    Engine* e = kf->e;
    DEVICE_SCOPE(e->device);
    const size_t I = size_t(e->rows) * e->cols;
    if (e->cfg.candidate_mode == VORS_CANDIDATES_DENSE || (e->g.L == 1 && e->cfg.candidate_mode != VORS_CANDIDATES_DSO)) {
        std::memset(mask, 1, I);
        return VORS_OK;
    }
    CU_TRY(cudaMemcpy(mask, e->d_mask, I, cudaMemcpyDeviceToHost));
    return VORS_OK;
}

int vors_keyframe_idepth_map(const vors_keyframe* kf, uint32_t level, float* idepth) {
    if (!kf || !idepth || level >= uint32_t(kf->e->g.L)) return fail(VORS_E_INVALID, "bad argument");
    Engine* e = kf->e;
    DEVICE_SCOPE(e->device);
    CU_TRY(cudaMemcpy(idepth, e->d_idepth + e->g.off[level], size_t(e->g.rows[level]) * e->g.cols[level] * 4,
                      cudaMemcpyDeviceToHost));
    return VORS_OK;
}

void vors_keyframe_destroy(vors_keyframe* kf) {
    if (!kf) return;
    delete kf->e;
    delete kf;
}

// Run one ad-hoc job on a keyframe engine (stream 0) and read the result back.
static int run_single_job(Engine* e, int lvl_first, int lvl_last, int flow_level, int pass_only, const vors_pose* init,
                          bool want_trace) {
    DEVICE_SCOPE(e->device);
    e->fill_job(e->h_jobs[0], 0, lvl_first, lvl_last, flow_level, pass_only);
    e->h_init[0] = to_pose(*init);
    CU_TRY(cudaMemcpyAsync(e->d_jobs, e->h_jobs, sizeof(AlignJob), cudaMemcpyHostToDevice, e->L.stream));
    CU_TRY(cudaMemcpyAsync(e->d_init, e->h_init, sizeof(Pose), cudaMemcpyHostToDevice, e->L.stream));
    const bool keep = e->tracing;
    e->tracing = want_trace;
    int rc = e->run_align(1, e->h_n_points[lvl_last]);
    e->tracing = keep;
    if (rc != VORS_OK) return rc;
    CU_TRY(cudaMemcpyAsync(e->h_results, e->d_results, sizeof(AlignResult), cudaMemcpyDeviceToHost, e->L.stream));
    CU_TRY(cudaStreamSynchronize(e->L.stream));
    return VORS_OK;
}

static int fetch_trace(Engine* e, vors_trace_rec* trace, int trace_cap, int* trace_len) {
    const int k = std::max(0, std::min(e->h_results[0].trace_len, trace_cap));
    if (k > 0) CU_TRY(cudaMemcpy(trace, e->d_trace, size_t(k) * sizeof(vors_trace_rec), cudaMemcpyDeviceToHost));
    if (trace_len) *trace_len = k;
    return VORS_OK;
}

int vors_align_pass(const vors_keyframe* kf, uint32_t level, const uint8_t* image, const vors_pose* model, float* energy,
                    int32_t* n_inside, float g[6], float H[36]) {
    if (!kf || !image || !model || level >= uint32_t(kf->e->g.L)) return fail(VORS_E_INVALID, "bad argument");
    Engine* e = kf->e;
    DEVICE_SCOPE(e->device);
    const size_t cnt = size_t(e->g.rows[level]) * e->g.cols[level];
    CU_TRY(cudaMemcpyAsync(e->d_pyr + e->g.off[level], image, cnt, cudaMemcpyHostToDevice, e->L.stream));
    int rc = run_single_job(e, int(level), int(level), -1, 1, model, false);
    if (rc != VORS_OK) return rc;
    const AlignResult& r = e->h_results[0];
    if (energy) *energy = r.pass_energy;
    if (n_inside) *n_inside = r.pass_n_inside;
    if (g)
        for (int c = 0; c < 6; ++c) g[c] = r.pass_g[c];
    if (H) {
        int t = 0;
        for (int a = 0; a < 6; ++a)
            for (int b = a; b < 6; ++b, ++t) H[a * 6 + b] = H[b * 6 + a] = r.pass_H[t];
    }
    return VORS_OK;
}

int vors_align_level(const vors_keyframe* kf, uint32_t level, const uint8_t* image, const vors_pose* init, vors_pose* out,
                     int32_t* n_iter, float* energy, vors_trace_rec* trace, int trace_cap, int* trace_len) {
    if (!kf || !image || !init || level >= uint32_t(kf->e->g.L)) return fail(VORS_E_INVALID, "bad argument");
    Engine* e = kf->e;
    DEVICE_SCOPE(e->device);
    const size_t cnt = size_t(e->g.rows[level]) * e->g.cols[level];
    CU_TRY(cudaMemcpyAsync(e->d_pyr + e->g.off[level], image, cnt, cudaMemcpyHostToDevice, e->L.stream));
    int rc = run_single_job(e, int(level), int(level), -1, 0, init, trace != nullptr);
    if (rc != VORS_OK) return rc;
    const AlignResult& r = e->h_results[0];
    if (out) *out = from_pose(r.model);
    if (n_iter) *n_iter = r.n_iters[level];
    if (energy) *energy = r.energy[level];
    if (trace && (rc = fetch_trace(e, trace, trace_cap, trace_len)) != VORS_OK) return rc;
    return r.status;
}

int vors_align(const vors_keyframe* kf, const uint8_t* img, int layout, const vors_pose* init, vors_pose* out,
               vors_track_stats* stats, vors_trace_rec* trace, int trace_cap, int* trace_len) {
    if (!kf || !img || !init) return fail(VORS_E_INVALID, "bad argument");
    Engine* e = kf->e;
    DEVICE_SCOPE(e->device);
    const int keep_layout = e->layout;
    e->layout = layout;
    const uint8_t* one[1] = {img};
    int rc = e->upload_images_host(one);
    e->layout = keep_layout;
    if (rc != VORS_OK) return rc;
    launch_pyramid(e->L, e->g, e->d_pyr, nullptr, 1);
    rc = run_single_job(e, e->g.L - 1, 0, e->g.L - 1, 0, init, trace != nullptr);
    if (rc != VORS_OK) return rc;
    const AlignResult& r = e->h_results[0];
    if (out) *out = from_pose(r.model);
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        stats->status = r.status;
        stats->optical_flow = r.optical_flow;
        stats->keyframe_changed = r.optical_flow >= e->cfg.keyframe_flow_threshold ? 1 : 0;
        for (int l = 0; l < e->g.L; ++l) {
            stats->n_iters[l] = r.n_iters[l];
            stats->energy[l] = r.energy[l];
            stats->n_points[l] = r.n_points[l];
        }
        stats->n_passes = r.n_passes;
    }
    if (trace && (rc = fetch_trace(e, trace, trace_cap, trace_len)) != VORS_OK) return rc;
    return r.status;
}

// se3::log / so3::exp / so3::log (src/math/se3.rs:99-130, so3.rs:61-99) on the device: utilities for trajectory error
// metrics, not on the tracking path.
static int run_lie(int op, const float* in, int n_in, float* out, int n_out) {
    if (!in || !out) return fail(VORS_E_INVALID, "null argument");
    if (vors_device_count() == 0) return fail(VORS_E_CUDA, "no CUDA device available (libvors_b200 has no CPU fallback)");
    float* d = nullptr;
    CU_TRY(cudaMalloc(&d, 16 * sizeof(float)));
    Launcher L{};
    cudaError_t ce = cudaMemcpy(d, in, size_t(n_in) * sizeof(float), cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) {
        launch_lie(L, op, d, d + 8);
        ce = cudaMemcpy(out, d + 8, size_t(n_out) * sizeof(float), cudaMemcpyDeviceToHost);
    }
    cudaFree(d);
    if (ce != cudaSuccess) return fail(VORS_E_CUDA, "lie-group utility", cudaGetErrorString(ce));
    return VORS_OK;
}
int vors_se3_log(const vors_pose* pose, float xi[6]) {
    if (!pose) return fail(VORS_E_INVALID, "null argument");
    const float in[7] = {pose->t[0], pose->t[1], pose->t[2], pose->q[0], pose->q[1], pose->q[2], pose->q[3]};
    return run_lie(1, in, 7, xi, 6);
}
int vors_so3_exp(const float w[3], float q[4]) { return run_lie(2, w, 3, q, 4); }
int vors_so3_log(const float q[4], float w[3]) { return run_lie(3, q, 4, w, 3); }

int vors_se3_exp(const float xi[6], vors_pose* out) {
    if (!xi || !out) return fail(VORS_E_INVALID, "null argument");
    if (vors_device_count() == 0) return fail(VORS_E_CUDA, "no CUDA device available (libvors_b200 has no CPU fallback)");
    float* d = nullptr;
    CU_TRY(cudaMalloc(&d, 6 * sizeof(float) + sizeof(Pose)));
    Launcher L{};
    Launcher LC{};                   // copy stream: H2D + transposes of the next chunk overlap the align kernel
    cudaEvent_t ev_up[2]{};
    L.stream = nullptr;
    cudaError_t ce = cudaMemcpy(d, xi, 6 * sizeof(float), cudaMemcpyHostToDevice);
    Pose p{};
    if (ce == cudaSuccess) {
        launch_se3_exp(L, d, reinterpret_cast<Pose*>(d + 6));
        ce = cudaMemcpy(&p, d + 6, sizeof(Pose), cudaMemcpyDeviceToHost);
    }
    cudaFree(d);
    if (ce != cudaSuccess) return fail(VORS_E_CUDA, "vors_se3_exp", cudaGetErrorString(ce));
    *out = from_pose(p);
    return VORS_OK;
}

}  // extern "C"
