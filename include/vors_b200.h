/*
 * vors_b200.h — C ABI of libvors_b200.so: the B200-native (sm_100a) implementation of the
 * direct RGB-D image-alignment hot path of mpizenberg/visual-odometry-rs ("vors").
 *
 * The reference has NO FFI / plugin boundary (README.md:63 lists "Making a C FFI" as future
 * work), so the boundary sits at the public Rust API of its Tracker plus the inner seams the
 * Tracker itself calls.  Each entry point below names the reference interface it replaces
 * (path:line under /root/reference); INTEGRATION.md shows the Rust `extern "C"` binding a
 * maintainer would add.  Plain pointers and sizes only; no torch / C++ types.
 *
 * Conventions
 *  - Images are u8 gray, depth maps u16 (TUM: 5000 = 1 m, 0 = unknown), shape (rows = height,
 *    cols = width).  `layout` says how the caller's buffer is ordered: VORS_COL_MAJOR is what
 *    nalgebra's DMatrix::as_slice() yields (element (r,c) at [c*rows + r]); VORS_ROW_MAJOR is
 *    decoder output (src/misc/interop.rs:53-56 converts it with from_row_slice).
 *  - Inner-seam functions exchange COLUMN-MAJOR buffers; "concat" buffers hold the pyramid
 *    levels back to back, finest first (use vors_pyramid_shapes for the sizes).
 *  - Input pointers are HOST memory borrowed for the duration of the call unless the function
 *    name ends in `_device`.
 *  - Return value: 0 OK, 1 optimisation failed (Cholesky; tracker state updated exactly like the
 *    reference, src/core/track/inverse_compositional.rs:191-208), negative = misuse / runtime
 *    error (see VORS_E_*).  Nothing aborts.  vors_last_error() gives a thread-local message.
 *  - A handle is not thread-safe; distinct handles are independent.  There is no CPU fallback:
 *    every compute entry point fails with VORS_E_CUDA when no sm_100 device is usable.
 */
#ifndef VORS_B200_H
#define VORS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VORS_MAX_LEVELS 8

#define VORS_OK 0
#define VORS_OPTIMIZATION_FAILED 1
#define VORS_E_INVALID (-1) /* bad argument / pyramid shorter than nb_levels / image too large */
#define VORS_E_CUDA (-2)    /* CUDA runtime error or no usable device */
#define VORS_E_NOMEM (-3)

#define VORS_COL_MAJOR 0
#define VORS_ROW_MAJOR 1

#define VORS_CANDIDATES_COARSE_TO_FINE 0 /* reference: src/core/candidates/coarse_to_fine.rs */
#define VORS_CANDIDATES_DENSE 1          /* extension: every pixel with depth != 0 */
#define VORS_CANDIDATES_DSO 2            /* extension (BASELINE config 3): src/core/candidates/dso.rs at level 0 */

/* Replaces `track::Config` (src/core/track/inverse_compositional.rs:37-49) plus the constants the
 * reference hard-codes (src/core/track/lm_optimizer.rs:115,157,173,179,186; inverse_compositional.rs:224).
 * Always initialise with vors_config_default() (values of src/bin/vors_track.rs:34-40), then edit. */
/* inverse_depth.rs:81-98 `strategy_dso_mean` (Tracker default) / :105-152 `strategy_statistically_similar` (values whose
 * squared distance to the fused inverse depth reaches the fused variance discard the bloc; a discarded bloc is not a candidate) */
enum { VORS_FUSION_DSO_MEAN = 0, VORS_FUSION_STATISTICALLY_SIMILAR = 1 };
/* Scharr: gx = (3 (I[r-1,c+1] - I[r-1,c-1]) + 10 (I[r,c+1] - I[r,c-1]) + 3 (I[r+1,c+1] - I[r+1,c-1])) / 32, gy likewise along rows,
 * i16 division truncating toward zero, 1-px border 0: the same scale as the centred difference it replaces. */
enum { VORS_GRADIENT_REFERENCE = 0, VORS_GRADIENT_SCHARR = 1 };

typedef struct vors_config {
    uint32_t nb_levels;                 /* Config::nb_levels */
    uint32_t candidates_diff_threshold; /* Config::candidates_diff_threshold (u16 range) */
    float depth_scale;                  /* Config::depth_scale */
    float fx, fy, cx, cy, skew;         /* Config::intrinsics (src/core/camera.rs:84-91) */
    float idepth_variance;              /* Config::idepth_variance */
    /* ---- extensions; defaults reproduce the reference ---- */
    uint32_t candidate_mode;       /* VORS_CANDIDATES_* */
    uint32_t fixed_iters;          /* 0 = reference's adaptive stop rule; k = exactly k LM rounds/level */
    float lm_coef_init;            /* 0.1  */
    float lm_coef_reject_mult;     /* 10   */
    float lm_coef_accept_mult;     /* 0.1  */
    float energy_delta_stop;       /* 1.0  */
    uint32_t max_iters;            /* 20   */
    float keyframe_flow_threshold; /* 1.0 px at the coarsest level */
    int32_t device;                /* CUDA ordinal; -1 = current device */
    uint32_t team_size;            /* CTAs cooperating on one alignment; 0 = auto */
    uint32_t dso_nb_target;        /* VORS_CANDIDATES_DSO: target number of level-0 candidates (2000) */
    uint32_t idepth_fusion;        /* VORS_FUSION_*: how 2x2 blocs of inverse depths merge up the pyramid (multires::halve with
                                      inverse_depth::fuse); 0 = strategy_dso_mean, what the Tracker uses (inverse_compositional.rs:135-138) */
    float huber_delta;             /* > 0: Huber-weighted residuals (the north_star's extra; the reference has no robust weighting,
                                      lm_optimizer.rs:94-100): energy = mean rho_delta(r), g = sum w J r, H = sum w J J^T with
                                      w = min(1, delta / |r|) (grey levels).  0 = the reference's plain L2.  Excluded from parity
                                      with the reference; checked against the oracle's same option. */
    uint32_t gradient_operator;    /* VORS_GRADIENT_*: 0 = the Tracker's recipe (centred differences at level 0, 2x2-bloc differences of
                                      the finer image above, inverse_compositional.rs:112-117); 1 = 3x3 Scharr on every level's own
                                      image (the north_star's other extra; not in the reference, excluded from parity with it) */
} vors_config;

/* Replaces `Iso3 = Isometry3<f32>` (src/misc/type_aliases.rs:30); printed by the reference as
 * `tx ty tz qx qy qz qw` (src/dataset/tum_rgbd.rs:78-85). */
typedef struct vors_pose {
    float t[3];
    float q[4]; /* x y z w */
} vors_pose;

/* One record per energy evaluation inside the LM loop (debug / parity tracing). */
typedef struct vors_trace_rec {
    int32_t level;
    int32_t iter; /* 0 = init evaluation of the level */
    float energy;
    int32_t n_inside;
    float lm_coef;
    int32_t accepted;
} vors_trace_rec;

/* What the reference prints to stderr per frame (inverse_compositional.rs:196,222,229), as data. */
typedef struct vors_track_stats {
    int32_t status; /* VORS_OK / VORS_OPTIMIZATION_FAILED */
    int32_t keyframe_changed;
    float optical_flow;
    int32_t n_iters[VORS_MAX_LEVELS];
    float energy[VORS_MAX_LEVELS];
    int32_t n_points[VORS_MAX_LEVELS];
    int32_t n_passes; /* energy evaluations executed on the device for this frame */
    int32_t reserved;
} vors_track_stats;

void vors_config_default(vors_config* cfg);
const char* vors_last_error(void);
const char* vors_version(void);
/* "src=<sha256/16 of the CUDA sources and headers the library was compiled from> built=<UTC time> arch=sm_100a":
 * lets a caller (and __graft_entry__.build() / smoke()) prove which sources the loaded binary came from. */
const char* vors_build_info(void);
/* Number of usable sm_100 devices (0 when none; never fails). */
int vors_device_count(void);

/* ---------------------------------------------------------------------------------------------
 * Outer seam: the Tracker (one RGB-D stream).
 * ------------------------------------------------------------------------------------------- */
typedef struct vors_tracker vors_tracker;

/* Replaces `Config::init(self, f64, &DMatrix<u16>, f64, DMatrix<u8>) -> Tracker`
 * (inverse_compositional.rs:74-100; called at src/bin/vors_track.rs:46). */
int vors_tracker_create(const vors_config* cfg, double depth_ts, const uint16_t* depth, double img_ts,
                        const uint8_t* img, uint32_t rows, uint32_t cols, int layout, vors_tracker** out);

/* Replaces `Tracker::track(&mut self, f64, &DMatrix<u16>, f64, DMatrix<u8>)`
 * (inverse_compositional.rs:170-240; called at vors_track.rs:54).  `stats` may be NULL. */
int vors_tracker_track(vors_tracker* t, double depth_ts, const uint16_t* depth, double img_ts,
                       const uint8_t* img, vors_track_stats* stats);

/* Replaces `Tracker::current_frame(&self) -> (f64, Iso3)` (inverse_compositional.rs:243-248). */
int vors_tracker_current_frame(const vors_tracker* t, double* depth_ts, vors_pose* pose);
int vors_tracker_keyframe_pose(const vors_tracker* t, vors_pose* pose);
/* Copy out the LM trace of the last track() call (tracing must be enabled before it). */
int vors_tracker_set_tracing(vors_tracker* t, int enabled);
int vors_tracker_last_trace(const vors_tracker* t, vors_trace_rec* out, int cap, int* len);
void vors_tracker_destroy(vors_tracker* t);

/* ---------------------------------------------------------------------------------------------
 * Batched outer seam: n independent RGB-D streams (n Trackers) advanced by one frame per call in
 * ONE persistent device launch.  Semantically n x the calls above; this is the throughput path
 * (one alignment per CTA team) and the unit sharded across GPUs (one batch per rank).
 * ------------------------------------------------------------------------------------------- */
typedef struct vors_batch vors_batch;

/* n x Config::init.  depth[i] / img[i] are per-stream host pointers. */
int vors_batch_create(const vors_config* cfg, uint32_t n, const double* depth_ts, const uint16_t* const* depth,
                      const double* img_ts, const uint8_t* const* img, uint32_t rows, uint32_t cols,
                      int layout, vors_batch** out);
/* n x Tracker::track with HOST buffers (H2D copies happen inside the call).  `status` (n ints) and
 * `stats` (n records) may be NULL.  Returns 0 if every stream is OK, 1 if any stream's optimisation
 * failed (see status[i]), negative on error. */
int vors_batch_track(vors_batch* b, const double* depth_ts, const uint16_t* const* depth,
                     const double* img_ts, const uint8_t* const* img, int* status, vors_track_stats* stats);
/* Same, and `next_img` (n host pointers, or NULL) announces the frames the NEXT call will be given: their upload and
 * pyramid build (multires.rs:21-31, run by Tracker::track at inverse_compositional.rs:178) are started on a copy
 * stream so that they overlap this call's alignment; the next call recognises them by pointer identity and skips its
 * own upload.  The announced buffers must stay valid and unchanged until that call returns.  An extension: the
 * reference's track() is strictly sequential. */
int vors_batch_track_next(vors_batch* b, const double* depth_ts, const uint16_t* const* depth,
                          const double* img_ts, const uint8_t* const* img, const uint8_t* const* next_img,
                          int* status, vors_track_stats* stats);
/* Same with DEVICE-resident inputs in the internal layout (column-major, one image per stream,
 * `img_dev` = n*rows*cols u8 contiguous, `depth_dev` = n*rows*cols u16 contiguous). */
int vors_batch_track_device(vors_batch* b, const double* depth_ts, const uint16_t* depth_dev,
                            const double* img_ts, const uint8_t* img_dev, int* status,
                            vors_track_stats* stats);
/* n x Tracker::current_frame. */
/* Device-resident variant of vors_batch_track_next: `next_img_dev` (or NULL) is the device buffer the next call will pass as
 * `img_dev`; its copy into the frame pyramids and the pyramid build overlap this call's alignment. */
int vors_batch_track_device_next(vors_batch* b, const double* depth_ts, const uint16_t* depth_dev,
                                 const double* img_ts, const uint8_t* img_dev, const uint8_t* next_img_dev,
                                 int* status, vors_track_stats* stats);
/* Forgets the frames announced by the last vors_batch_track_next / _device_next call and waits for their copy to finish:
 * afterwards the announced buffers may be reused, refilled or freed.  Announced frames are recognised in the next call by
 * POINTER identity and are read asynchronously until that call returns: a caller that refills the same buffers in place
 * (ring buffer, in-place decode) must cancel the announcement first, or not announce. */
int vors_batch_cancel_prefetch(vors_batch* b);
int vors_batch_current_frames(const vors_batch* b, double* depth_ts, vors_pose* poses);
int vors_batch_size(const vors_batch* b);
/* Device time of the last track call's kernels, by stage (ms; CUDA events on the batch's stream):
 * [0] upload/transposes [1] pyramid [2] align (persistent LM kernel) [3] keyframe rebuilds. */
int vors_batch_last_timing(const vors_batch* b, float ms[4]);
/* Kernel launches issued by the last track call and the sum of candidate-point evaluations
 * (points x passes) the align kernel executed — feeds bench.py's roofline. */
int vors_batch_last_counters(const vors_batch* b, uint64_t* launches, uint64_t* point_passes);
/* Shape of the last align launch: CTAs cooperating on one alignment (`team`; 1 = the throughput configuration chosen when
 * the batch fills the device) and alignments in flight at once (`n_teams`). */
int vors_batch_last_launch_shape(const vors_batch* b, int* team, int* n_teams);
int vors_batch_set_tracing(vors_batch* b, int enabled);
int vors_batch_last_trace(const vors_batch* b, uint32_t stream, vors_trace_rec* out, int cap, int* len);
void vors_batch_destroy(vors_batch* b);

/* ---------------------------------------------------------------------------------------------
 * Inner seams (kernel-level parity tests; also what a Rust shim that keeps `Tracker` intact
 * would bind).  Host buffers, column-major.
 * ------------------------------------------------------------------------------------------- */

/* Level shapes of `multires::limited_sequence` + `halve` (src/core/multires.rs:38-88): halves with
 * floor, stops at max_levels or when a side would become 0.  Returns the number of levels. */
int vors_pyramid_shapes(uint32_t rows, uint32_t cols, uint32_t max_levels, uint32_t* out_rows,
                        uint32_t* out_cols);

/* Replaces `multires::mean_pyramid(max_levels, DMatrix<u8>) -> Vec<DMatrix<u8>>` (multires.rs:21-31).
 * out_concat receives all levels (level 0 is a copy of img).  Returns the number of levels or <0. */
int vors_mean_pyramid(const uint8_t* img, uint32_t rows, uint32_t cols, uint32_t max_levels,
                      uint8_t* out_concat);

/* Replaces the Tracker's gradient recipe (inverse_compositional.rs:112-117): level 0
 * `gradient::centered` (src/core/gradient.rs:15-33), levels >= 1 `multires::gradients_xy`
 * (multires.rs:112-126, bloc_x / bloc_y gradient.rs:74-93), and `gradient::squared_norm`
 * (gradient.rs:38-44) of each.  Any output may be NULL. */
int vors_gradients(const uint8_t* img, uint32_t rows, uint32_t cols, uint32_t max_levels,
                   int16_t* gx_concat, int16_t* gy_concat, uint16_t* g2_concat);

/* Replaces `candidates::coarse_to_fine::select(diff_threshold, &[DMatrix<u16>]) -> Vec<DMatrix<bool>>`
 * (coarse_to_fine.rs:15-32).  g2_concat finest first; masks_concat (0/1 bytes) finest first. */
int vors_candidates_coarse_to_fine(uint16_t diff_threshold, const uint16_t* g2_concat, uint32_t rows,
                                   uint32_t cols, uint32_t n_levels, uint8_t* masks_concat);

/* Replaces `candidates::dso::select(&DMatrix<u16>, DEFAULT_REGION_CONFIG, DEFAULT_BLOCK_CONFIG, RecursiveConfig{
 * nb_iterations_left, ..DEFAULT}, nb_target) -> DMatrix<bool>` (src/core/candidates/dso.rs:98-150).  `gradients` is the
 * column-major u16 gradient-magnitude map.  The reference's thinning branch draws from thread_rng (dso.rs:140-143,
 * not reproducible); here the r-th picked pixel (column-major order) uses the r-th output of splitmix64(seed).
 * Returns the number of block candidates of the last recursion (before thinning) or <0; VORS_E_INVALID also stands
 * for the reference's `expect("woops")` panic (threshold does not fit u16). */
int vors_candidates_dso(const uint16_t* gradients, uint32_t rows, uint32_t cols, uint32_t nb_target,
                        uint32_t nb_iterations_left, uint64_t seed, uint8_t* mask_out, int* used_random_branch);

/* Replaces the example gradient-norm recipe: level 0 `gradient::squared_norm_direct`
 * (src/core/gradient.rs:49-65), levels >= 1 `multires::gradients_squared_norm` (multires.rs:96-106 with
 * gradient::bloc_squared_norm, gradient.rs:102-111) — what examples/candidates_coarse-to-fine.rs:55-69 feeds the
 * selector; numerically different from vors_gradients (no intermediate truncation).  Returns the number of levels. */
int vors_gradient_norms_example(const uint8_t* img, uint32_t rows, uint32_t cols, uint32_t max_levels,
                                uint16_t* g2_concat);

/* Replaces `precompute_multires_data` (inverse_compositional.rs:105-161): keyframe precompute.
 * A keyframe handle owns device scratch (the frame pyramid slot, job descriptors) that vors_align_pass / vors_align_level /
 * vors_align overwrite: the `const` in their signatures says that the keyframe's CANDIDATES are not modified, not that the
 * calls are re-entrant.  One caller at a time per handle, like every other handle of this library. */
typedef struct vors_keyframe vors_keyframe;
int vors_keyframe_create(const vors_config* cfg, const uint16_t* depth, const uint8_t* img, uint32_t rows,
                         uint32_t cols, int layout, vors_keyframe** out);
int vors_keyframe_levels(const vors_keyframe* kf);
int vors_keyframe_n_points(const vors_keyframe* kf, uint32_t level);
/* `usable_candidates_multires[level]` (extract_z, inverse_compositional.rs:260-279) in the reference's
 * column-major scan order, plus the per-candidate gradient and template value the kernel reads.
 * Any output may be NULL.  xy: 2 u32 per point (x = col, y = row). */
int vors_keyframe_points(const vors_keyframe* kf, uint32_t level, uint32_t* xy, float* idepth,
                         int16_t* grad_xy, uint8_t* tmpl);
/* `jacobians_multires[level]` (warp_jacobians, inverse_compositional.rs:284-341), 6 f32 per point,
 * evaluated on the device by the same code the align kernel uses. */
int vors_keyframe_jacobians(const vors_keyframe* kf, uint32_t level, float* jac6);
int vors_keyframe_mask0(const vors_keyframe* kf, uint8_t* mask);
/* idepth pyramid map (inverse_depth.rs:24-29, 49-98): NaN where unknown. */
int vors_keyframe_idepth_map(const vors_keyframe* kf, uint32_t level, float* idepth);
void vors_keyframe_destroy(vors_keyframe* kf);

/* One evaluation of `eval_energy` + `compute_eval_data` (lm_optimizer.rs:68-107) at `model` on one
 * level: energy = sum r^2 / n_inside, g = sum J r, H = sum J J^T (full 6x6, row-major = symmetric).
 * `image` is the current frame's pyramid level (column-major, the keyframe's level shape). */
int vors_align_pass(const vors_keyframe* kf, uint32_t level, const uint8_t* image, const vors_pose* model,
                    float* energy, int32_t* n_inside, float g[6], float H[36]);

/* Replaces `LMOptimizerState::iterative_solve(&Obs, Iso3)` (src/math/optimizer.rs:57-70 with
 * lm_optimizer.rs:113-192) on one level.  Returns VORS_OK / VORS_OPTIMIZATION_FAILED / <0. */
int vors_align_level(const vors_keyframe* kf, uint32_t level, const uint8_t* image, const vors_pose* init,
                     vors_pose* out, int32_t* n_iter, float* energy, vors_trace_rec* trace, int trace_cap,
                     int* trace_len);

/* The level loop of `Tracker::track` (inverse_compositional.rs:181-201) for one frame against a
 * keyframe: builds the frame's pyramid and runs all levels coarse to fine on the device.
 * `img` is the full-resolution frame in `layout`. */
int vors_align(const vors_keyframe* kf, const uint8_t* img, int layout, const vors_pose* init, vors_pose* out,
               vors_track_stats* stats, vors_trace_rec* trace, int trace_cap, int* trace_len);

/* `se3::exp` (src/math/se3.rs:65-95) evaluated by the device code the LM step uses. */
int vors_se3_exp(const float xi[6], vors_pose* out);
/* `se3::log` (src/math/se3.rs:99-130), `so3::exp` / `so3::log` (src/math/so3.rs:61-99; quaternion as x y z w): utilities for
 * trajectory error metrics, evaluated on the device in f32 like the reference.  Not on the tracking path. */
int vors_se3_log(const vors_pose* pose, float xi[6]);
int vors_so3_exp(const float w[3], float q[4]);
int vors_so3_log(const float q[4], float w[3]);

#ifdef __cplusplus
}
#endif
#endif /* VORS_B200_H */
