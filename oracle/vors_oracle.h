/*
 * vors_oracle.h — C interface of the CPU ORACLE (test infrastructure, NOT product code).
 *
 * The oracle is a single-threaded C++17 restatement of the hot path of
 * mpizenberg/visual-odometry-rs ("vors", /root/reference, commit 5f77605): the
 * inverse-compositional direct RGB-D alignment in src/core/track/ and the
 * pyramid / gradient / candidate / inverse-depth precompute it consumes.
 * Every function cites the reference file:line it follows (see vors_oracle.cpp).
 *
 * PARITY STATUS: "parity unpinned" for SURVEY.md §8a rows A-O, Q-S — the
 * reference ships no tests, fixtures or golden vectors for them, and the Rust
 * toolchain (cargo/rustc) plus its un-vendored dependency nalgebra ^0.17 are not
 * available in this image, so the reference itself cannot be run here.  What IS
 * pinned: the three `prune_with_thresh` doc-comment vectors
 * (src/core/candidates/coarse_to_fine.rs:68-71) and the so3/se3 properties
 * restated from src/math/so3.rs:115-142 and src/math/se3.rs:145-173 (row P).
 * nalgebra arithmetic (6x6 Cholesky, quaternion algebra, Isometry3 products) is
 * restated from its published 0.17 behaviour; see the comments in the .cpp.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product (libvors_b200.so) never
 * links, includes or calls anything under oracle/.
 *
 * Conventions (same as the reference): images are COLUMN-MAJOR, shape
 * (rows = height, cols = width), element (row, col) at [col * rows + row];
 * candidate coordinates are (x = col, y = row); everything is f32.
 * "concat" buffers hold pyramid levels back to back, finest first.
 */
#ifndef VORS_ORACLE_H
#define VORS_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define REF_MAX_LEVELS 16

/* Same field layout as vors_config in include/vors_b200.h so one ctypes
 * structure serves both; fields the oracle does not use are ignored. */
typedef struct ref_config {
    uint32_t nb_levels;
    uint32_t candidates_diff_threshold; /* u16 range */
    float depth_scale;
    float fx, fy, cx, cy, skew;
    float idepth_variance;
    uint32_t candidate_mode; /* 0 coarse-to-fine (reference), 1 dense (extension), 2 dso at level 0 (extension) */
    uint32_t fixed_iters;    /* 0 = reference's adaptive stop rule; k>0 = exactly k step+eval rounds (benchmark extension) */
    float lm_coef_init;
    float lm_coef_reject_mult;
    float lm_coef_accept_mult;
    float energy_delta_stop;
    uint32_t max_iters;
    float keyframe_flow_threshold;
    int32_t device;     /* unused by the oracle */
    uint32_t team_size; /* unused by the oracle */
    uint32_t dso_nb_target;
    uint32_t idepth_fusion; /* 0 strategy_dso_mean (Tracker), 1 strategy_statistically_similar (inverse_depth.rs:105-152) */
    float huber_delta;      /* > 0: Huber weights (extension, not in the reference); 0 = plain L2 like the reference */
    uint32_t gradient_operator; /* 0 the Tracker's recipe; 1 Scharr 3x3 on every level (extension, not in the reference) */
} ref_config;

typedef struct ref_pose {
    float t[3];
    float q[4]; /* x y z w */
} ref_pose;

/* One record per energy evaluation of the LM loop (init evaluation has iter 0). */
typedef struct ref_trace_rec {
    int32_t level;
    int32_t iter;
    float energy;
    int32_t n_inside;
    float lm_coef;    /* coefficient the step that produced this model used (0.1 for init) */
    int32_t accepted; /* 1 accepted / init, 0 rejected */
} ref_trace_rec;

typedef struct ref_track_stats {
    int32_t status; /* 0 ok, 1 optimisation failed (Cholesky) */
    int32_t keyframe_changed;
    float optical_flow;
    int32_t n_iters[REF_MAX_LEVELS];
    float energy[REF_MAX_LEVELS];
    int32_t n_points[REF_MAX_LEVELS];
} ref_track_stats;

void ref_config_default(ref_config* cfg);
/* Test-only: 1 = accumulate E / g / H in f64 (calling thread only); 0 = the reference's sequential f32 (default). */
void ref_set_accum_f64(int on);

/* ---- rows A-D, S: pyramid and gradients ------------------------------------------- */
int ref_pyramid_shapes(int rows, int cols, int max_levels, int* out_rows, int* out_cols);
int ref_mean_pyramid(const uint8_t* img, int rows, int cols, int max_levels, uint8_t* out_concat);
void ref_gradient_centered(const uint8_t* img, int rows, int cols, int16_t* gx, int16_t* gy);
/* Extension (gradient_operator = 1, not in the reference): 3x3 Scharr / 32, truncating, 1-px border 0. */
void ref_gradient_scharr(const uint8_t* img, int rows, int cols, int16_t* gx, int16_t* gy);
/* Tracker recipe: level 0 centered, level l>=1 2x2-block gradients of level l-1. */
void ref_gradients_tracker(const uint8_t* pyr_concat, int rows, int cols, int n_levels,
                           int16_t* gx_concat, int16_t* gy_concat, uint16_t* g2_concat);
void ref_squared_norm_direct(const uint8_t* img, int rows, int cols, uint16_t* out);
/* Example recipe (row S): level 0 squared_norm_direct, level l>=1 bloc_squared_norm of level l-1. */
void ref_gradients_squared_norm_example(const uint8_t* pyr_concat, int rows, int cols, int n_levels,
                                        uint16_t* g2_concat);

/* ---- row E: coarse-to-fine candidates ----------------------------------------------- */
void ref_prune_with_thresh(uint16_t thresh, uint16_t a, uint16_t b, uint16_t c, uint16_t d,
                           uint8_t out[4]);
/* g2_concat / masks_concat are finest-first; masks hold 0/1 bytes for every level. */
void ref_c2f_select(uint16_t thresh, const uint16_t* g2_concat, int rows, int cols, int n_levels,
                    uint8_t* masks_concat);

/* ---- row R: DSO candidates (deterministic branches; random branch seeded) ----------- */
/* DEFAULT_* configs of dso.rs:72-90; nb_iterations_left is 1 by default, 2 in examples/candidates_dso.rs:46.
 * Returns the number of block candidates of the last recursion (before random thinning), <0 on the
 * reference's "woops" panic. */
int ref_dso_select(const uint16_t* gradients, int rows, int cols, int nb_target, int nb_iterations_left,
                   uint64_t seed, uint8_t* mask_out, int* used_random_branch);

/* ---- rows F-K: keyframe precompute -------------------------------------------------- */
typedef struct ref_keyframe ref_keyframe;
ref_keyframe* ref_keyframe_create(const ref_config* cfg, const uint16_t* depth, const uint8_t* img,
                                  int rows, int cols);
void ref_keyframe_destroy(ref_keyframe* kf);
int ref_keyframe_levels(const ref_keyframe* kf);
void ref_keyframe_level_shape(const ref_keyframe* kf, int level, int* rows, int* cols);
void ref_keyframe_intrinsics(const ref_keyframe* kf, int level, float out5[5]); /* fx fy cx cy skew */
int ref_keyframe_n_points(const ref_keyframe* kf, int level);
void ref_keyframe_points(const ref_keyframe* kf, int level, uint32_t* xy /*2n*/, float* idepth /*n*/,
                         float* jac /*6n or NULL*/);
void ref_keyframe_image(const ref_keyframe* kf, int level, uint8_t* out);
void ref_keyframe_mask0(const ref_keyframe* kf, uint8_t* out);
/* idepth map of a level: NaN where unknown; weight 0 where unknown. */
void ref_keyframe_idepth_map(const ref_keyframe* kf, int level, float* idepth, float* weight);

/* ---- rows M, N: one evaluation -------------------------------------------------------
 * accum: 0 = reference-faithful sequential f32 sums, 1 = f64 sums (used to grade the GPU
 * reduction independently of the reference's own f32 round-off).
 * H is the full 6x6 (36 floats, symmetric).  Returns n_inside. */
int ref_eval(const ref_keyframe* kf, int level, const uint8_t* image, int rows, int cols,
             const ref_pose* model, int accum, float* energy, float g[6], float H[36]);

/* ---- rows O, P: LM loop on one level -------------------------------------------------
 * returns 0 ok, 1 Cholesky failure. */
int ref_iterative_solve(const ref_config* cfg, const ref_keyframe* kf, int level, const uint8_t* image,
                        int rows, int cols, const ref_pose* init, ref_pose* out, int* n_iter,
                        float* final_energy, ref_trace_rec* trace, int trace_cap, int* trace_len);

/* ---- row Q: tracker ------------------------------------------------------------------ */
typedef struct ref_tracker ref_tracker;
/* layout: 0 = column-major (nalgebra as_slice), 1 = row-major (decoder output). */
ref_tracker* ref_tracker_create(const ref_config* cfg, double depth_ts, const uint16_t* depth,
                                double img_ts, const uint8_t* img, int rows, int cols, int layout);
int ref_tracker_track(ref_tracker* t, double depth_ts, const uint16_t* depth, double img_ts,
                      const uint8_t* img, ref_track_stats* stats, ref_trace_rec* trace, int trace_cap,
                      int* trace_len);
void ref_tracker_current_frame(const ref_tracker* t, double* depth_ts, ref_pose* pose);
void ref_tracker_keyframe_pose(const ref_tracker* t, ref_pose* pose);
const ref_keyframe* ref_tracker_keyframe(const ref_tracker* t);
void ref_tracker_destroy(ref_tracker* t);

/* ---- row P + nalgebra pieces ---------------------------------------------------------- */
void ref_so3_hat(const float w[3], float out9[9]);   /* row-major 3x3 for readability */
void ref_so3_hat2(const float w[3], float out9[9]);
void ref_so3_vee(const float m9[9], float w[3]);
void ref_so3_exp(const float w[3], float q[4]);
void ref_so3_log(const float q[4], float w[3]);
void ref_se3_hat(const float xi[6], float out16[16]);
void ref_se3_vee(const float m16[16], float xi[6]);
void ref_se3_exp(const float xi[6], ref_pose* out);
void ref_se3_log(const ref_pose* p, float xi[6]);
void ref_pose_mul(const ref_pose* a, const ref_pose* b, ref_pose* out);
void ref_pose_inverse(const ref_pose* a, ref_pose* out);
void ref_pose_transform(const ref_pose* a, const float p[3], float out[3]);
void ref_quat_from_euler(float roll, float pitch, float yaw, float q[4]);
int ref_cholesky_solve6(const float H[36], const float g[6], float x[6]);
void ref_warp(const ref_pose* model, float x, float y, float idepth, const float intr5[5], float uv[2]);
void ref_warp_jacobian_at(float gu, float gv, float u, float v, float idepth, const float intr5[5],
                          float out6[6]);
int ref_interpolate(float x, float y, const uint8_t* image, int rows, int cols, float* out);

#ifdef __cplusplus
}
#endif
#endif
