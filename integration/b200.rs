// integration/b200.rs - the Rust side of the drop-in boundary: the file a vors maintainer adds as
// `src/core/track/b200.rs` so that `src/bin/vors_track.rs` keeps driving the tracker unchanged through
// libvors_b200.so (include/vors_b200.h).  Each item names the reference interface it stands in for
// (path:line under the reference checkout).
//
// STATUS: UNBUILT AND UNTESTED - the build image has no cargo / rustc.  What IS checked here, without a Rust toolchain
// (tests/test_rust_shim.py): every `extern "C"` function declared below exists in include/vors_b200.h with the same
// number of parameters, and `VorsConfig` lists the fields of `vors_config` in the same order with matching widths.
// The same C ABI is exercised end to end from Python (ctypes) and from C++ (tools/vors_track.cpp).
//
// Wiring (see also integration/build.rs):
//   src/core/track/mod.rs:   pub mod b200;
//   src/bin/vors_track.rs:13 use vors::core::track::b200 as track;      // lines 34-63 compile unchanged
use nalgebra::{DMatrix, Quaternion, Translation3, UnitQuaternion};
use std::os::raw::{c_char, c_int};
use crate::core::camera::Intrinsics;
use crate::misc::type_aliases::{Float, Iso3};

#[repr(C)]
#[derive(Clone, Copy)]
pub struct VorsConfig {
    pub nb_levels: u32,
    pub candidates_diff_threshold: u32,
    pub depth_scale: f32,
    pub fx: f32, pub fy: f32, pub cx: f32, pub cy: f32, pub skew: f32,
    pub idepth_variance: f32,
    pub candidate_mode: u32,
    pub fixed_iters: u32,
    pub lm_coef_init: f32,
    pub lm_coef_reject_mult: f32,
    pub lm_coef_accept_mult: f32,
    pub energy_delta_stop: f32,
    pub max_iters: u32,
    pub keyframe_flow_threshold: f32,
    pub device: i32,
    pub team_size: u32,
    pub dso_nb_target: u32,
    pub idepth_fusion: u32, // 0 = strategy_dso_mean (what Tracker uses), 1 = strategy_statistically_similar
    pub huber_delta: f32,   // 0 = the reference's plain L2
    pub gradient_operator: u32, // 0 = the Tracker's gradient recipe, 1 = Scharr (extension)
}
#[repr(C)] #[derive(Clone, Copy, Default)] pub struct VorsPose { pub t: [f32; 3], pub q: [f32; 4] }
#[repr(C)] pub struct VorsTracker { _private: [u8; 0] }

#[link(name = "vors_b200")]
extern "C" {
    fn vors_config_default(cfg: *mut VorsConfig);
    fn vors_last_error() -> *const c_char;
    fn vors_tracker_create(cfg: *const VorsConfig, depth_ts: f64, depth: *const u16, img_ts: f64, img: *const u8,
                           rows: u32, cols: u32, layout: c_int, out: *mut *mut VorsTracker) -> c_int;
    fn vors_tracker_track(t: *mut VorsTracker, depth_ts: f64, depth: *const u16, img_ts: f64, img: *const u8,
                          stats: *mut std::ffi::c_void) -> c_int;
    fn vors_tracker_current_frame(t: *const VorsTracker, depth_ts: *mut f64, pose: *mut VorsPose) -> c_int;
    fn vors_tracker_destroy(t: *mut VorsTracker);
}
const VORS_COL_MAJOR: c_int = 0;

/// Same fields as the reference's `Config` (inverse_compositional.rs:37-49).
pub struct Config {
    pub nb_levels: usize,
    pub candidates_diff_threshold: u16,
    pub depth_scale: Float,
    pub intrinsics: Intrinsics,
    pub idepth_variance: Float,
}
pub struct Tracker { raw: *mut VorsTracker }

impl Config {
    /// `Config::init` (inverse_compositional.rs:74-100): same signature.
    pub fn init(self, depth_ts: f64, depth_map: &DMatrix<u16>, img_ts: f64, img: DMatrix<u8>) -> Tracker {
        let mut c: VorsConfig = unsafe { std::mem::zeroed() };
        unsafe { vors_config_default(&mut c) };
        c.nb_levels = self.nb_levels as u32;
        c.candidates_diff_threshold = self.candidates_diff_threshold as u32;
        c.depth_scale = self.depth_scale;
        c.fx = self.intrinsics.focal.0; c.fy = self.intrinsics.focal.1;
        c.cx = self.intrinsics.principal_point.0; c.cy = self.intrinsics.principal_point.1;
        c.skew = self.intrinsics.skew;
        c.idepth_variance = self.idepth_variance;
        let (rows, cols) = img.shape();
        let mut raw = std::ptr::null_mut();
        let rc = unsafe {
            vors_tracker_create(&c, depth_ts, depth_map.as_slice().as_ptr(), img_ts, img.as_slice().as_ptr(),
                                rows as u32, cols as u32, VORS_COL_MAJOR, &mut raw)
        };
        // the reference panics on degenerate sizes (index out of bounds); keep that contract
        assert!(rc == 0, "vors_tracker_create failed: {}", unsafe { std::ffi::CStr::from_ptr(vors_last_error()).to_string_lossy() });
        Tracker { raw }
    }
}

impl Tracker {
    /// `Tracker::track` (inverse_compositional.rs:170-240): same signature; an optimisation failure is logged, not returned.
    pub fn track(&mut self, depth_time: f64, depth_map: &DMatrix<u16>, img_time: f64, img: DMatrix<u8>) {
        let rc = unsafe {
            vors_tracker_track(self.raw, depth_time, depth_map.as_slice().as_ptr(), img_time, img.as_slice().as_ptr(),
                               std::ptr::null_mut())
        };
        if rc == 1 { eprintln!("Error at Cholesky decomposition of hessian"); }   // lm_optimizer.rs:133
        assert!(rc >= 0, "vors_tracker_track failed");
    }
    /// `Tracker::current_frame` (inverse_compositional.rs:243-248).
    pub fn current_frame(&self) -> (f64, Iso3) {
        let (mut ts, mut p) = (0.0f64, VorsPose::default());
        unsafe { vors_tracker_current_frame(self.raw, &mut ts, &mut p) };
        let q = UnitQuaternion::new_unchecked(Quaternion::new(p.q[3], p.q[0], p.q[1], p.q[2]));
        (ts, Iso3::from_parts(Translation3::new(p.t[0], p.t[1], p.t[2]), q))
    }
}
impl Drop for Tracker { fn drop(&mut self) { unsafe { vors_tracker_destroy(self.raw) } } }

// ---- inner seams: for a maintainer who keeps the reference's own `Tracker` (inverse_compositional.rs) and only swaps
// ---- the hot functions it calls.  Buffers are nalgebra's column-major slices: no copy, no transpose.

#[repr(C)] pub struct VorsKeyframe { _private: [u8; 0] }
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct VorsTraceRec { pub level: i32, pub iter: i32, pub energy: f32, pub n_inside: i32, pub lm_coef: f32, pub accepted: i32 }

#[link(name = "vors_b200")]
extern "C" {
    fn vors_pyramid_shapes(rows: u32, cols: u32, max_levels: u32, out_rows: *mut u32, out_cols: *mut u32) -> c_int;
    fn vors_mean_pyramid(img: *const u8, rows: u32, cols: u32, max_levels: u32, out_concat: *mut u8) -> c_int;
    fn vors_keyframe_create(cfg: *const VorsConfig, depth: *const u16, img: *const u8, rows: u32, cols: u32, layout: c_int,
                            out: *mut *mut VorsKeyframe) -> c_int;
    fn vors_keyframe_destroy(kf: *mut VorsKeyframe);
    fn vors_align_level(kf: *const VorsKeyframe, level: u32, image: *const u8, init: *const VorsPose, out: *mut VorsPose,
                        n_iter: *mut i32, energy: *mut f32, trace: *mut VorsTraceRec, trace_cap: c_int, trace_len: *mut c_int) -> c_int;
}

fn to_pose(m: &Iso3) -> VorsPose {
    let q = m.rotation.as_ref().coords; // (x, y, z, w), the order tum_rgbd.rs:78-85 prints
    VorsPose { t: [m.translation.vector[0], m.translation.vector[1], m.translation.vector[2]], q: [q[0], q[1], q[2], q[3]] }
}
fn from_pose(p: &VorsPose) -> Iso3 {
    let q = UnitQuaternion::new_unchecked(Quaternion::new(p.q[3], p.q[0], p.q[1], p.q[2]));
    Iso3::from_parts(Translation3::new(p.t[0], p.t[1], p.t[2]), q)
}

/// `multires::mean_pyramid(max_levels, DMatrix<u8>) -> Vec<DMatrix<u8>>` (src/core/multires.rs:21-31).
pub fn mean_pyramid(max_levels: usize, img: DMatrix<u8>) -> Vec<DMatrix<u8>> {
    let (rows, cols) = img.shape();
    let (mut lr, mut lc) = ([0u32; 32], [0u32; 32]);
    let n = unsafe { vors_pyramid_shapes(rows as u32, cols as u32, max_levels as u32, lr.as_mut_ptr(), lc.as_mut_ptr()) } as usize;
    let total: usize = (0..n).map(|l| (lr[l] * lc[l]) as usize).sum();
    let mut concat = vec![0u8; total];
    let rc = unsafe { vors_mean_pyramid(img.as_slice().as_ptr(), rows as u32, cols as u32, max_levels as u32, concat.as_mut_ptr()) };
    assert!(rc >= 0, "vors_mean_pyramid failed");
    let mut out = Vec::with_capacity(n);
    let mut off = 0;
    for l in 0..n {
        let sz = (lr[l] * lc[l]) as usize;
        out.push(DMatrix::from_column_slice(lr[l] as usize, lc[l] as usize, &concat[off..off + sz]));
        off += sz;
    }
    out
}

/// Device-resident `MultiresData` (inverse_compositional.rs:64-70) built by `precompute_multires_data` (:105-161).
pub struct Keyframe { raw: *mut VorsKeyframe }
impl Keyframe {
    pub fn new(cfg: &VorsConfig, depth_map: &DMatrix<u16>, img: &DMatrix<u8>) -> Keyframe {
        let (rows, cols) = img.shape();
        let mut raw = std::ptr::null_mut();
        let rc = unsafe { vors_keyframe_create(cfg, depth_map.as_slice().as_ptr(), img.as_slice().as_ptr(), rows as u32, cols as u32,
                                               VORS_COL_MAJOR, &mut raw) };
        assert!(rc == 0, "vors_keyframe_create failed");
        Keyframe { raw }
    }
    /// `LMOptimizerState::iterative_solve(&Obs, Iso3) -> Result<(Self, usize), String>` (src/math/optimizer.rs:57-70 with
    /// lm_optimizer.rs:113-192) on one level: Ok((model, nb_iter)) or the reference's Cholesky error string.
    pub fn iterative_solve(&self, level: usize, image: &DMatrix<u8>, model: Iso3) -> Result<(Iso3, usize), String> {
        let (init, mut out, mut n_iter, mut energy) = (to_pose(&model), VorsPose::default(), 0i32, 0f32);
        let rc = unsafe { vors_align_level(self.raw, level as u32, image.as_slice().as_ptr(), &init, &mut out, &mut n_iter, &mut energy,
                                           std::ptr::null_mut(), 0, std::ptr::null_mut()) };
        match rc {
            0 => Ok((from_pose(&out), n_iter as usize)),
            1 => Err("Error at Cholesky decomposition of hessian".to_string()), // lm_optimizer.rs:133
            _ => panic!("vors_align_level failed"),
        }
    }
}
impl Drop for Keyframe { fn drop(&mut self) { unsafe { vors_keyframe_destroy(self.raw) } } }
