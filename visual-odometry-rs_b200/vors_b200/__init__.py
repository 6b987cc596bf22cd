"""vors_b200 — Python host layer over libvors_b200.so (ctypes; no torch types cross the boundary).

Mirrors the reference's public interface for the direct-alignment path
(src/core/track/inverse_compositional.rs): `Config` -> `Config.init(...)` -> `Tracker.track(...)` /
`Tracker.current_frame()`, plus `BatchTracker` (n independent streams per device launch) and the
inner seams (`mean_pyramid`, `gradients`, `candidates_coarse_to_fine`, `Keyframe`).

There is no CPU fallback: importing works anywhere, but every compute call raises `VorsError`
unless the CUDA library is built (visual-odometry-rs_b200/lib/libvors_b200.so) and an sm_100 GPU
is visible.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libvors_b200.so")
MAX_LEVELS = 8
COL_MAJOR, ROW_MAJOR = 0, 1
CANDIDATES_COARSE_TO_FINE, CANDIDATES_DENSE, CANDIDATES_DSO = 0, 1, 2
OK, OPTIMIZATION_FAILED, E_INVALID, E_CUDA, E_NOMEM = 0, 1, -1, -2, -3


class VorsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"vors_b200 error {code}: {msg}")
        self.code = code


class ConfigStruct(C.Structure):
    """vors_config (include/vors_b200.h)."""

    _fields_ = [
        ("nb_levels", C.c_uint32),
        ("candidates_diff_threshold", C.c_uint32),
        ("depth_scale", C.c_float),
        ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("skew", C.c_float),
        ("idepth_variance", C.c_float),
        ("candidate_mode", C.c_uint32),
        ("fixed_iters", C.c_uint32),
        ("lm_coef_init", C.c_float),
        ("lm_coef_reject_mult", C.c_float),
        ("lm_coef_accept_mult", C.c_float),
        ("energy_delta_stop", C.c_float),
        ("max_iters", C.c_uint32),
        ("keyframe_flow_threshold", C.c_float),
        ("device", C.c_int32),
        ("team_size", C.c_uint32),
        ("dso_nb_target", C.c_uint32),
        ("idepth_fusion", C.c_uint32),
        ("huber_delta", C.c_float),
        ("gradient_operator", C.c_uint32),
    ]


class Pose(C.Structure):
    """vors_pose: translation + unit quaternion (x, y, z, w)."""

    _fields_ = [("t", C.c_float * 3), ("q", C.c_float * 4)]

    @staticmethod
    def identity():
        return Pose((C.c_float * 3)(0, 0, 0), (C.c_float * 4)(0, 0, 0, 1))

    @staticmethod
    def from_arrays(t, q):
        return Pose((C.c_float * 3)(*[float(v) for v in t]), (C.c_float * 4)(*[float(v) for v in q]))

    def as_array(self):
        return np.array(list(self.t) + list(self.q), dtype=np.float32)


class TraceRec(C.Structure):
    _fields_ = [("level", C.c_int32), ("iter", C.c_int32), ("energy", C.c_float), ("n_inside", C.c_int32),
                ("lm_coef", C.c_float), ("accepted", C.c_int32)]


class TrackStats(C.Structure):
    _fields_ = [("status", C.c_int32), ("keyframe_changed", C.c_int32), ("optical_flow", C.c_float),
                ("n_iters", C.c_int32 * MAX_LEVELS), ("energy", C.c_float * MAX_LEVELS),
                ("n_points", C.c_int32 * MAX_LEVELS), ("n_passes", C.c_int32), ("reserved", C.c_int32)]


# Every symbol include/vors_b200.h declares: name -> (restype, argtypes)
_vp = C.c_void_p
_P = C.POINTER
SIGNATURES = {
    "vors_config_default": (None, [_P(ConfigStruct)]),
    "vors_last_error": (C.c_char_p, []),
    "vors_version": (C.c_char_p, []),
    "vors_build_info": (C.c_char_p, []),
    "vors_device_count": (C.c_int, []),
    "vors_tracker_create": (C.c_int, [_P(ConfigStruct), C.c_double, _vp, C.c_double, _vp, C.c_uint32, C.c_uint32, C.c_int, _P(_vp)]),
    "vors_tracker_track": (C.c_int, [_vp, C.c_double, _vp, C.c_double, _vp, _P(TrackStats)]),
    "vors_tracker_current_frame": (C.c_int, [_vp, _P(C.c_double), _P(Pose)]),
    "vors_tracker_keyframe_pose": (C.c_int, [_vp, _P(Pose)]),
    "vors_tracker_set_tracing": (C.c_int, [_vp, C.c_int]),
    "vors_tracker_last_trace": (C.c_int, [_vp, _P(TraceRec), C.c_int, _P(C.c_int)]),
    "vors_tracker_destroy": (None, [_vp]),
    "vors_batch_create": (C.c_int, [_P(ConfigStruct), C.c_uint32, _vp, _vp, _vp, _vp, C.c_uint32, C.c_uint32, C.c_int, _P(_vp)]),
    "vors_batch_track": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "vors_batch_track_next": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "vors_batch_track_device": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "vors_batch_track_device_next": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "vors_batch_cancel_prefetch": (C.c_int, [_vp]),
    "vors_batch_current_frames": (C.c_int, [_vp, _vp, _vp]),
    "vors_batch_size": (C.c_int, [_vp]),
    "vors_batch_last_timing": (C.c_int, [_vp, _P(C.c_float)]),
    "vors_batch_last_counters": (C.c_int, [_vp, _P(C.c_uint64), _P(C.c_uint64)]),
    "vors_batch_last_launch_shape": (C.c_int, [_vp, _P(C.c_int), _P(C.c_int)]),
    "vors_batch_set_tracing": (C.c_int, [_vp, C.c_int]),
    "vors_batch_last_trace": (C.c_int, [_vp, C.c_uint32, _P(TraceRec), C.c_int, _P(C.c_int)]),
    "vors_batch_destroy": (None, [_vp]),
    "vors_pyramid_shapes": (C.c_int, [C.c_uint32, C.c_uint32, C.c_uint32, _vp, _vp]),
    "vors_mean_pyramid": (C.c_int, [_vp, C.c_uint32, C.c_uint32, C.c_uint32, _vp]),
    "vors_gradients": (C.c_int, [_vp, C.c_uint32, C.c_uint32, C.c_uint32, _vp, _vp, _vp]),
    "vors_candidates_coarse_to_fine": (C.c_int, [C.c_uint16, _vp, C.c_uint32, C.c_uint32, C.c_uint32, _vp]),
    "vors_candidates_dso": (C.c_int, [_vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, _vp, _P(C.c_int)]),
    "vors_gradient_norms_example": (C.c_int, [_vp, C.c_uint32, C.c_uint32, C.c_uint32, _vp]),
    "vors_keyframe_create": (C.c_int, [_P(ConfigStruct), _vp, _vp, C.c_uint32, C.c_uint32, C.c_int, _P(_vp)]),
    "vors_keyframe_levels": (C.c_int, [_vp]),
    "vors_keyframe_n_points": (C.c_int, [_vp, C.c_uint32]),
    "vors_keyframe_points": (C.c_int, [_vp, C.c_uint32, _vp, _vp, _vp, _vp]),
    "vors_keyframe_jacobians": (C.c_int, [_vp, C.c_uint32, _vp]),
    "vors_keyframe_mask0": (C.c_int, [_vp, _vp]),
    "vors_keyframe_idepth_map": (C.c_int, [_vp, C.c_uint32, _vp]),
    "vors_keyframe_destroy": (None, [_vp]),
    "vors_align_pass": (C.c_int, [_vp, C.c_uint32, _vp, _P(Pose), _P(C.c_float), _P(C.c_int32), _vp, _vp]),
    "vors_align_level": (C.c_int, [_vp, C.c_uint32, _vp, _P(Pose), _P(Pose), _P(C.c_int32), _P(C.c_float), _P(TraceRec), C.c_int, _P(C.c_int)]),
    "vors_align": (C.c_int, [_vp, _vp, C.c_int, _P(Pose), _P(Pose), _P(TrackStats), _P(TraceRec), C.c_int, _P(C.c_int)]),
    "vors_se3_exp": (C.c_int, [_vp, _P(Pose)]),
    "vors_se3_log": (C.c_int, [_P(Pose), _vp]),
    "vors_so3_exp": (C.c_int, [_vp, _vp]),
    "vors_so3_log": (C.c_int, [_vp, _vp]),
}

_lib = None


def load_library() -> C.CDLL:
    """Load libvors_b200.so; fails loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VorsError(E_CUDA, f"{LIB_PATH} is missing: build it with `make -C visual-odometry-rs_b200` "
                                    "(or __graft_entry__.build()); there is no CPU fallback")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def source_hash() -> str:
    """sha256/16 over the CUDA sources and headers, computed the way the Makefile does (SRC then HDR order)."""
    import hashlib

    pkg = os.path.dirname(_HERE)
    files = ["csrc/image_kernels.cu", "csrc/dso_kernels.cu", "csrc/align_kernel.cu", "csrc/engine.cu",
             "csrc/vors_device.cuh", "csrc/lie.cuh", "../include/vors_b200.h"]
    h = hashlib.sha256()
    for f in files:
        with open(os.path.join(pkg, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def build_info() -> dict:
    """What the loaded libvors_b200.so says it was built from, and whether that matches the sources in the tree."""
    info = dict(kv.split("=", 1) for kv in load_library().vors_build_info().decode().split())
    info["tree_src"] = source_hash()
    info["matches_tree"] = info.get("src") == info["tree_src"]
    return info


def _check(rc: int, allow_failed: bool = False) -> int:
    if rc < 0 or (rc == OPTIMIZATION_FAILED and not allow_failed):
        raise VorsError(rc, load_library().vors_last_error().decode())
    return rc


def _count(rc: int) -> int:
    """For entry points that return a count (levels, points): only negative values are errors."""
    if rc < 0:
        raise VorsError(rc, load_library().vors_last_error().decode())
    return rc


def device_count() -> int:
    return load_library().vors_device_count()


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _cm(a: np.ndarray, dtype) -> np.ndarray:
    """[row, col] array -> flat column-major buffer (what nalgebra's as_slice() yields)."""
    return np.ascontiguousarray(np.asarray(a, dtype).T).reshape(-1)


def _from_cm(flat, rows, cols):
    return flat.reshape(cols, rows).T


class Config:
    """Mirror of `track::Config` (inverse_compositional.rs:37-49) with the reference defaults of
    src/bin/vors_track.rs:34-40; extension fields default to the reference's hard-coded constants."""

    def __init__(self, **kw):
        self.c = ConfigStruct()
        load_library().vors_config_default(C.byref(self.c))
        for k, v in kw.items():
            if not hasattr(self.c, k):
                raise AttributeError(k)
            setattr(self.c, k, v)

    def __getattr__(self, k):
        return getattr(self.__dict__["c"], k)

    def init(self, keyframe_depth_timestamp, depth_map, keyframe_img_timestamp, img) -> "Tracker":
        """`Config::init` (inverse_compositional.rs:74-100): images are [row, col] numpy arrays."""
        return Tracker(self, keyframe_depth_timestamp, depth_map, keyframe_img_timestamp, img)


class Tracker:
    """Mirror of `Tracker` (inverse_compositional.rs:30-34, 170-248)."""

    def __init__(self, cfg: Config, depth_ts, depth, img_ts, img):
        self._lib = load_library()
        img = np.ascontiguousarray(img, np.uint8)
        depth = np.ascontiguousarray(depth, np.uint16)
        if img.shape != depth.shape or img.ndim != 2:
            raise VorsError(E_INVALID, "image and depth map must be 2-D arrays of the same shape")
        self.rows, self.cols = img.shape
        h = C.c_void_p()
        # numpy [row, col] buffers are row-major: same situation as src/bin/vors_track.rs:142
        _check(self._lib.vors_tracker_create(C.byref(cfg.c), depth_ts, _ptr(depth), img_ts, _ptr(img), self.rows, self.cols,
                                             ROW_MAJOR, C.byref(h)))
        self._h = h

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.vors_tracker_destroy(self._h)
            self._h = None

    def set_tracing(self, enabled=True):
        _check(self._lib.vors_tracker_set_tracing(self._h, int(enabled)))

    def track(self, depth_time, depth_map, img_time, img) -> TrackStats:
        """`Tracker::track`.  An optimisation failure is not an exception (the reference only logs it):
        check `stats.status`."""
        img = np.ascontiguousarray(img, np.uint8)
        depth = np.ascontiguousarray(depth_map, np.uint16)
        if img.shape != (self.rows, self.cols) or depth.shape != (self.rows, self.cols):
            raise VorsError(E_INVALID, "frame shape differs from the tracker's")
        stats = TrackStats()
        _check(self._lib.vors_tracker_track(self._h, depth_time, _ptr(depth), img_time, _ptr(img), C.byref(stats)), True)
        return stats

    def current_frame(self):
        """`Tracker::current_frame` -> (depth timestamp, Pose)."""
        ts = C.c_double()
        p = Pose()
        _check(self._lib.vors_tracker_current_frame(self._h, C.byref(ts), C.byref(p)))
        return ts.value, p

    def keyframe_pose(self) -> Pose:
        p = Pose()
        _check(self._lib.vors_tracker_keyframe_pose(self._h, C.byref(p)))
        return p

    def last_trace(self, cap=256):
        tr = (TraceRec * cap)()
        n = C.c_int()
        _check(self._lib.vors_tracker_last_trace(self._h, tr, cap, C.byref(n)))
        return [tr[i] for i in range(n.value)]


class BatchTracker:
    """n independent streams advanced together: one persistent device launch per frame set."""

    def __init__(self, cfg: Config, depth_ts, depths, img_ts, imgs, layout=ROW_MAJOR):
        """depths / imgs: arrays [n, rows, cols] (u16 / u8) in `layout` order per frame."""
        self._lib = load_library()
        imgs = np.ascontiguousarray(imgs, np.uint8)
        depths = np.ascontiguousarray(depths, np.uint16)
        self.n, self.rows, self.cols = imgs.shape
        self.layout = layout
        h = C.c_void_p()
        dts = np.ascontiguousarray(depth_ts, np.float64)
        its = np.ascontiguousarray(img_ts, np.float64)
        ip, dp = self._ptr_arrays(imgs, depths)
        _check(self._lib.vors_batch_create(C.byref(cfg.c), self.n, _ptr(dts), dp, _ptr(its), ip, self.rows, self.cols, layout,
                                           C.byref(h)))
        self._h = h

    def _ptr_arrays(self, imgs, depths):
        I = self.rows * self.cols
        ip = (C.c_void_p * self.n)(*[imgs.ctypes.data + i * I for i in range(self.n)])
        dp = (C.c_void_p * self.n)(*[depths.ctypes.data + i * I * 2 for i in range(self.n)]) if depths is not None else None
        return ip, dp

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.vors_batch_destroy(self._h)
            self._h = None

    def set_tracing(self, enabled=True):
        _check(self._lib.vors_batch_set_tracing(self._h, int(enabled)))

    def track(self, depth_ts, depths, img_ts, imgs, next_imgs=None):
        """Host buffers [n, rows, cols]; returns (status[n], stats list).  `next_imgs`: the C-contiguous uint8 array the NEXT
        call will be given as `imgs` (same object): its upload then overlaps this call's alignment."""
        imgs = np.ascontiguousarray(imgs, np.uint8)
        depths = np.ascontiguousarray(depths, np.uint16)
        dts = np.ascontiguousarray(depth_ts, np.float64)
        its = np.ascontiguousarray(img_ts, np.float64)
        ip, dp = self._ptr_arrays(imgs, depths)
        status = np.zeros(self.n, np.int32)
        stats = (TrackStats * self.n)()
        if next_imgs is not None:
            if not (isinstance(next_imgs, np.ndarray) and next_imgs.dtype == np.uint8 and next_imgs.flags.c_contiguous):
                raise ValueError("next_imgs must be a C-contiguous uint8 array (it is recognised by address in the next call)")
            self._announced = next_imgs  # keep the announced buffer alive until the next call
            nip, _ = self._ptr_arrays(next_imgs, depths)
            _check(self._lib.vors_batch_track_next(self._h, _ptr(dts), dp, _ptr(its), ip, nip, _ptr(status), stats), True)
        else:
            _check(self._lib.vors_batch_track(self._h, _ptr(dts), dp, _ptr(its), ip, _ptr(status), stats), True)
        return status, list(stats)

    def track_raw(self, dts_ptr, depth_ptrs, its_ptr, img_ptrs, status_ptr=None, stats_ptr=None, next_img_ptrs=None):
        """Pre-marshalled pointers (bench hot loop): no numpy work inside the timed region.  `next_img_ptrs` announces the
        frames of the next call (their upload then overlaps this call's alignment)."""
        if next_img_ptrs is not None:
            return _check(self._lib.vors_batch_track_next(self._h, dts_ptr, depth_ptrs, its_ptr, img_ptrs, next_img_ptrs, status_ptr,
                                                          stats_ptr), True)
        return _check(self._lib.vors_batch_track(self._h, dts_ptr, depth_ptrs, its_ptr, img_ptrs, status_ptr, stats_ptr), True)

    def track_device(self, dts_ptr, depth_dev_ptr, its_ptr, img_dev_ptr, status_ptr=None, stats_ptr=None, next_img_dev_ptr=None):
        """Device-resident column-major inputs ([n, cols, rows] memory order).  `next_img_dev_ptr` announces the buffer of the
        next call."""
        if next_img_dev_ptr is not None:
            return _check(self._lib.vors_batch_track_device_next(self._h, dts_ptr, depth_dev_ptr, its_ptr, img_dev_ptr, next_img_dev_ptr,
                                                                 status_ptr, stats_ptr), True)
        return _check(self._lib.vors_batch_track_device(self._h, dts_ptr, depth_dev_ptr, its_ptr, img_dev_ptr, status_ptr,
                                                        stats_ptr), True)

    def cancel_prefetch(self):
        """Forget the announced next frames (and wait for their copy): the buffers may then be reused."""
        _check(self._lib.vors_batch_cancel_prefetch(self._h))
        self._announced = None

    def current_frames(self):
        ts = np.zeros(self.n, np.float64)
        poses = np.zeros((self.n, 7), np.float32)  # vors_pose is 7 packed floats
        _check(self._lib.vors_batch_current_frames(self._h, _ptr(ts), _ptr(poses)))
        return ts, poses

    def last_timing(self):
        ms = (C.c_float * 4)()
        _check(self._lib.vors_batch_last_timing(self._h, ms))
        return dict(upload_ms=ms[0], pyramid_ms=ms[1], align_ms=ms[2], keyframe_ms=ms[3])

    def last_counters(self):
        a, b = C.c_uint64(), C.c_uint64()
        _check(self._lib.vors_batch_last_counters(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def last_launch_shape(self):
        """(CTAs per alignment, alignments in flight) of the last align launch."""
        a, b = C.c_int(), C.c_int()
        _check(self._lib.vors_batch_last_launch_shape(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def last_trace(self, stream, cap=256):
        tr = (TraceRec * cap)()
        n = C.c_int()
        _check(self._lib.vors_batch_last_trace(self._h, stream, tr, cap, C.byref(n)))
        return [tr[i] for i in range(n.value)]


# ---- inner seams ---------------------------------------------------------------------------------

def pyramid_shapes(rows, cols, max_levels):
    r = np.zeros(32, np.uint32)
    c = np.zeros(32, np.uint32)
    n = _count(load_library().vors_pyramid_shapes(rows, cols, max_levels, _ptr(r), _ptr(c)))
    return [(int(r[i]), int(c[i])) for i in range(n)]


def _split(flat, shapes):
    out, off = [], 0
    for r, c in shapes:
        out.append(_from_cm(flat[off:off + r * c], r, c))
        off += r * c
    return out


def mean_pyramid(img, max_levels):
    """`multires::mean_pyramid` (multires.rs:21-31) on the GPU; img is [row, col] u8."""
    rows, cols = img.shape
    shapes = pyramid_shapes(rows, cols, max_levels)
    out = np.zeros(sum(r * c for r, c in shapes), np.uint8)
    n = _count(load_library().vors_mean_pyramid(_ptr(_cm(img, np.uint8)), rows, cols, max_levels, _ptr(out)))
    assert n == len(shapes)
    return _split(out, shapes)


def gradients(img, max_levels):
    """Tracker gradient recipe (inverse_compositional.rs:112-117) -> (gx levels, gy levels, g2 levels)."""
    rows, cols = img.shape
    shapes = pyramid_shapes(rows, cols, max_levels)
    total = sum(r * c for r, c in shapes)
    gx, gy, g2 = np.zeros(total, np.int16), np.zeros(total, np.int16), np.zeros(total, np.uint16)
    _count(load_library().vors_gradients(_ptr(_cm(img, np.uint8)), rows, cols, max_levels, _ptr(gx), _ptr(gy), _ptr(g2)))
    return _split(gx, shapes), _split(gy, shapes), _split(g2, shapes)


def candidates_coarse_to_fine(diff_threshold, g2_levels):
    """`candidates::coarse_to_fine::select` (coarse_to_fine.rs:15-32); finest first in and out."""
    rows, cols = g2_levels[0].shape
    shapes = [g.shape for g in g2_levels]
    cat = np.concatenate([_cm(g, np.uint16) for g in g2_levels])
    out = np.zeros(cat.size, np.uint8)
    _check(load_library().vors_candidates_coarse_to_fine(diff_threshold, _ptr(cat), rows, cols, len(g2_levels), _ptr(out)))
    return [m.astype(bool) for m in _split(out, shapes)]


def candidates_dso(gradients, nb_target, nb_iterations_left=1, seed=0):
    """`candidates::dso::select` (dso.rs:98-150) -> (mask [row, col] bool, nb block candidates, used_random_branch)."""
    rows, cols = gradients.shape
    out = np.zeros(rows * cols, np.uint8)
    used = C.c_int()
    nb = _count(load_library().vors_candidates_dso(_ptr(_cm(gradients, np.uint16)), rows, cols, nb_target, nb_iterations_left, seed,
                                                   _ptr(out), C.byref(used)))
    return _from_cm(out, rows, cols).astype(bool), nb, bool(used.value)


def gradient_norms_example(img, max_levels):
    """Example recipe: squared_norm_direct at level 0, bloc_squared_norm above; finest first."""
    rows, cols = img.shape
    shapes = pyramid_shapes(rows, cols, max_levels)
    out = np.zeros(sum(r * c for r, c in shapes), np.uint16)
    _count(load_library().vors_gradient_norms_example(_ptr(_cm(img, np.uint8)), rows, cols, max_levels, _ptr(out)))
    return _split(out, shapes)


def se3_exp(xi) -> Pose:
    p = Pose()
    x = np.ascontiguousarray(xi, np.float32)
    _check(load_library().vors_se3_exp(_ptr(x), C.byref(p)))
    return p


def se3_log(pose: Pose) -> np.ndarray:
    xi = np.zeros(6, np.float32)
    _check(load_library().vors_se3_log(C.byref(pose), _ptr(xi)))
    return xi


def so3_exp(w) -> np.ndarray:
    q = np.zeros(4, np.float32)
    _check(load_library().vors_so3_exp(_ptr(np.ascontiguousarray(w, np.float32)), _ptr(q)))
    return q


def so3_log(q) -> np.ndarray:
    w = np.zeros(3, np.float32)
    _check(load_library().vors_so3_log(_ptr(np.ascontiguousarray(q, np.float32)), _ptr(w)))
    return w


class Keyframe:
    """`precompute_multires_data` (inverse_compositional.rs:105-161) result, resident on the device."""

    def __init__(self, cfg: Config, depth, img):
        self._lib = load_library()
        img = np.ascontiguousarray(img, np.uint8)
        depth = np.ascontiguousarray(depth, np.uint16)
        self.rows, self.cols = img.shape
        self.cfg = cfg
        h = C.c_void_p()
        _check(self._lib.vors_keyframe_create(C.byref(cfg.c), _ptr(depth), _ptr(img), self.rows, self.cols, ROW_MAJOR, C.byref(h)))
        self._h = h
        self.shapes = pyramid_shapes(self.rows, self.cols, cfg.nb_levels)

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.vors_keyframe_destroy(self._h)
            self._h = None

    @property
    def levels(self):
        return _count(self._lib.vors_keyframe_levels(self._h))

    def n_points(self, lvl):
        return _count(self._lib.vors_keyframe_n_points(self._h, lvl))

    def points(self, lvl):
        n = self.n_points(lvl)
        xy = np.zeros((n, 2), np.uint32)
        idepth = np.zeros(n, np.float32)
        grad = np.zeros((n, 2), np.int16)
        tmpl = np.zeros(n, np.uint8)
        _check(self._lib.vors_keyframe_points(self._h, lvl, _ptr(xy), _ptr(idepth), _ptr(grad), _ptr(tmpl)))
        return xy, idepth, grad, tmpl

    def jacobians(self, lvl):
        jac = np.zeros((self.n_points(lvl), 6), np.float32)
        _check(self._lib.vors_keyframe_jacobians(self._h, lvl, _ptr(jac)))
        return jac

    def mask0(self):
        out = np.zeros(self.rows * self.cols, np.uint8)
        _check(self._lib.vors_keyframe_mask0(self._h, _ptr(out)))
        return _from_cm(out, self.rows, self.cols).astype(bool)

    def idepth_map(self, lvl):
        r, c = self.shapes[lvl]
        out = np.zeros(r * c, np.float32)
        _check(self._lib.vors_keyframe_idepth_map(self._h, lvl, _ptr(out)))
        return _from_cm(out, r, c)

    def align_pass(self, lvl, image, model: Pose):
        """One `eval_energy` + `compute_eval_data` (lm_optimizer.rs:68-107): (energy, n_inside, g[6], H[6,6])."""
        e = C.c_float()
        n = C.c_int32()
        g = np.zeros(6, np.float32)
        H = np.zeros(36, np.float32)
        _check(self._lib.vors_align_pass(self._h, lvl, _ptr(_cm(image, np.uint8)), C.byref(model), C.byref(e), C.byref(n), _ptr(g), _ptr(H)))
        return e.value, n.value, g, H.reshape(6, 6)

    def align_level(self, lvl, image, init: Pose, trace_cap=256):
        """`LMOptimizerState::iterative_solve` on one level -> (status, pose, n_iter, energy, trace)."""
        out = Pose()
        it = C.c_int32()
        en = C.c_float()
        tr = (TraceRec * trace_cap)()
        tl = C.c_int()
        st = _check(self._lib.vors_align_level(self._h, lvl, _ptr(_cm(image, np.uint8)), C.byref(init), C.byref(out), C.byref(it),
                                               C.byref(en), tr, trace_cap, C.byref(tl)), True)
        return st, out, it.value, en.value, [tr[i] for i in range(tl.value)]

    def align(self, img, init: Pose, trace_cap=256):
        """Level loop of `Tracker::track` for one full-resolution frame -> (status, pose, stats, trace)."""
        img = np.ascontiguousarray(img, np.uint8)
        out = Pose()
        stats = TrackStats()
        tr = (TraceRec * trace_cap)()
        tl = C.c_int()
        st = _check(self._lib.vors_align(self._h, _ptr(img), ROW_MAJOR, C.byref(init), C.byref(out), C.byref(stats), tr, trace_cap,
                                         C.byref(tl)), True)
        return st, out, stats, [tr[i] for i in range(tl.value)]
