"""ctypes binding of the CPU oracle (oracle/_build/libvors_oracle*.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
PARITY UNPINNED except for the vectors listed in oracle/vors_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
MAX_LEVELS = 16


class Config(C.Structure):
    """Field-for-field the same layout as vors_config (include/vors_b200.h)."""

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
                ("n_points", C.c_int32 * MAX_LEVELS)]


def build(force: bool = False) -> None:
    """Compile the oracle with oracle/Makefile (g++ only)."""
    need = force or not all(
        os.path.exists(os.path.join(_HERE, "_build", n)) for n in ("libvors_oracle.so", "libvors_oracle_fast.so"))
    if not need:
        src_m = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("vors_oracle.cpp", "vors_oracle.h"))
        out_m = min(os.path.getmtime(os.path.join(_HERE, "_build", n))
                    for n in ("libvors_oracle.so", "libvors_oracle_fast.so"))
        need = src_m > out_m
    if need:
        subprocess.check_call(["make", "-C", _HERE, "-s", "all"])


_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_u16p = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")
_i16p = np.ctypeslib.ndpointer(np.int16, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def _load(name: str) -> C.CDLL:
    path = os.path.join(_HERE, "_build", name)
    if not os.path.exists(path):
        build()
    lib = C.CDLL(path)
    P = C.POINTER
    vp = C.c_void_p
    sig = {
        "ref_config_default": (None, [P(Config)]),
        "ref_set_accum_f64": (None, [C.c_int]),
        "ref_pyramid_shapes": (C.c_int, [C.c_int, C.c_int, C.c_int, _i32p, _i32p]),
        "ref_mean_pyramid": (C.c_int, [_u8p, C.c_int, C.c_int, C.c_int, _u8p]),
        "ref_gradient_centered": (None, [_u8p, C.c_int, C.c_int, _i16p, _i16p]),
        "ref_gradient_scharr": (None, [_u8p, C.c_int, C.c_int, _i16p, _i16p]),
        "ref_gradients_tracker": (None, [_u8p, C.c_int, C.c_int, C.c_int, _i16p, _i16p, _u16p]),
        "ref_squared_norm_direct": (None, [_u8p, C.c_int, C.c_int, _u16p]),
        "ref_gradients_squared_norm_example": (None, [_u8p, C.c_int, C.c_int, C.c_int, _u16p]),
        "ref_prune_with_thresh": (None, [C.c_uint16] * 5 + [_u8p]),
        "ref_c2f_select": (None, [C.c_uint16, _u16p, C.c_int, C.c_int, C.c_int, _u8p]),
        "ref_dso_select": (C.c_int, [_u16p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, _u8p, P(C.c_int)]),
        "ref_keyframe_create": (vp, [P(Config), _u16p, _u8p, C.c_int, C.c_int]),
        "ref_keyframe_destroy": (None, [vp]),
        "ref_keyframe_levels": (C.c_int, [vp]),
        "ref_keyframe_level_shape": (None, [vp, C.c_int, P(C.c_int), P(C.c_int)]),
        "ref_keyframe_intrinsics": (None, [vp, C.c_int, _f32p]),
        "ref_keyframe_n_points": (C.c_int, [vp, C.c_int]),
        "ref_keyframe_points": (None, [vp, C.c_int, vp, vp, vp]),
        "ref_keyframe_image": (None, [vp, C.c_int, _u8p]),
        "ref_keyframe_mask0": (None, [vp, _u8p]),
        "ref_keyframe_idepth_map": (None, [vp, C.c_int, _f32p, _f32p]),
        "ref_eval": (C.c_int, [vp, C.c_int, _u8p, C.c_int, C.c_int, P(Pose), C.c_int, P(C.c_float), _f32p, _f32p]),
        "ref_iterative_solve": (C.c_int, [P(Config), vp, C.c_int, _u8p, C.c_int, C.c_int, P(Pose), P(Pose),
                                          P(C.c_int), P(C.c_float), P(TraceRec), C.c_int, P(C.c_int)]),
        "ref_tracker_create": (vp, [P(Config), C.c_double, _u16p, C.c_double, _u8p, C.c_int, C.c_int, C.c_int]),
        "ref_tracker_track": (C.c_int, [vp, C.c_double, _u16p, C.c_double, _u8p, P(TrackStats), P(TraceRec),
                                        C.c_int, P(C.c_int)]),
        "ref_tracker_current_frame": (None, [vp, P(C.c_double), P(Pose)]),
        "ref_tracker_keyframe_pose": (None, [vp, P(Pose)]),
        "ref_tracker_keyframe": (vp, [vp]),
        "ref_tracker_destroy": (None, [vp]),
        "ref_so3_hat": (None, [_f32p, _f32p]),
        "ref_so3_hat2": (None, [_f32p, _f32p]),
        "ref_so3_vee": (None, [_f32p, _f32p]),
        "ref_so3_exp": (None, [_f32p, _f32p]),
        "ref_so3_log": (None, [_f32p, _f32p]),
        "ref_se3_hat": (None, [_f32p, _f32p]),
        "ref_se3_vee": (None, [_f32p, _f32p]),
        "ref_se3_exp": (None, [_f32p, P(Pose)]),
        "ref_se3_log": (None, [P(Pose), _f32p]),
        "ref_pose_mul": (None, [P(Pose), P(Pose), P(Pose)]),
        "ref_pose_inverse": (None, [P(Pose), P(Pose)]),
        "ref_pose_transform": (None, [P(Pose), _f32p, _f32p]),
        "ref_quat_from_euler": (None, [C.c_float, C.c_float, C.c_float, _f32p]),
        "ref_cholesky_solve6": (C.c_int, [_f32p, _f32p, _f32p]),
        "ref_warp": (None, [P(Pose), C.c_float, C.c_float, C.c_float, _f32p, _f32p]),
        "ref_warp_jacobian_at": (None, [C.c_float] * 5 + [_f32p, _f32p]),
        "ref_interpolate": (C.c_int, [C.c_float, C.c_float, _u8p, C.c_int, C.c_int, P(C.c_float)]),
    }
    for name_, (res, args) in sig.items():
        fn = getattr(lib, name_)
        fn.restype = res
        fn.argtypes = args
    return lib


_libs: dict[str, C.CDLL] = {}


def lib(fast: bool = False) -> C.CDLL:
    """Parity build by default; fast=True gives the -O3 timing build."""
    key = "libvors_oracle_fast.so" if fast else "libvors_oracle.so"
    if key not in _libs:
        _libs[key] = _load(key)
    return _libs[key]


def default_config(**kw) -> Config:
    cfg = Config()
    lib().ref_config_default(C.byref(cfg))
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


# ------------------------------------------------------------------------------------------
# numpy-friendly wrappers.  Images are passed as 2-D numpy arrays indexed [row, col]; the
# wrappers hand the oracle column-major bytes (np.asfortranarray) like nalgebra would.

def _cm(a: np.ndarray) -> np.ndarray:
    """Flat column-major copy of a [row, col] array."""
    return np.ascontiguousarray(np.asarray(a).T).reshape(-1)


def _from_cm(flat: np.ndarray, rows: int, cols: int) -> np.ndarray:
    return flat.reshape(cols, rows).T


def pyramid_shapes(rows: int, cols: int, max_levels: int):
    r = np.zeros(MAX_LEVELS * 2, np.int32)
    c = np.zeros(MAX_LEVELS * 2, np.int32)
    n = lib().ref_pyramid_shapes(rows, cols, max_levels, r, c)
    return [(int(r[i]), int(c[i])) for i in range(n)]


def split_concat(flat: np.ndarray, shapes):
    out, off = [], 0
    for (r, c) in shapes:
        out.append(_from_cm(flat[off:off + r * c], r, c))
        off += r * c
    return out


def mean_pyramid(img: np.ndarray, max_levels: int, fast: bool = False):
    rows, cols = img.shape
    shapes = pyramid_shapes(rows, cols, max_levels)
    out = np.zeros(sum(r * c for r, c in shapes), np.uint8)
    n = lib(fast).ref_mean_pyramid(_cm(img), rows, cols, max_levels, out)
    assert n == len(shapes)
    return split_concat(out, shapes)


def gradient_scharr(img):
    rows, cols = img.shape
    gx = np.zeros(rows * cols, np.int16)
    gy = np.zeros(rows * cols, np.int16)
    lib().ref_gradient_scharr(_cm(img), rows, cols, gx, gy)
    return _from_cm(gx, rows, cols), _from_cm(gy, rows, cols)


def gradients_tracker(pyr):
    rows, cols = pyr[0].shape
    shapes = [p.shape for p in pyr]
    total = sum(r * c for r, c in shapes)
    cat = np.concatenate([_cm(p) for p in pyr])
    gx = np.zeros(total, np.int16)
    gy = np.zeros(total, np.int16)
    g2 = np.zeros(total, np.uint16)
    lib().ref_gradients_tracker(cat, rows, cols, len(pyr), gx, gy, g2)
    return split_concat(gx, shapes), split_concat(gy, shapes), split_concat(g2, shapes)


def c2f_select(thresh: int, g2_levels):
    """g2_levels finest first -> masks finest first (bool arrays)."""
    rows, cols = g2_levels[0].shape
    shapes = [g.shape for g in g2_levels]
    cat = np.concatenate([_cm(g) for g in g2_levels]).astype(np.uint16)
    out = np.zeros(cat.size, np.uint8)
    lib().ref_c2f_select(thresh, cat, rows, cols, len(g2_levels), out)
    return [m.astype(bool) for m in split_concat(out, shapes)]


def prune_with_thresh(thresh, a, b, c, d):
    out = np.zeros(4, np.uint8)
    lib().ref_prune_with_thresh(thresh, a, b, c, d, out)
    return [bool(v) for v in out]


class Keyframe:
    def __init__(self, cfg: Config, depth: np.ndarray, img: np.ndarray, fast: bool = False, _borrowed=None):
        self._lib = lib(fast)
        self._owned = _borrowed is None
        if _borrowed is not None:
            self._h = _borrowed
        else:
            rows, cols = img.shape
            self._h = self._lib.ref_keyframe_create(C.byref(cfg), _cm(depth).astype(np.uint16), _cm(img), rows, cols)
            if not self._h:
                raise ValueError("pyramid shorter than nb_levels")

    def __del__(self):
        if getattr(self, "_owned", False) and self._h:
            self._lib.ref_keyframe_destroy(self._h)
            self._h = None

    @property
    def levels(self) -> int:
        return self._lib.ref_keyframe_levels(self._h)

    def level_shape(self, lvl):
        r, c = C.c_int(), C.c_int()
        self._lib.ref_keyframe_level_shape(self._h, lvl, C.byref(r), C.byref(c))
        return r.value, c.value

    def intrinsics(self, lvl):
        out = np.zeros(5, np.float32)
        self._lib.ref_keyframe_intrinsics(self._h, lvl, out)
        return out  # fx fy cx cy skew

    def n_points(self, lvl) -> int:
        return self._lib.ref_keyframe_n_points(self._h, lvl)

    def points(self, lvl, with_jac=True):
        n = self.n_points(lvl)
        xy = np.zeros((n, 2), np.uint32)
        idepth = np.zeros(n, np.float32)
        jac = np.zeros((n, 6), np.float32) if with_jac else None
        self._lib.ref_keyframe_points(self._h, lvl, xy.ctypes.data, idepth.ctypes.data,
                                      jac.ctypes.data if with_jac else None)
        return xy, idepth, jac

    def image(self, lvl):
        r, c = self.level_shape(lvl)
        out = np.zeros(r * c, np.uint8)
        self._lib.ref_keyframe_image(self._h, lvl, out)
        return _from_cm(out, r, c)

    def mask0(self):
        r, c = self.level_shape(0)
        out = np.zeros(r * c, np.uint8)
        self._lib.ref_keyframe_mask0(self._h, out)
        return _from_cm(out, r, c).astype(bool)

    def idepth_map(self, lvl):
        r, c = self.level_shape(lvl)
        d = np.zeros(r * c, np.float32)
        w = np.zeros(r * c, np.float32)
        self._lib.ref_keyframe_idepth_map(self._h, lvl, d, w)
        return _from_cm(d, r, c), _from_cm(w, r, c)

    def eval(self, lvl, image: np.ndarray, model: Pose, accum: int = 0):
        rows, cols = image.shape
        e = C.c_float()
        g = np.zeros(6, np.float32)
        H = np.zeros(36, np.float32)
        n = self._lib.ref_eval(self._h, lvl, _cm(image), rows, cols, C.byref(model), accum, C.byref(e), g, H)
        return e.value, n, g, H.reshape(6, 6)

    def iterative_solve(self, cfg: Config, lvl, image: np.ndarray, init: Pose, trace_cap=256):
        rows, cols = image.shape
        out = Pose()
        n_iter = C.c_int()
        en = C.c_float()
        tr = (TraceRec * trace_cap)()
        tl = C.c_int()
        st = self._lib.ref_iterative_solve(C.byref(cfg), self._h, lvl, _cm(image), rows, cols, C.byref(init),
                                           C.byref(out), C.byref(n_iter), C.byref(en), tr, trace_cap, C.byref(tl))
        return st, out, n_iter.value, en.value, [tr[i] for i in range(tl.value)]


class Tracker:
    """Oracle restatement of Config::init / Tracker::track / Tracker::current_frame."""

    def __init__(self, cfg: Config, depth_ts, depth: np.ndarray, img_ts, img: np.ndarray, fast: bool = False):
        self._lib = lib(fast)
        rows, cols = img.shape
        self.rows, self.cols = rows, cols
        # row-major numpy buffers go in with layout=1, exactly what vors_track.rs:142 does
        self._h = self._lib.ref_tracker_create(C.byref(cfg), depth_ts, np.ascontiguousarray(depth, np.uint16),
                                               img_ts, np.ascontiguousarray(img, np.uint8), rows, cols, 1)
        if not self._h:
            raise ValueError("invalid tracker configuration")

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.ref_tracker_destroy(self._h)
            self._h = None

    def track(self, depth_ts, depth, img_ts, img, trace_cap=0):
        stats = TrackStats()
        tr = (TraceRec * max(trace_cap, 1))()
        tl = C.c_int(0)
        st = self._lib.ref_tracker_track(self._h, depth_ts, np.ascontiguousarray(depth, np.uint16), img_ts,
                                         np.ascontiguousarray(img, np.uint8), C.byref(stats),
                                         tr if trace_cap else None, trace_cap, C.byref(tl))
        return st, stats, [tr[i] for i in range(tl.value)]

    def current_frame(self):
        ts = C.c_double()
        p = Pose()
        self._lib.ref_tracker_current_frame(self._h, C.byref(ts), C.byref(p))
        return ts.value, p

    def keyframe_pose(self):
        p = Pose()
        self._lib.ref_tracker_keyframe_pose(self._h, C.byref(p))
        return p

    def keyframe(self) -> Keyframe:
        return Keyframe(None, None, None, _borrowed=self._lib.ref_tracker_keyframe(self._h))


def se3_exp(xi) -> Pose:
    p = Pose()
    lib().ref_se3_exp(np.asarray(xi, np.float32), C.byref(p))
    return p


def se3_log(p: Pose) -> np.ndarray:
    xi = np.zeros(6, np.float32)
    lib().ref_se3_log(C.byref(p), xi)
    return xi


def pose_mul(a: Pose, b: Pose) -> Pose:
    o = Pose()
    lib().ref_pose_mul(C.byref(a), C.byref(b), C.byref(o))
    return o


def pose_inverse(a: Pose) -> Pose:
    o = Pose()
    lib().ref_pose_inverse(C.byref(a), C.byref(o))
    return o


def pose_error(a, b):
    """(rotation angle of qa^-1 qb in rad, |ta - tb| in m) for two 7-vectors (t, q xyzw)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    qa, qb = a[3:] / np.linalg.norm(a[3:]), b[3:] / np.linalg.norm(b[3:])
    dot = abs(float(np.dot(qa, qb)))
    ang = 2.0 * np.arccos(min(1.0, dot))
    return ang, float(np.linalg.norm(a[:3] - b[:3]))
