"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/vors_b200.h declares,
and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vors_b200.h")


@pytest.fixture(scope="module")
def vb():
    import vors_b200

    if not os.path.exists(vors_b200.LIB_PATH):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "visual-odometry-rs_b200"), "-s"])
    return vors_b200


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vors_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound(vb):
    lib = vb.load_library()
    names = _declared_symbols()
    assert len(names) >= 35
    for name in names:
        assert hasattr(lib, name), f"{name} declared in vors_b200.h but not exported"
        assert name in vb.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(vb.SIGNATURES) == names


def test_struct_layouts_match_header(vb):
    # vors_config: 23 scalar 4-byte fields (include/vors_b200.h)
    assert C.sizeof(vb.ConfigStruct) == 4 * 23
    assert C.sizeof(vb.Pose) == 28
    assert C.sizeof(vb.TraceRec) == 24
    assert C.sizeof(vb.TrackStats) == 12 + 3 * 4 * vb.MAX_LEVELS + 8


def test_default_config_is_the_reference_binary_config(vb):
    cfg = vb.Config()
    # src/bin/vors_track.rs:34-40, src/dataset/tum_rgbd.rs:15,31-35, lm_optimizer.rs:115,157,173,179,186
    assert (cfg.nb_levels, cfg.candidates_diff_threshold, cfg.depth_scale) == (6, 7, 5000.0)
    assert np.isclose(cfg.idepth_variance, 1e-4) and np.isclose(cfg.fx, 517.306408) and np.isclose(cfg.cy, 255.313989)
    assert (cfg.candidate_mode, cfg.fixed_iters, cfg.max_iters) == (0, 0, 20)
    assert np.isclose(cfg.lm_coef_init, 0.1) and cfg.lm_coef_reject_mult == 10.0 and np.isclose(cfg.lm_coef_accept_mult, 0.1)
    assert cfg.energy_delta_stop == 1.0 and cfg.keyframe_flow_threshold == 1.0


def test_pyramid_shapes_match_oracle(vb, oracle):
    for shape in [(480, 640), (1080, 1920), (37, 53), (2, 2), (1, 9), (960, 1280)]:
        for L in (1, 2, 5, 6, 8):
            assert vb.pyramid_shapes(*shape, L) == oracle.pyramid_shapes(*shape, L)


def test_no_gpu_means_loud_failure_not_fallback(vb):
    if vb.device_count() > 0:
        pytest.skip("a GPU is visible")
    img = np.zeros((48, 64), np.uint8)
    depth = np.ones((48, 64), np.uint16)
    cfg = vb.Config(nb_levels=3)
    with pytest.raises(vb.VorsError) as ei:
        cfg.init(0.0, depth, 0.0, img)
    assert ei.value.code == vb.E_CUDA
    with pytest.raises(vb.VorsError):
        vb.mean_pyramid(img, 3)
    with pytest.raises(vb.VorsError):
        vb.se3_exp(np.zeros(6))


def test_invalid_arguments_are_reported(vb):
    lib = vb.load_library()
    cfg = vb.Config(nb_levels=99)
    h = C.c_void_p()
    img = np.zeros((48, 64), np.uint8)
    depth = np.ones((48, 64), np.uint16)
    rc = lib.vors_tracker_create(C.byref(cfg.c), 0.0, depth.ctypes.data, 0.0, img.ctypes.data, 48, 64, 0, C.byref(h))
    assert rc == vb.E_INVALID and b"nb_levels" in lib.vors_last_error()
    cfg = vb.Config(nb_levels=6)  # 16x16 image: pyramid would be shorter than 6 levels -> reference panics
    rc = lib.vors_tracker_create(C.byref(cfg.c), 0.0, depth.ctypes.data, 0.0, img.ctypes.data, 16, 16, 0, C.byref(h))
    assert rc == vb.E_INVALID
    assert lib.vors_tracker_track(None, 0.0, None, 0.0, None, None) == vb.E_INVALID
