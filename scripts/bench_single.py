"""Single-stream latency: one Tracker (reference configuration: coarse-to-fine candidates, adaptive LM) on a synthetic
640x480 sequence, GPU (vors_tracker_track through the C ABI, host buffers) vs the CPU oracle on the same frames.
BASELINE configs[0].  Prints one JSON line."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "visual-odometry-rs_b200"))
import numpy as np
import vors_b200 as vb
from oracle import oracle_py as O
from vors_b200 import synth

def run(levels, n_frames, rows, cols, **kw):
    scene, frames, poses = synth.make_sequence(seed=77, n_frames=n_frames, rows=rows, cols=cols)
    base = dict(nb_levels=levels, **synth.scene_config_kwargs(scene)); base.update(kw)
    t = vb.Config(**base).init(0.0, frames[0][1], 0.0, frames[0][0])
    ot = O.Tracker(O.default_config(**base), 0.0, frames[0][1], 0.0, frames[0][0], fast=True)   # -O3 build: timing only
    op = O.Tracker(O.default_config(**base), 0.0, frames[0][1], 0.0, frames[0][0], fast=False)  # parity build: pose check
    gpu, cpu, errs, passes, switches = [], [], [], [], 0
    for k in range(1, n_frames):
        g, d = frames[k]
        t0 = time.perf_counter(); st = t.track(float(k), d, float(k), g); gpu.append(time.perf_counter() - t0)
        t0 = time.perf_counter(); ot.track(float(k), d, float(k), g); cpu.append(time.perf_counter() - t0)
        op.track(float(k), d, float(k), g)
        errs.append(O.pose_error(t.current_frame()[1].as_array(), op.current_frame()[1].as_array()))
        passes.append(st.n_passes); switches += st.keyframe_changed
    g, c = np.median(gpu[2:]) * 1e3, np.median(cpu[2:]) * 1e3
    return dict(shape=[rows, cols], levels=levels, mode=kw.get("candidate_mode", 0), gpu_ms=g, cpu_ms=c, gpu_fps=1e3 / g, cpu_fps=1e3 / c,
                speedup=c / g, max_pose_diff_rad=max(e[0] for e in errs), max_pose_diff_m=max(e[1] for e in errs),
                median_passes=float(np.median(passes)), keyframe_switches=switches, frames=n_frames - 1)

if __name__ == "__main__":
    out = [run(5, 24, 480, 640), run(6, 24, 480, 640), run(6, 12, 960, 1280), run(6, 10, 1080, 1920),
           run(5, 12, 480, 640, candidate_mode=1, fixed_iters=10)]
    print(json.dumps(out))
