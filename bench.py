#!/usr/bin/env python
"""bench.py — RGB-D frames aligned / s at 640x480, 5 pyramid levels (BASELINE.json metric).

Workload (BASELINE.json configs[1]): per GPU, B independent synthetic 640x480
RGB-D streams, dense candidates (every pixel with depth != 0), 5 levels, 10 fixed LM rounds per level
(11 energy passes / level), Tracker semantics (keyframe switch when the optical flow reaches 1 px).
One step = every stream tracks its next frame = B alignments in ONE persistent kernel launch.

  value   frames/s with the step's inputs already resident in HBM (vors_batch_track_device)
  e2e     frames/s through the C ABI with HOST buffers (vors_batch_track_next: every call announces the next step's frames
          so that their upload overlaps the alignment): pinned row-major frames in,
          poses out, H2D/D2H copies inside the timed region
  roofline   align kernel: algorithmic bytes (10 B per candidate-pass) / its device time
  cpu_baseline  the C++ oracle (a restatement of the reference's algorithm, not rustc output), 1 core,
          on a bounded sample of the same streams
  --impl reference   the same oracle on all host cores (independent streams, one per thread)

Launch: python bench.py --gpus N --steps K --warmup W   (N>1: torchrun, one rank per GPU, NCCL pose gather)
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "visual-odometry-rs_b200"))

ROWS, COLS, LEVELS, FIXED_ITERS = 480, 640, 5, 10
METRIC = "RGB-D frames aligned/sec at 640x480, 5 pyramid levels"
WORKLOAD = ("configs[1]: 640x480 synthetic RGB-D sequences, dense (all-pixel) candidates, 5 levels, "
            "10 fixed LM rounds/level, Tracker semantics incl. keyframe switches")
MAX_FRAMES = 12  # distinct synthetic frames rendered per stream
ALGO_BYTES_PER_POINT_PASS = 10.0  # idepth 4 + template 1 + gradient pair 4 + image texel 1


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=296, help="independent RGB-D streams per GPU (2 CTAs x 148 SMs)")
    ap.add_argument("--team", type=int, default=0, help="CTAs per alignment (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_streams(n_streams, n_frames, seed0, device):
    """Per stream: its own textured-plane scene and smooth trajectory.  Returns torch tensors
    gray u8 [n_frames, n_streams, ROWS, COLS], depth u16 (same shape), gt poses [n_frames, n_streams, 7], scene0."""
    import torch
    from vors_b200 import synth

    gray = torch.empty((n_frames, n_streams, ROWS, COLS), dtype=torch.uint8, device=device)
    depth = torch.empty((n_frames, n_streams, ROWS, COLS), dtype=torch.uint16, device=device)
    gt = np.zeros((n_frames, n_streams, 7))
    scene0 = None
    for s in range(n_streams):
        scene = synth.make_scene(seed0 + s, ROWS, COLS)
        scene0 = scene0 or scene
        poses = synth.trajectory(seed0 + s, n_frames)
        g, d = synth.render_batch_torch(scene, poses, device, frame_seed=s, chunk=n_frames)
        gray[:, s] = g
        depth[:, s] = d.to(torch.uint16)
        for k, p in enumerate(poses):
            gt[k, s, :3], gt[k, s, 3:] = p[0], p[1]
    return gray, depth, gt, scene0


def ptr_array(base_ptr, n, stride_bytes):
    return (C.c_void_p * n)(*[base_ptr + i * stride_bytes for i in range(n)])


def run_ours(args):
    import torch
    import torch.distributed as dist

    import vors_b200 as vb
    from vors_b200 import synth

    rank, world, local = dist_env()
    assert torch.cuda.is_available(), "bench.py (impl=ours) needs a GPU: libvors_b200 has no CPU fallback"
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    if not os.path.exists(vb.LIB_PATH):
        raise SystemExit(f"{vb.LIB_PATH} missing: run __graft_entry__.build() first")

    B, K, W = args.streams, args.steps, args.warmup
    T = K + W
    # At most MAX_FRAMES distinct frames per stream are rendered (3.3 GB of device and of pinned host memory at 296 streams);
    # longer runs walk the rendered trajectory back and forth, so consecutive steps always see adjacent frames.
    F = min(T, MAX_FRAMES - 1)
    def fi(k):
        m = k % (2 * F)
        return m if m <= F else 2 * F - m
    t0 = time.time()
    gray, depth, gt, scene = make_streams(B, F + 1, 100000 * (rank + 1), device)
    # device-resident inputs in the library's internal layout (column-major per frame)
    gray_cm = gray.transpose(-1, -2).contiguous()
    depth_cm = depth.transpose(-1, -2).contiguous()
    # host inputs: pinned, row-major (decoder layout, what src/bin/vors_track.rs:140-145 gets from PNG files)
    gray_h = torch.empty(gray.shape, dtype=torch.uint8, pin_memory=True)
    depth_h = torch.empty(depth.shape, dtype=torch.uint16, pin_memory=True)
    gray_h.copy_(gray)
    depth_h.copy_(depth)
    torch.cuda.synchronize()
    gen_s = time.time() - t0
    I = ROWS * COLS

    kw = dict(nb_levels=LEVELS, candidate_mode=vb.CANDIDATES_DENSE, fixed_iters=FIXED_ITERS, device=local,
              team_size=args.team, **synth.scene_config_kwargs(scene))
    cfg = vb.Config(**kw)

    ts = [np.full(B, float(k)) for k in range(T + 1)]
    status = np.zeros(B, np.int32)
    stats = (vb.TrackStats * B)()
    from vors_b200 import shard

    def gather_poses(bt):
        """Streams shard across ranks with no data-path collective; the only exchange is this all-gather of 32-byte
        pose records (NCCL), once per step."""
        if world > 1:
            _, p = bt.current_frames()
            return shard.gather_poses(shard.pack_records(p, status), B * world, device=device)
        return None

    def new_tracker():
        g0 = gray_h[0].numpy()
        d0 = depth_h[0].numpy()
        return vb.BatchTracker(cfg, ts[0], d0, ts[0], g0, layout=vb.ROW_MAJOR)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    sampler = ClockSampler(local)

    # ---- arm 1: inputs resident in HBM ------------------------------------------------------------------
    bt = new_tracker()
    def step_device(k):
        # the next step's device buffer is announced: its copy into the frame pyramids and the pyramid build overlap this alignment
        return bt.track_device(ts[k].ctypes.data, depth_cm[fi(k)].data_ptr(), ts[k].ctypes.data, gray_cm[fi(k)].data_ptr(),
                               status.ctypes.data, C.addressof(stats), gray_cm[fi(k + 1)].data_ptr() if k < T else None)
    for k in range(1, W + 1):
        step_device(k)
        gather_poses(bt)
    align_ms = pyr_ms = kf_ms = up_ms = 0.0
    launches = point_passes = switches = failed = 0
    barrier()
    if rank == 0:
        sampler.start()
    t_start = time.perf_counter()
    for k in range(W + 1, T + 1):
        step_device(k)
        gather_poses(bt)
        tm = bt.last_timing()
        align_ms += tm["align_ms"]; pyr_ms += tm["pyramid_ms"]; kf_ms += tm["keyframe_ms"]; up_ms += tm["upload_ms"]
        l, pp = bt.last_counters()
        launches += l; point_passes += pp
        switches += sum(s.keyframe_changed for s in stats)
        failed += int((status != 0).sum())
    barrier()
    dev_s = max_over_ranks(time.perf_counter() - t_start)
    clocks = sampler.stop() if rank == 0 else None
    _, poses_a = bt.current_frames()
    del bt

    # ---- arm 2: end to end through the C ABI with host buffers -----------------------------------------
    bt = new_tracker()
    img_ptrs = [ptr_array(gray_h[k].data_ptr(), B, I) for k in range(F + 1)]
    dep_ptrs = [ptr_array(depth_h[k].data_ptr(), B, I * 2) for k in range(F + 1)]
    # every call announces the next step's host frames (a streaming caller has them decoded by then): their H2D copy,
    # transpose and pyramid build overlap this step's alignment.  Still one H2D of every frame per step, inside the timed
    # region; the last step has nothing to announce.
    nxt = lambda k: img_ptrs[fi(k + 1)] if k < T else None
    for k in range(1, W + 1):
        bt.track_raw(ts[k].ctypes.data, dep_ptrs[fi(k)], ts[k].ctypes.data, img_ptrs[fi(k)], status.ctypes.data, C.addressof(stats), nxt(k))
        gather_poses(bt)
    e2e_switches = 0
    barrier()
    t_start = time.perf_counter()
    for k in range(W + 1, T + 1):
        bt.track_raw(ts[k].ctypes.data, dep_ptrs[fi(k)], ts[k].ctypes.data, img_ptrs[fi(k)], status.ctypes.data, C.addressof(stats), nxt(k))
        bt.current_frames()  # device -> host read of the step's result (poses) is part of the call above; this is the accessor
        gather_poses(bt)
        e2e_switches += sum(s.keyframe_changed for s in stats)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t_start)
    _, poses_b = bt.current_frames()
    del bt

    frames = B * K * world
    value = frames / dev_s
    e2e_value = frames / e2e_s
    # accuracy of the run itself (not a parity claim): error of the tracked poses against the synthetic ground truth
    def err(p):
        q = gt[fi(T)][:, 3:]
        d = np.abs(np.sum(p[:, 3:] * q, 1)).clip(max=1.0)
        return float(np.max(2 * np.arccos(d))), float(np.max(np.linalg.norm(p[:, :3] - gt[fi(T)][:, :3], axis=1)))
    rot_err, trans_err = err(poses_a)

    out = None
    if rank == 0:
        peak, peak_src = peaks()
        algo_bytes = ALGO_BYTES_PER_POINT_PASS * point_passes  # rank 0's launches
        achieved = algo_bytes / (align_ms * 1e-3) / 1e9 if align_ms > 0 else 0.0
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "align_traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        out = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_s / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "streams_per_gpu": B, "frames_per_step": B * world,
                       "inputs": f"larger than L2: {B * 12.3e-3:.1f} GB of per-stream keyframe+frame data touched per step",
                       "frames": f"{F + 1} rendered frames per stream, walked back and forth",
                       "team_size": args.team or "auto", "keyframe_switches_per_step": switches / K,
                       "failed_alignments": failed, "pose_gather": "NCCL all_gather per step" if world > 1 else "none (1 GPU)",
                       "max_pose_error_vs_ground_truth": {"rad": rot_err, "m": trans_err},
                       "arms_max_abs_pose_diff": float(np.max(np.abs(poses_a - poses_b))), "synth_seconds": gen_s},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "frames/s", "ms_per_step": e2e_s / K * 1e3,
                    "h2d_bytes_per_step": B * I + (e2e_switches / K) * I * 2 + B * 28,
                    "d2h_bytes_per_step": B * 272 + B * 32,
                    "host_layout": "row-major pinned (decoder layout); depth uploaded only on keyframe switches"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "k_align (persistent LM alignment)", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": algo_bytes / K, "avg_launch_ms": align_ms / K,
                         "point_passes_per_launch": point_passes / K,
                         "step_share": {"upload_ms": up_ms / K, "pyramid_ms": pyr_ms / K, "align_ms": align_ms / K,
                                        "keyframe_ms": kf_ms / K}},
        }
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline_port(gray_h, depth_h, kw, n_streams=9, n_frames=F)  # ~10 s of CPU work
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


def oracle_cfg(kw):
    from oracle import oracle_py as O

    kw = {k: v for k, v in kw.items() if k not in ("device", "team_size")}
    return O.default_config(**kw)


def cpu_baseline_port(gray_h, depth_h, kw, n_streams, n_frames):
    """C++ oracle (-O3 build), ONE thread (the reference is single-threaded), bounded sample of the same streams."""
    from oracle import oracle_py as O

    O.build()
    cfg = oracle_cfg(kw)
    total, frames = 0.0, 0
    for s in range(n_streams):
        tr = O.Tracker(cfg, 0.0, depth_h[0, s].numpy(), 0.0, gray_h[0, s].numpy(), fast=True)
        for k in range(1, n_frames + 1):
            g, d = gray_h[k, s].numpy(), depth_h[k, s].numpy()
            t0 = time.perf_counter()
            tr.track(float(k), d, float(k), g)
            total += time.perf_counter() - t0
            frames += 1
    return {"value": frames / total, "unit": "frames/s", "cores": 1, "kind": "port",
            "sample": f"{n_streams} streams x {n_frames} frames of the same workload (Tracker::track only, PNG decode excluded); "
                      "C++ restatement of the reference algorithm, not rustc output", "seconds": total}


def run_reference(args):
    """The reference's CPU path = the C++ oracle port (the Rust reference cannot be built in this image: no
    cargo/rustc), all host cores, one independent stream per thread, bounded sample per step."""
    rank, world, local = dist_env()
    if rank != 0:
        return None
    import concurrent.futures as cf

    from oracle import oracle_py as O
    from vors_b200 import synth

    O.build()
    cores = os.cpu_count() or 1
    K, W = args.steps, args.warmup
    T = K + W
    try:
        import torch
        device = torch.device("cuda", local) if torch.cuda.is_available() else torch.device("cpu")
    except Exception:
        device = None
    n = cores
    gray, depth, gt, scene = make_streams(n, T + 1, 100000, device)
    gray = gray.cpu().numpy()
    depth = depth.cpu().numpy()
    kw = dict(nb_levels=LEVELS, candidate_mode=1, fixed_iters=FIXED_ITERS, **synth.scene_config_kwargs(scene))
    cfg = oracle_cfg(kw)
    trackers = [O.Tracker(cfg, 0.0, depth[0, s], 0.0, gray[0, s], fast=True) for s in range(n)]
    pool = cf.ThreadPoolExecutor(max_workers=cores)

    def step(k):
        list(pool.map(lambda s: trackers[s].track(float(k), depth[k, s], float(k), gray[k, s]), range(n)))

    for k in range(1, W + 1):
        step(k)
    t0 = time.perf_counter()
    for k in range(W + 1, T + 1):
        step(k)
    secs = time.perf_counter() - t0
    value = n * K / secs
    return {"impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus, "steps": K, "warmup": W,
            "ms_per_step": secs / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "streams": n, "note": "each step = one frame on each of `cores` independent streams"},
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port",
                             "sample": f"{n} independent streams x {K} timed frames, one stream per thread; C++ restatement "
                                       "of the reference algorithm (-O3), not rustc output"},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def main():
    args = parse_args()
    # stdout carries exactly one JSON line: anything a library writes to fd 1 meanwhile (NCCL prints its version banner
    # there when NCCL_DEBUG is set) goes to stderr instead
    sys.stdout.flush()
    result_fd = os.dup(1)
    os.dup2(2, 1)
    out = run_reference(args) if args.impl == "reference" else run_ours(args)
    sys.stdout.flush()
    if out is not None:
        os.write(result_fd, (json.dumps(out) + "\n").encode())
    os.close(result_fd)


if __name__ == "__main__":
    main()
