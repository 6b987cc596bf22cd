"""Parity of the code path bench.py measures (VERDICT r01 "weak" item 2): a batch that fills the device so the engine picks
team == 1 (one alignment per CTA), 640x480 dense candidates (tiled records, texture-gather sampling), 5 levels, 10 fixed LM
rounds per level, consecutive frames walked back and forth, through BOTH entry points the bench times (device-resident frames
announced one step ahead, and pinned host frames announced one step ahead), against one oracle Tracker per stream.

296 streams are 8 distinct scenes replicated 37 times: the oracle only has to track 8 streams, and the replicas double as a
determinism check (identical inputs must give bit-identical poses whatever CTA / SM they ran on).

What "parity" can mean here.  Dense candidates and a FIXED number of LM rounds are extensions; the reference stops a level as
soon as the energy gain drops below 1.0 (lm_optimizer.rs:179), long before the summation round-off matters.  Ten forced
rounds instead walk into the flat floor of the energy valley, where the accept / reject test `E' > E` compares numbers that
differ by less than their own round-off: the decisions there are coin flips in ANY arithmetic, each flip changes the damping
by a factor of 10 or 100, and the pose ends up somewhere on that floor.  The reference-faithful oracle (sequential f32 sums,
~4e-4 of relative round-off over 3e5 terms) and its f64-accumulating twin already disagree in 1-11 of 55 decisions per
alignment and by up to ~1e-4 m (`oracle_f32_vs_f64`).  So:
The floor is a valley: with a scene at ~2 m a rotation about y and a translation along x (same for x / y) move every pixel
almost alike, so the energy pins their combination 100 times worse than either - the deviations below all have
rad ~ m / depth - and a coin-flip decision at a coarse level can send a frame to another floor of it (measured: one frame of
200, 3.5e-4 m away, final energy 0.2 % higher, every single evaluation of it identical to the oracle's to 1e-6).  The bar is
therefore: at least 95 % of the alignments within the north_star's 1e-4 rad / 1e-4 m, none beyond ten times that, final
energies of every level within 1 %, at every step, for both entry points:
  * phase A: the benchmark's own camera speed, 26 steps;
  * phase B: twice that speed, 13 steps, so that streams cross the 1 px keyframe threshold (keyframe switches);
  * in both: decision statistics (same decision / strict comparator / near-ties) are reported, not hidden."""
import concurrent.futures as cf
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N_DISTINCT, N_STREAMS, N_RENDER = 8, 296, 13
ROWS, COLS = 480, 640


def _walk(k, F):
    m = k % (2 * F)
    return m if m <= F else 2 * F - m


def _run_phase(oracle, vb, bench, torch, n_steps, seed, **traj_kw):
    from test_gpu_parity import _same_trace

    device = torch.device("cuda", 0)
    cfg = bench.CONFIGS[2]
    gray, depth, _, scene = bench.make_streams(cfg, N_DISTINCT, N_RENDER, seed, device, **traj_kw)  # [F+1, 8, rows, cols]
    F = N_RENDER - 1
    rep = torch.arange(N_STREAMS, device=device) % N_DISTINCT
    repl = rep.cpu().numpy()
    gray_h, depth_h = gray.cpu(), depth.cpu()
    kw = bench.tracker_kwargs(cfg, scene)
    vcfg = vb.Config(device=0, **kw)
    ocfg = oracle.default_config(**kw)
    I = ROWS * COLS
    ts = [np.full(N_STREAMS, float(k)) for k in range(n_steps + 1)]
    status = np.zeros(N_STREAMS, np.int32)
    stats = (vb.TrackStats * N_STREAMS)()

    # ---- oracles: one Tracker per distinct stream, parity build; f64 sums (grades the GPU) and the faithful sequential f32
    def run_oracle(job):
        s, f64 = job
        oracle.lib().ref_set_accum_f64(f64)  # thread-local
        tr = oracle.Tracker(ocfg, 0.0, depth_h[0, s].numpy(), 0.0, gray_h[0, s].numpy())
        poses, traces, energies, switches = [], [], [], 0
        for k in range(1, n_steps + 1):
            f = _walk(k, F)
            _, st, trace = tr.track(float(k), depth_h[f, s].numpy(), float(k), gray_h[f, s].numpy(), trace_cap=512)
            poses.append(tr.current_frame()[1].as_array())
            traces.append(trace)
            energies.append(list(st.energy)[:5])
            switches += st.keyframe_changed
        oracle.lib().ref_set_accum_f64(0)
        return poses, traces, switches, energies

    with cf.ThreadPoolExecutor(max_workers=min(2 * N_DISTINCT, os.cpu_count() or 1)) as pool:
        both = list(pool.map(run_oracle, [(s, a) for a in (1, 0) for s in range(N_DISTINCT)]))
    ref, ref32 = both[:N_DISTINCT], both[N_DISTINCT:]
    spread = [0.0, 0.0]
    for s in range(N_DISTINCT):
        for a, b in zip(ref[s][0], ref32[s][0]):
            e = oracle.pose_error(a, b)
            spread = [max(spread[0], e[0]), max(spread[1], e[1])]

    def new_tracker():
        bt = vb.BatchTracker(vcfg, ts[0], depth_h[0].numpy()[repl], ts[0], gray_h[0].numpy()[repl], layout=vb.ROW_MAJOR)
        bt.set_tracing(True)
        return bt

    def check_step(bt, k, acc, announced=True):
        # (an unannounced host batch is aligned as two half batches while the second half uploads: two CTAs per alignment)
        assert bt.last_launch_shape()[0] == (1 if announced else 2), "the benchmarked configuration is one alignment per CTA (team == 1)"
        assert not status.any()
        _, poses = bt.current_frames()
        for s in range(N_DISTINCT):
            assert np.array_equal(poses[s::N_DISTINCT], np.broadcast_to(poses[s], poses[s::N_DISTINCT].shape)), \
                f"replicas of stream {s} differ at step {k}: the reduction is not order-deterministic"
            ang, dist = oracle.pose_error(poses[s], ref[s][0][k - 1])
            a32, d32 = oracle.pose_error(poses[s], ref32[s][0][k - 1])
            acc["max_rad"], acc["max_m"] = max(acc["max_rad"], ang), max(acc["max_m"], dist)
            acc["within"] += int(ang <= 1e-4 and dist <= 1e-4)
            acc["alignments"] += 1
            acc["vs_f32_oracle_rad"], acc["vs_f32_oracle_m"] = max(acc["vs_f32_oracle_rad"], a32), max(acc["vs_f32_oracle_m"], d32)
            assert stats[s].n_passes == 55 and list(stats[s].n_iters)[:5] == [10] * 5
            e_gpu, e_ref = np.array(list(stats[s].energy)[:5]), np.array(ref[s][3][k - 1])
            # (a frame identical to its keyframe has an energy near zero: absolute floor of 0.01 grey levels^2)
            acc["max_energy_rel"] = max(acc["max_energy_rel"], float(np.max(np.abs(e_gpu - e_ref) / (e_ref + 1.0))))
            got, want = bt.last_trace(s, 512), ref[s][1][k - 1]
            assert len(got) == len(want) == 55
            acc["records"] += len(want)
            acc["same_decision"] += sum(int((a.level, a.iter, a.accepted) == (b.level, b.iter, b.accepted)) for a, b in zip(got, want))
            try:
                acc["near_ties"] += _same_trace(got, want, max_ties=5)
                acc["strict_ok"] += 1
            except AssertionError:
                acc["diverged"] += 1

    def fresh():
        return dict(within=0, alignments=0, max_rad=0.0, max_m=0.0, vs_f32_oracle_rad=0.0, vs_f32_oracle_m=0.0, max_energy_rel=0.0, records=0,
                    same_decision=0, near_ties=0, strict_ok=0, diverged=0)

    results = {}
    # ---- (a) device-resident frames, next step announced (bench.py `value`)
    bt = new_tracker()
    acc = fresh()
    depth_i16 = depth.view(torch.int16)  # same bits; index_select has no uint16 kernel
    cm = lambda t, f: t[f].index_select(0, rep).transpose(-1, -2).contiguous()  # [296, cols, rows] = column-major frames
    nxt_g = cm(gray, _walk(1, F))
    for k in range(1, n_steps + 1):
        f = _walk(k, F)
        g, d = nxt_g, cm(depth_i16, f)
        nxt_g = cm(gray, _walk(k + 1, F)) if k < n_steps else None
        torch.cuda.synchronize()  # the library's streams are not ordered after torch's: the buffers must be complete
        bt.track_device(ts[k].ctypes.data, d.data_ptr(), ts[k].ctypes.data, g.data_ptr(), status.ctypes.data, C.addressof(stats),
                        nxt_g.data_ptr() if nxt_g is not None else None)
        check_step(bt, k, acc)
    results["device"] = acc
    del bt

    # ---- (b) pinned host frames, next step announced (bench.py `e2e`)
    gray_p, depth_p = gray_h.pin_memory(), depth_h.pin_memory()
    iptr = [(C.c_void_p * N_STREAMS)(*[gray_p[f].data_ptr() + int(r) * I for r in repl]) for f in range(F + 1)]
    dptr = [(C.c_void_p * N_STREAMS)(*[depth_p[f].data_ptr() + int(r) * I * 2 for r in repl]) for f in range(F + 1)]
    bt = new_tracker()
    acc = fresh()
    for k in range(1, n_steps + 1):
        f = _walk(k, F)
        bt.track_raw(ts[k].ctypes.data, dptr[f], ts[k].ctypes.data, iptr[f], status.ctypes.data, C.addressof(stats),
                     iptr[_walk(k + 1, F)] if k < n_steps else None)
        check_step(bt, k, acc, announced=k > 1)
    results["host_announced"] = acc
    del bt
    results["oracle_f32_vs_f64"] = dict(max_rad=spread[0], max_m=spread[1])
    results["oracle_keyframe_switches"] = int(sum(r[2] for r in ref))
    results["steps"] = n_steps
    return results


def test_benchmarked_path_team1_dense_296_streams(oracle, record_property):
    import torch

    import bench
    import vors_b200 as vb

    assert vb.device_count() > 0
    # phase A: the benchmark's own camera speed
    a = _run_phase(oracle, vb, bench, torch, 26, 424200)
    print("phase A (bench camera speed) vs oracle:", a)
    record_property("benchmarked_path_parity_phase_a", a)
    # phase B: twice the speed, so that every stream crosses the 1 px keyframe threshold during the walk
    b = _run_phase(oracle, vb, bench, torch, 13, 424300, step_v=0.008, step_w=0.006)
    print("phase B (2x speed, keyframe switches) vs oracle:", b)
    record_property("benchmarked_path_parity_phase_b", b)
    for name, ph in (("A", a), ("B", b)):
        for arm in ("device", "host_announced"):
            r = ph[arm]
            # >= 95 % of the alignments within the north_star bar, none beyond ten times it (a basin flip of the fixed-round LM)
            assert r["within"] >= 0.95 * r["alignments"], (name, arm, r)
            assert r["max_rad"] <= 1e-3 and r["max_m"] <= 1e-3, (name, arm, r)
            assert r["max_energy_rel"] <= 1e-2, (name, arm, r)  # another floor of the valley is within 1 % in energy
            assert r["same_decision"] >= 0.8 * r["records"], f"phase {name} {arm}: only {r['same_decision']} of {r['records']} LM decisions agree"
    assert b["oracle_keyframe_switches"] >= 1, "phase B must include keyframe switches"
