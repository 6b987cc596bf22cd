"""N>1 host logic on CPU: world_size-2 gloo run of the stream sharding + pose all-gather.
The per-stream compute is the CPU oracle here (no GPU in this test); the collective and ordering are the product's."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vors_b200 import shard, synth

N_STREAMS = 5


def _track_streams(oracle, idx):
    out = []
    for s in idx:
        scene, frames, _ = synth.make_sequence(seed=900 + int(s), n_frames=3, rows=60, cols=80, step_v=0.01, step_w=0.005)
        cfg = oracle.default_config(nb_levels=3, **synth.scene_config_kwargs(scene))
        tr = oracle.Tracker(cfg, 0.0, frames[0][1], 0.0, frames[0][0])
        st = 0
        for k in (1, 2):
            st, _, _ = tr.track(float(k), frames[k][1], float(k), frames[k][0])
        out.append(np.concatenate([tr.current_frame()[1].as_array(), [st]]))
    return np.asarray(out, np.float32).reshape(-1, 8)


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "visual-odometry-rs_b200"))
    from oracle import oracle_py

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx = shard.partition(N_STREAMS, rank, world)
    local = _track_streams(oracle_py, idx)
    full = shard.gather_poses(local, N_STREAMS)
    # the asynchronous, preallocated form used by bench.py: two exchanges in flight, collected in submission order
    pg = shard.PoseGatherer(N_STREAMS, device="cpu", depth=2)
    pg.submit(local[:, :7], local[:, 7])
    pg.submit(local[:, :7] + 1.0, local[:, 7])
    a, b = pg.collect(), pg.collect()
    assert pg.in_flight() == 0
    assert np.array_equal(a, full), "PoseGatherer differs from gather_poses"
    assert np.array_equal(b[:, :7], full[:, :7] + 1.0) and np.array_equal(b[:, 7], full[:, 7])
    # bench.py's pattern: the exchange is submitted from a helper thread while the main thread works; every main-thread
    # collective first waits for the pending exchange, so that all ranks enqueue collectives in the same order
    import concurrent.futures as cf
    pool = cf.ThreadPoolExecutor(max_workers=1)
    pg2 = shard.PoseGatherer(N_STREAMS, device="cpu", depth=2)
    pending = None
    got = []
    def work(p, st):
        if pg2.in_flight() == pg2.depth:
            got.append(pg2.collect())
        pg2.submit(p, st)
    for k in range(5):
        if pending is not None:
            pending.result()
        pending = pool.submit(work, local[:, :7] + float(k), local[:, 7].copy())
        if k == 2:  # a main-thread collective in the middle of the loop (bench.py: the barrier before the timed region)
            pending.result()
            dist.barrier()
    pending.result()
    while pg2.in_flight():
        got.append(pg2.collect())
    assert len(got) == 5 and all(np.array_equal(g[:, :7], full[:, :7] + float(k)) for k, g in enumerate(got))
    # the grouped form bench.py steps with: 3 steps per block, 7 steps (so the last block is partial), a main-thread
    # collective in the middle, 4 streams per rank; record (step k, global stream s) carries k + s / 100
    n_loc, G = 4, 3
    xch = shard.StepExchange(n_loc, device="cpu", group=G, keep=True)
    mine = rank + world * np.arange(n_loc)
    for k in range(7):
        p = np.repeat((k + mine / 100.0).astype(np.float32)[:, None], 7, 1)
        xch.push(p, (mine % 2).astype(np.float32))
        if k == 3:
            xch.wait_enqueued()
            dist.barrier()
    xch.drain()
    assert xch.blocks == 3 and len(xch.collected) == 3
    every = np.arange(n_loc * world)
    for blk_i, blk in enumerate(xch.collected):
        assert blk.shape == (G, n_loc * world, 8)
        for gstep in range(G):
            k = blk_i * G + gstep
            want = (k + every / 100.0).astype(np.float32) if k < 7 else np.zeros(n_loc * world, np.float32)
            assert np.array_equal(blk[gstep, :, 0], want) and np.array_equal(blk[gstep, :, 6], want), (blk_i, gstep)
            assert np.array_equal(blk[gstep, :, 7], (every % 2).astype(np.float32) if k < 7 else np.zeros(n_loc * world, np.float32))
    q.put((rank, full))
    dist.barrier()
    dist.destroy_process_group()


def test_partition_covers_every_item_once():
    for n in (1, 5, 8, 296):
        for world in (1, 2, 4, 8):
            got = np.sort(np.concatenate([shard.partition(n, r, world) for r in range(world)]))
            assert np.array_equal(got, np.arange(n))


@pytest.mark.parametrize("world", [2, 4])
def test_gloo_gather_matches_single_process(oracle, world):
    """world 2 (the contract's CPU test) and world 4 (more ranks than some shards have streams: 5 streams over 4 ranks)."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    single = _track_streams(oracle, np.arange(N_STREAMS))
    for r in range(1, world):
        assert np.array_equal(results[0], results[r])  # every rank holds the same gathered table
    assert np.array_equal(results[0], single)      # in global stream order, identical to the unsharded run
    assert results[0].shape == (N_STREAMS, 8) and np.all(results[0][:, 7] == 0)


def test_pose_gatherer_single_process_is_identity():
    pg = shard.PoseGatherer(7, device="cpu")
    poses = np.arange(49, dtype=np.float32).reshape(7, 7)
    status = np.arange(7, dtype=np.int32) % 2
    pg.submit(poses, status)
    out = pg.collect()
    assert np.array_equal(out[:, :7], poses) and np.array_equal(out[:, 7], status.astype(np.float32))
    with pytest.raises(RuntimeError):
        pg.collect()


def test_step_exchange_single_process_blocks():
    """Without a process group the grouped exchange is the identity: blocks of `group` steps, the last one zero-padded."""
    xch = shard.StepExchange(3, device="cpu", group=2, keep=True)
    for k in range(5):
        xch.push(np.full((3, 7), float(k + 1), np.float32), np.array([0, 1, 0], np.float32))
        xch.wait_enqueued()  # (what a caller does before a collective of its own)
    xch.drain()
    assert xch.blocks == 3 and [b.shape for b in xch.collected] == [(2, 3, 8)] * 3
    steps = np.concatenate([b[:, 0, 0] for b in xch.collected])
    assert np.array_equal(steps, [1, 2, 3, 4, 5, 0])
    assert np.array_equal(xch.collected[0][1, :, 7], [0, 1, 0]) and np.all(xch.collected[2][1] == 0)
    xch.drain()  # idempotent
    assert xch.blocks == 3
