"""Multi-GPU sharding of independent alignments.

The path shards across alignments only: a single `track` is a serial chain of small reductions and consecutive
frames of one stream are coupled through the pose prior (inverse_compositional.rs:177, :224-239).  So each rank
owns whole streams, runs them on its own GPU with no data-path collective, and the ONLY exchange is one
all-gather of the fixed-size pose records (7 f32 + status, 32 B per stream) per step.  torch.distributed is
plumbing: NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np

POSE_RECORD_FLOATS = 8  # t[3], q[4] (x y z w), status


def partition(n_items: int, rank: int, world: int) -> np.ndarray:
    """Round-robin ownership: rank r takes items r, r + world, ...  (SURVEY §8e)."""
    return np.arange(rank, n_items, world, dtype=np.int64)


def pack_records(poses: np.ndarray, status: np.ndarray) -> np.ndarray:
    rec = np.zeros((poses.shape[0], POSE_RECORD_FLOATS), np.float32)
    rec[:, :7] = poses
    rec[:, 7] = status
    return rec


def gather_poses(local_records, n_items: int, device=None):
    """All-gather the per-rank pose records and return them in GLOBAL item order: [n_items, 8].

    local_records: [len(partition(n_items, rank, world)), 8] float32 (numpy or torch).  Ranks may own different
    counts (n_items not divisible by world), so records are padded to the largest shard before the collective."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        rec = torch.as_tensor(local_records, dtype=torch.float32)
        return rec.cpu().numpy()
    world, rank = dist.get_world_size(), dist.get_rank()
    per = (n_items + world - 1) // world
    buf = torch.zeros((per, POSE_RECORD_FLOATS), dtype=torch.float32, device=device)
    loc = torch.as_tensor(local_records, dtype=torch.float32)
    buf[: loc.shape[0]].copy_(loc)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    full = np.zeros((n_items, POSE_RECORD_FLOATS), np.float32)
    for r in range(world):
        idx = partition(n_items, r, world)
        full[idx] = out[r][: len(idx)].cpu().numpy()
    return full
