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


class PoseGatherer:
    """The per-step pose exchange without a host round trip on the critical path.

    `gather_poses` above is the simple synchronous form (tests, one-off calls).  In a stepping loop it serialises a
    H2D copy, the collective and `world` blocking device->host copies between two launches of the align kernel, with
    nothing in flight on the GPU meanwhile.  This class keeps everything preallocated and asynchronous instead:

      submit(records)  packs the step's records into a pinned host slot, copies them to a preallocated device buffer,
                       issues ONE `all_gather_into_tensor` (NCCL; gloo falls back to the list form on CPU tensors) and
                       ONE device->host copy of the gathered block into pinned memory, all on a side stream, and
                       returns immediately - the caller launches its next step while the exchange is in flight;
      collect()        waits for the oldest outstanding exchange and returns its records in GLOBAL item order.

    `depth` slots are in flight at most (default 2: the gather of step k overlaps the kernels of step k+1).
    Ranks own `partition(n_items, rank, world)`; shards are padded to the largest shard.
    """

    def __init__(self, n_items: int, device=None, depth: int = 2):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.n_items = int(n_items)
        self.active = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.world = dist.get_world_size() if self.active else 1
        self.rank = dist.get_rank() if self.active else 0
        self.per = (self.n_items + self.world - 1) // self.world
        self.n_local = len(partition(self.n_items, self.rank, self.world))
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        self.cuda = self.device.type == "cuda"
        self.depth = max(1, int(depth))
        pin = self.cuda
        self.h_in = [torch.zeros((self.per, POSE_RECORD_FLOATS), dtype=torch.float32, pin_memory=pin) for _ in range(self.depth)]
        self.h_out = [torch.zeros((self.world * self.per, POSE_RECORD_FLOATS), dtype=torch.float32, pin_memory=pin)
                      for _ in range(self.depth)]
        if self.cuda:
            self.d_in = [torch.zeros((self.per, POSE_RECORD_FLOATS), dtype=torch.float32, device=self.device) for _ in range(self.depth)]
            self.d_out = [torch.zeros((self.world * self.per, POSE_RECORD_FLOATS), dtype=torch.float32, device=self.device)
                          for _ in range(self.depth)]
            self.stream = torch.cuda.Stream(device=self.device)
            self.done = [torch.cuda.Event() for _ in range(self.depth)]
        # global item index of row j of the gathered block: rank r's i-th record is item r + i * world
        idx = np.full(self.world * self.per, -1, np.int64)
        for r in range(self.world):
            own = partition(self.n_items, r, self.world)
            idx[r * self.per: r * self.per + len(own)] = own
        self._rows = np.nonzero(idx >= 0)[0]
        self._items = idx[self._rows]
        self._head = 0          # next slot to submit into
        self._pending = []      # slots in flight, oldest first
        self.submitted = 0

    def submit(self, poses: np.ndarray, status: np.ndarray) -> None:
        """Start the exchange of this step's local records (poses [n_local, 7], status [n_local])."""
        torch, dist = self.torch, self.dist
        if len(self._pending) == self.depth:
            raise RuntimeError("PoseGatherer: collect() the oldest exchange before submitting another")
        s = self._head
        self._head = (s + 1) % self.depth
        h = self.h_in[s].numpy()
        h[: self.n_local, :7] = poses
        h[: self.n_local, 7] = status
        if not self.active:
            self.h_out[s][: self.per].copy_(self.h_in[s])
        elif self.cuda:
            with torch.cuda.stream(self.stream):
                self.d_in[s].copy_(self.h_in[s], non_blocking=True)
                dist.all_gather_into_tensor(self.d_out[s], self.d_in[s])
                self.h_out[s].copy_(self.d_out[s], non_blocking=True)
                self.done[s].record(self.stream)
        else:  # gloo on CPU tensors (tests)
            parts = list(self.h_out[s].view(self.world, self.per, POSE_RECORD_FLOATS).unbind(0))
            dist.all_gather(parts, self.h_in[s])
        self._pending.append(s)
        self.submitted += 1

    def collect(self) -> np.ndarray:
        """Records of the oldest outstanding exchange in global item order: [n_items, 8] (a fresh array)."""
        if not self._pending:
            raise RuntimeError("PoseGatherer: nothing in flight")
        s = self._pending.pop(0)
        if self.active and self.cuda:
            self.done[s].synchronize()
        full = np.zeros((self.n_items, POSE_RECORD_FLOATS), np.float32)
        full[self._items] = self.h_out[s].numpy()[self._rows]
        return full

    def in_flight(self) -> int:
        return len(self._pending)


class StepExchange:
    """The stepping loop's side of the pose exchange: local records of `group` consecutive steps are accumulated on the host
    and exchanged as ONE block ("or every K steps"), through a `PoseGatherer` driven from a helper thread, so
    that neither the collective nor its host work sits between two launches of the align kernel.

      push(poses, status)   after every step (main thread); every `group`-th call hands the block to the helper thread
      wait_enqueued()       the main thread calls this before any collective of its own: all ranks must enqueue their
                            collectives in the same order, and the helper's exchange comes first
      drain()               flushes a partial block and collects everything in flight

    Every rank owns `n_local` streams (weak scaling); stream j of rank r is global stream r + j * world, as in `partition`.
    A collected block is [group, n_local * world, 8] in global stream order (rows of steps not reached yet are zero)."""

    def __init__(self, n_local: int, device=None, group: int = 4, depth: int = 2, thread_init=None, keep: bool = False):
        import concurrent.futures as cf
        import torch.distributed as dist

        world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        self.n_local, self.group, self.world = int(n_local), max(1, int(group)), world
        self.n_step = self.n_local * world
        # items = (step of the group, global stream): item % world is the owner, so `partition` hands every rank its own
        # streams of every step, step-major - the order of the accumulation buffer below
        self.pg = PoseGatherer(self.n_step * self.group, device=device, depth=depth)
        assert self.pg.n_local == self.n_local * self.group
        self._p = np.zeros((self.group, self.n_local, 7), np.float32)
        self._s = np.zeros((self.group, self.n_local), np.float32)
        self._fill = 0
        self._pool = cf.ThreadPoolExecutor(max_workers=1)
        self._pending = None
        self._thread_init = thread_init
        self.keep = keep
        self.collected = []  # blocks, oldest first (only with keep=True)
        self.blocks = 0      # blocks collected so far

    def _collect_one(self):
        full = self.pg.collect().reshape(self.group, self.n_step, POSE_RECORD_FLOATS)
        self.blocks += 1
        if self.keep:
            self.collected.append(full)

    def _work(self, p, s):
        if self._thread_init is not None:
            self._thread_init()
        if self.pg.in_flight() == self.pg.depth:
            self._collect_one()
        self.pg.submit(p, s)

    def push(self, poses: np.ndarray, status: np.ndarray) -> None:
        self._p[self._fill] = poses
        self._s[self._fill] = status
        self._fill += 1
        if self._fill == self.group:
            self.flush()

    def flush(self) -> None:
        if self._fill == 0:
            return
        self._p[self._fill:] = 0.0
        self._s[self._fill:] = 0.0
        self.wait_enqueued()
        self._pending = self._pool.submit(self._work, self._p.reshape(-1, 7).copy(), self._s.reshape(-1).copy())
        self._fill = 0

    def wait_enqueued(self) -> None:
        if self._pending is not None:
            self._pending.result()
            self._pending = None

    def drain(self) -> None:
        self.flush()
        self.wait_enqueued()
        while self.pg.in_flight():
            self._collect_one()
