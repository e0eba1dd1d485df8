"""Multi-GPU plumbing: one process per GPU, the batch axis sharded in contiguous slices, no communication on the
solve path, one all-gather of the packed per-instance result record afterwards.

Only `torch.distributed` is used (NCCL over NVLink on the GPU box, gloo in the CPU tests); the send buffer is the
record buffer the step kernel's epilogue wrote, so there is no staging copy.
"""
from __future__ import annotations

from typing import List, Optional, Tuple


def shard_bounds(B: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of a batch of B instances owned by `rank`; sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(B: int, world: int) -> List[int]:
    return [shard_bounds(B, world, r)[1] - shard_bounds(B, world, r)[0] for r in range(world)]


def allgather_records(local, B: Optional[int] = None, group=None):
    """All-gather per-instance records.  `local` is this rank's [b_r, W] tensor (CUDA for NCCL, CPU for gloo).
    With equal shard sizes this is a single `all_gather_into_tensor` straight out of `local`; ragged shards are
    padded to the largest shard first.  Returns the [B, W] tensor in global instance order on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1:
        return local
    b_r, W = local.shape
    if B is None:
        B = b_r * world
    sizes = shard_sizes(B, world)
    if sizes[dist.get_rank(group)] != b_r:
        raise ValueError("local shard size does not match shard_bounds()")
    m = max(sizes)
    if all(s == m for s in sizes):
        out = torch.empty(world * m, W, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    send = local if b_r == m else torch.cat([local, local.new_zeros(m - b_r, W)], 0)
    buf = torch.empty(world * m, W, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, send.contiguous(), group=group)
    return torch.cat([buf[r * m:r * m + sizes[r]] for r in range(world)], 0)


class FusedRecordGather:
    """All-gather of the result records fused into the step kernel (no collective call on the data path).

    Every rank allocates `buffers` gathered tables [rows, RECORD_LEN] in symmetric memory; after the rendezvous each
    GPU holds NVLink peer pointers to all of them, and `BatchedMPC.set_record_peers` makes the step kernel's epilogue
    store each instance's record straight into row `row_offset + b` of the current table on every GPU.

    Synchronisation is separate from the stores and is the caller's choice:
      * `finish(sync=True)` after a step: one cross-GPU barrier (symmetric-memory signal pads); afterwards the table
        of that step is complete on every rank.  The tables rotate, so the next step's stores never land in the
        table a slower peer may still be reading (`buffers` = 2 is enough when every step is synchronised: a rank
        cannot pass the barrier of step s + 1 before every rank has enqueued, in stream order, its reads of step s).
      * `finish(sync=False)`: no barrier; the records of this step still reach every peer.  A closed loop that only
        needs the gathered records every k steps (or at episode end: SURVEY.md 8e) synchronises then:
        `finish(sync=True)`, read, `release()` -- the second barrier keeps fast ranks from overwriting the table
        while a peer is still reading it, since without per-step barriers ranks may be several steps apart.
    """

    def __init__(self, engine, rows: int, row_offset: int, group=None, buffers: int = 2):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        from . import _cabi
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if self.world > 8:
            raise ValueError("fused record gather addresses the GPUs of one NVSwitch box (<= 8)")
        dev = torch.device("cuda", engine.device)
        self.rows, self.row_offset, self.buffers = int(rows), int(row_offset), int(buffers)
        self.tables = symm_mem.empty(self.buffers, self.rows, _cabi.RECORD_LEN, dtype=torch.float64, device=dev)
        self.tables.zero_()
        self.handle = symm_mem.rendezvous(self.tables, self.group)
        self.ptrs = [int(p) for p in self.handle.buffer_ptrs]
        if len(self.ptrs) != self.world or any(p == 0 for p in self.ptrs):
            raise RuntimeError("symmetric-memory rendezvous returned no peer pointers")
        self.table_bytes = self.rows * _cabi.RECORD_LEN * 8
        self.engine = engine
        self.step_no = 0
        self.begin_step()

    @property
    def table(self):
        """The table the current (or, after finish(), the last finished) step writes."""
        return self.tables[self._cur]

    def begin_step(self, row_offset: Optional[int] = None):
        """Point the engine's epilogue at this step's table (a host-side call, no CUDA work).  `row_offset` overrides
        the row this rank's instance 0 goes to (several launches per step, each with its own slice of the table)."""
        self._cur = self.step_no % self.buffers
        off = self._cur * self.table_bytes
        self.engine.set_record_peers([p + off for p in self.ptrs], self.row_offset if row_offset is None else row_offset)

    def finish(self, sync: bool = True):
        """Call after the step's launches (same stream).  Returns the step's table: complete once the stream has
        passed the barrier (sync=True), otherwise only this rank's rows and whatever the peers have stored so far."""
        if sync:
            self.handle.barrier(channel=0)
        tab = self.tables[self._cur]
        self.step_no += 1
        self.begin_step()
        return tab

    def release(self):
        """Second barrier of a deferred synchronisation: every rank is done reading."""
        self.handle.barrier(channel=0)

    def close(self):
        self.engine.set_record_peers([], 0)


class ShardedMPC:
    """Runs this rank's slice of a global batch and gathers the result records.

    engine: a BatchedMPC on this rank's GPU.  Inputs to `step` are this rank's slices (device tensors).  With
    `fused=True` (NCCL group on one NVSwitch box) the gather is the step kernel's own epilogue plus one barrier;
    otherwise it is one `all_gather_into_tensor` of the record buffer."""

    def __init__(self, engine, B_global: int, group=None, fused: bool = False):
        import torch.distributed as dist
        self.engine, self.B, self.group = engine, int(B_global), group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.lo, self.hi = shard_bounds(self.B, self.world, self.rank)
        self.fused = FusedRecordGather(engine, self.B, self.lo, group) if (fused and self.world > 1) else None

    def step(self, state, target_ind, oa, od, out, sync: bool = True, **kw):
        out = self.engine.step(state, target_ind, oa, od, out, **kw)
        if self.fused is not None:
            return out, self.fused.finish(sync=sync)
        if self.world > 1:
            return out, allgather_records(out.record, self.B, self.group)
        return out, out.record
