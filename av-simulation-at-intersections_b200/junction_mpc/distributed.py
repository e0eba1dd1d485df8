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

    Every rank allocates a ring of gathered tables [rows, RECORD_LEN] in symmetric memory; after the rendezvous each
    GPU holds NVLink peer pointers to all of them, and `BatchedMPC.set_record_peers` makes the step kernel's epilogue
    store each instance's record straight into row `row_offset + b` of the step's table on every GPU.

    Completion (when is a table complete on this GPU?) has two implementations:

      * sync="flags" (default): the step kernel publishes it itself.  Its last retiring block stores the step number
        into this rank's slot of every peer's flag array (release, system scope, after all record stores), and a reader
        enqueues `wait(step)` -- a one-warp kernel that spins until all ranks' slots have reached `step` -- exactly
        where it needs the table.  Nothing else synchronises the ranks, so a loop that consumes table s after step
        s + 1 was launched (`wait(step - 1)`) overlaps one rank's tail and launch jitter with the next step of the
        others.  The ring has four tables: rank A overwrites the table of step s in its step s + 4, which it launches
        after `wait(s + 2)`, i.e. after every rank has finished step s + 2 and therefore (stream order) its reads of
        table s, which lie between its `wait(s)` and its launch of step s + 2.
      * sync="barrier": `finish()` runs one cross-GPU barrier (symmetric-memory signal pads) after every step; two
        tables are enough (a rank cannot pass the barrier of step s + 1 before every rank has enqueued, in stream
        order, its reads of step s).
    """

    def __init__(self, engine, rows: int, row_offset: int, group=None, sync: str = "flags", buffers: Optional[int] = None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        from . import _cabi
        if sync not in ("flags", "barrier"):
            raise ValueError("sync must be 'flags' or 'barrier'")
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if self.world > 8:
            raise ValueError("fused record gather addresses the GPUs of one NVSwitch box (<= 8)")
        dev = torch.device("cuda", engine.device)
        self.sync = sync
        self.rows, self.row_offset = int(rows), int(row_offset)
        self.buffers = int(buffers or (4 if sync == "flags" else 2))
        if sync == "flags" and self.buffers < 4:
            raise ValueError("the flag protocol needs a ring of at least four tables")
        self.tables = symm_mem.empty(self.buffers, self.rows, _cabi.RECORD_LEN, dtype=torch.float64, device=dev)
        self.tables.zero_()
        self.handle = symm_mem.rendezvous(self.tables, self.group)
        self.ptrs = [int(p) for p in self.handle.buffer_ptrs]
        if len(self.ptrs) != self.world or any(p == 0 for p in self.ptrs):
            raise RuntimeError("symmetric-memory rendezvous returned no peer pointers")
        self.table_bytes = self.rows * _cabi.RECORD_LEN * 8
        self.flags = symm_mem.empty(max(self.world, 8), dtype=torch.int64, device=dev)     # slot r: last step rank r published
        self.flags.zero_()
        self.flag_handle = symm_mem.rendezvous(self.flags, self.group)
        self.flag_ptrs = [int(p) + 8 * self.rank for p in self.flag_handle.buffer_ptrs]   # my slot in every rank's array
        torch.cuda.synchronize(dev)
        dist.barrier(self.group)                  # every rank's tables and flags are zeroed before anyone stores into them
        self.engine = engine
        self.step_no = 0                          # steps finished so far; step numbers start at 1
        self.begin_step()

    @property
    def table(self):
        """The table the current (or, after finish(), the last finished) step writes."""
        return self.tables[self._cur]

    def table_of(self, step: int):
        return self.tables[(step - 1) % self.buffers]

    def begin_step(self, row_offset: Optional[int] = None, publish: bool = True):
        """Point the engine's epilogue at this step's table (host-side calls, no CUDA work).  `row_offset` overrides
        the row this rank's instance 0 goes to (several launches per step, each with its own slice of the table);
        `publish` = this launch is the step's last one and announces the step as complete (flag protocol)."""
        self._cur = self.step_no % self.buffers
        off = self._cur * self.table_bytes
        self.engine.set_record_peers([p + off for p in self.ptrs], self.row_offset if row_offset is None else row_offset)
        if self.sync == "flags":
            self.engine.set_record_flags(self.flag_ptrs if publish else [], self.step_no + 1)

    def finish(self):
        """Call after the step's launches (same stream).  sync="barrier": runs the cross-GPU barrier and returns the
        complete table.  sync="flags": returns the table without waiting -- call `wait(step)` before reading it."""
        if self.sync == "barrier":
            self.handle.barrier(channel=0)
        tab = self.tables[self._cur]
        self.step_no += 1
        self.begin_step()
        return tab

    def wait(self, step: Optional[int] = None):
        """Flag protocol: enqueue the wait for all ranks' records of `step` (default: the last finished step) on the
        current stream and return that step's table."""
        step = self.step_no if step is None else int(step)
        if step < 1:
            return None
        if self.sync == "flags":
            self.engine.gather_wait(self.flags, self.world, step)
        return self.table_of(step)

    def close(self):
        self.engine.set_record_peers([], 0)
        self.engine.set_record_flags([], 0)


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

    def step(self, state, target_ind, oa, od, out, **kw):
        out = self.engine.step(state, target_ind, oa, od, out, **kw)
        if self.fused is not None:
            self.fused.finish()
            return out, self.fused.wait()           # complete table of this step (stream ordered)
        if self.world > 1:
            return out, allgather_records(out.record, self.B, self.group)
        return out, out.record
