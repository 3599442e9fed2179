"""Multi-GPU sharding of a batch of cutout pairs (one process per GPU).

Pairs are independent (the loop body of reference align.py:656-699 touches only
its own cutouts), so the source table is split into contiguous, equally sized
ranges - rank r gets ``[r*ceil(n/G), (r+1)*ceil(n/G))``, the tail padded - and
the only communication is ONE all-gather of the per-rank result table
(``[m, 4]`` float64: dx, dy, peak index, status), after which every rank holds
the full table for the linear fit.  The same code runs over NCCL (GPU) and
gloo (CPU tests of the host logic).
"""
import torch
import torch.distributed as dist

__all__ = ["shard_range", "pack_results", "allgather_results", "ResultGather",
           "cc_fit_frames_sharded", "cc_fit_stack_sharded", "frame_displacements_sharded"]


def shard_range(n, rank, world):
    """(start, stop, per_rank): rank's slice of range(n) and the padded shard
    length ``per_rank = ceil(n / world)``."""
    per = (n + world - 1) // world if world > 0 else n
    start = min(n, rank * per)
    return start, min(n, start + per), per


def pack_results(shifts, peak, status, per_rank):
    """[m,2] f64 + [m] i32 + [m] i32 -> [per_rank, 4] f64 table, NaN/-1 padded."""
    m = shifts.shape[0]
    tab = torch.full((per_rank, 4), float("nan"), dtype=torch.float64, device=shifts.device)
    tab[:, 2:] = -1.0
    tab[:m, :2] = shifts
    tab[:m, 2] = peak.to(torch.float64)
    tab[:m, 3] = status.to(torch.float64)
    return tab


def allgather_results(table, n, group=None):
    """One all_gather_into_tensor of the equally sized per-rank tables; returns
    ``(shifts [n,2] f64, peak [n] i64, status [n] i64)`` on every rank."""
    if dist.is_available() and dist.is_initialized():
        world = dist.get_world_size(group)
        full = torch.empty((world * table.shape[0], 4), dtype=table.dtype,
                           device=table.device)
        dist.all_gather_into_tensor(full, table.contiguous(), group=group)
    else:
        full = table
    full = full[:n]
    return full[:, :2].contiguous(), full[:, 2].to(torch.int64), full[:, 3].to(torch.int64)


class ResultGather(object):
    """Persistent buffers for the ONE collective of the path: the kernel writes this
    rank's packed rows (dx, dy, peak, status; float64) straight into ``self.local``
    (``batch.cc_fit_*(..., table=gather.local)``), ``gather()`` sends them with one
    ``all_gather_into_tensor`` into ``self.full`` and returns VIEWS of it - no packing
    or unpacking kernels, no allocation per step.  Rows past a rank's own pair count
    keep the NaN / -1 padding they were created with.

    ``depth`` > 1 keeps a ring of buffer sets for pipelined batches: ``advance()`` moves to
    the next set (waiting for the all-gather that last used it), ``gather(async_op=True)``
    only enqueues the collective on NCCL's stream, so the all-gather of batch i overlaps
    the kernel of batch i + 1; ``wait()`` joins everything outstanding."""

    def __init__(self, per_rank, world, device, group=None, depth=1):
        self.per, self.world, self.group = int(per_rank), int(world), group
        self._sets, self._works, self._cur = [], [], 0
        for _ in range(max(1, int(depth))):
            local = torch.empty((self.per, 4), dtype=torch.float64, device=device)
            local[:, :2] = float("nan")
            local[:, 2:] = -1.0
            full = (torch.empty((self.world * self.per, 4), dtype=torch.float64, device=device)
                    if self.world > 1 else local)
            self._sets.append((local, full))
            self._works.append(None)

    @property
    def local(self):
        return self._sets[self._cur][0]

    @property
    def full(self):
        return self._sets[self._cur][1]

    def advance(self):
        """Switch to the next buffer set of the ring; returns its send buffer."""
        self._cur = (self._cur + 1) % len(self._sets)
        w = self._works[self._cur]
        if w is not None:
            w.wait()
            self._works[self._cur] = None
        return self.local

    def wait(self):
        for k, w in enumerate(self._works):
            if w is not None:
                w.wait()
                self._works[k] = None

    def gather(self, n=None, async_op=False):
        """All-gather ``self.local``; returns float64 views ``(shifts [n,2], peak [n],
        status [n])`` of the gathered table (peak / status hold exact integers).  With
        ``async_op`` the views are valid after ``wait()`` (or ``advance()`` back to this set)."""
        if self.world > 1:
            w = dist.all_gather_into_tensor(self.full, self.local, group=self.group,
                                            async_op=async_op)
            self._works[self._cur] = w if async_op else None
        full = self.full
        n = full.shape[0] if n is None else n
        return full[:n, :2], full[:n, 2], full[:n, 3]

    def __call__(self, shifts, peak, status):
        """Results that were NOT written through ``table=self.local`` (separate tensors):
        copy them into the send buffer, then gather."""
        m = shifts.shape[0]
        if shifts.data_ptr() != self.local.data_ptr():
            self.local[:m, :2] = shifts
            self.local[:m, 2] = peak
            self.local[:m, 3] = status
        return self.gather()


def _rank_world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def cc_fit_frames_sharded(frames, boxes, ny, nx, cc_type="NCC", group=None, **kw):
    """Frames are replicated on every rank (<= 67 MB each); the box table is
    sharded by source.  Returns the full result table on every rank."""
    from . import batch
    rank, world = _rank_world(group)
    n = int(boxes.shape[0])
    lo, hi, per = shard_range(n, rank, world)
    g = ResultGather(per, world, boxes.device if boxes.is_cuda else frames[0].device, group)
    batch.cc_fit_frames(frames, boxes[lo:hi], ny, nx, cc_type, table=g.local, **kw)
    sh, pk, st = g.gather(n)
    return sh.contiguous(), pk.to(torch.int64), st.to(torch.int64)


def cc_fit_stack_sharded(local_stack, n_total, cc_type="NCC", group=None):
    """``local_stack`` is this rank's ``[m, 5, ny, nx]`` shard (m = its part of
    shard_range(n_total)); returns the full result table on every rank."""
    from . import batch
    rank, world = _rank_world(group)
    _, _, per = shard_range(n_total, rank, world)
    g = ResultGather(per, world, local_stack.device, group)
    batch.cc_fit_stack(local_stack, cc_type, table=g.local)
    sh, pk, st = g.gather(n_total)
    return sh.contiguous(), pk.to(torch.int64), st.to(torch.int64)


def frame_displacements_sharded(img_frame, drz_frame, img_boxes, drz_boxes, frame_map, group=None,
                                maps=None, flux_scale=1.0, src_ids=None, **kw):
    """``align.frame_displacements`` (gather + gather + blot + CC, all on the device)
    with the source table sharded by rank; both frames are replicated.  Returns the
    full ``(shifts, peak, status)`` table on every rank after ONE all-gather.  Keyword
    arguments (``cc_type``, ``drz_seg``, ``drz_dq``, ``drz_dqbitmask``, ``interp`` ...)
    are forwarded."""
    import numpy as np
    from . import align
    rank, world = _rank_world(group)
    ib = np.asarray(img_boxes).reshape(-1, 4)
    db = np.asarray(drz_boxes).reshape(-1, 4)
    n = ib.shape[0]
    lo, hi, per = shard_range(n, rank, world)
    part = lambda v: None if v is None else np.asarray(v)[lo:hi]
    fs = flux_scale if np.ndim(flux_scale) == 0 else np.asarray(flux_scale)[lo:hi]
    sh, pk, st = align.frame_displacements(img_frame, drz_frame, ib[lo:hi], db[lo:hi], frame_map,
                                           flux_scale=fs, src_ids=part(src_ids), maps=part(maps), **kw)
    return allgather_results(pack_results(sh, pk, st, per), n, group)
