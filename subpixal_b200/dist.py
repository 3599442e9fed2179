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

__all__ = ["shard_range", "pack_results", "allgather_results", "cc_fit_frames_sharded",
           "cc_fit_stack_sharded", "frame_displacements_sharded"]


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
    sh, pk, st, _ = batch.cc_fit_frames(frames, boxes[lo:hi], ny, nx, cc_type, **kw)
    return allgather_results(pack_results(sh, pk, st, per), n, group)


def cc_fit_stack_sharded(local_stack, n_total, cc_type="NCC", group=None):
    """``local_stack`` is this rank's ``[m, 5, ny, nx]`` shard (m = its part of
    shard_range(n_total)); returns the full result table on every rank."""
    from . import batch
    rank, world = _rank_world(group)
    _, _, per = shard_range(n_total, rank, world)
    sh, pk, st, _ = batch.cc_fit_stack(local_stack, cc_type)
    return allgather_results(pack_results(sh, pk, st, per), n_total, group)


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
