"""world_size-2 gloo test (CPU) of the sharding + single all-gather logic."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, n):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from subpixal_b200 import dist as d
        lo, hi, per = d.shard_range(n, rank, world)
        idx = torch.arange(lo, hi, dtype=torch.float64)
        shifts = torch.stack([idx * 0.5, -idx * 0.25], dim=1)
        peak = torch.arange(lo, hi, dtype=torch.int32) * 3
        status = torch.full((hi - lo,), rank, dtype=torch.int32)
        sh, pk, st = d.allgather_results(d.pack_results(shifts, peak, status, per), n)
        full = torch.arange(n, dtype=torch.float64)
        assert sh.shape == (n, 2) and torch.equal(sh[:, 0], full * 0.5)
        assert torch.equal(sh[:, 1], -full * 0.25)
        assert torch.equal(pk, torch.arange(n) * 3)
        owner = torch.clamp(torch.arange(n) // per, max=world - 1)
        assert torch.equal(st, owner)
    finally:
        dist.destroy_process_group()


def test_two_rank_allgather_of_shift_table():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    for n in (11, 10):
        mp.spawn(_worker, args=(2, port, n), nprocs=2, join=True)
