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
        # persistent-buffer form: the producer writes packed rows into g.local (what the
        # kernel does with table=g.local), one all-gather, views come back
        g = d.ResultGather(per, world, torch.device("cpu"))
        assert torch.isnan(g.local[:, :2]).all() and bool((g.local[:, 2:] == -1).all())
        g.local[:hi - lo, :2] = shifts
        g.local[:hi - lo, 2] = peak
        g.local[:hi - lo, 3] = status
        for _ in range(2):                                   # buffers are reused
            sh2, pk2, st2 = g.gather(n)
            assert torch.equal(sh2, sh) and torch.equal(pk2.long(), pk) and torch.equal(st2.long(), st)
            assert sh2.data_ptr() == g.full.data_ptr()       # a view, nothing unpacked
        if n % world:                                         # the tail keeps its padding
            assert torch.isnan(g.full[n:, :2]).all()
        sh3, _, _ = g(shifts, peak, status)                   # separate tensors -> copied in
        assert torch.equal(sh3[:n], sh)
        # pipelined form: ring of buffer sets, asynchronous all-gathers
        ring = d.ResultGather(per, world, torch.device("cpu"), depth=2)
        outs = []
        for step in range(5):
            loc = ring.advance()
            loc[:hi - lo, :2] = shifts + step
            loc[:hi - lo, 2] = peak
            loc[:hi - lo, 3] = status
            outs.append((step, ring.gather(n, async_op=True)[0]))
            if step >= 1:                                     # the previous step's table is still intact
                ring._works[(ring._cur + 1) % 2].wait()
                assert torch.equal(outs[step - 1][1], sh + (step - 1))
        ring.wait()
        assert torch.equal(outs[-1][1], sh + 4)
    finally:
        dist.destroy_process_group()


def test_two_rank_allgather_of_shift_table():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    for n in (11, 10):
        mp.spawn(_worker, args=(2, port, n), nprocs=2, join=True)
