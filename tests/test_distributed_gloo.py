"""world_size-2 gloo tests (CPU) of the N>1 host logic: contiguous profile shards, uneven blocks,
one all-gather of the stacked outputs (distributed.py).  The compute itself needs a GPU and is
replaced here by a deterministic stand-in so that only the plumbing is tested."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mwr_fast_forward_operators_and_lbls_b200.distributed import (gather_outputs, gather_profiles_axis,
                                                                  shard_range, shard_sizes)


def test_shard_ranges_cover_everything_once():
    for P in (1, 7, 10, 10000, 100003):
        for world in (1, 2, 3, 8):
            blocks = [shard_range(P, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == P
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = shard_sizes(P, world)
            assert sum(sizes) == P and max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, P, ok):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        F, A = 3, 2
        full = torch.arange(P * F * A, dtype=torch.float64).reshape(P, F, A)
        a, b = shard_range(P, rank, world)
        local = {"tbtotal": full[a:b].clone(), "tauwet": -full[a:b].clone(),
                 "status": torch.zeros(b - a, dtype=torch.int32)}
        got = gather_outputs(local, P, fields=("tbtotal", "tauwet"))
        assert torch.equal(got["tbtotal"], full) and torch.equal(got["tauwet"], -full)
        g2 = gather_profiles_axis(full[a:b].clone(), P)
        assert torch.equal(g2, full)
        with pytest.raises(ValueError):
            gather_profiles_axis(full[: (b - a) + 1].clone(), P)
        ok[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("P", [10, 11])
def test_gather_world2_gloo(P):
    world = 2
    ctx = mp.get_context("spawn")
    ok = ctx.Array("i", [0] * world)
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, P, ok)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert list(ok) == [1] * world
