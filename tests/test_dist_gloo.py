"""
World-size-2 test of the N>1 host logic on CPU (gloo): shard bounds and the final gather.  The kernels
themselves never communicate, so this is all the multi-rank logic there is.
"""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from galois_b200._dist import gather_rows, shard_bounds, shard_sizes


def test_shard_bounds_cover_everything():
    for total in (0, 1, 7, 4096, 10**6 + 3):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = shard_sizes(total, world)
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == total
    assert shard_bounds(4096, 3, 8) == (1536, 2048)  # BASELINE config 3: 512 transforms per GPU at 8 GPUs
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(total * 5, dtype=torch.int64).reshape(total, 5) * 3 + 1
        lo, hi = shard_bounds(total, rank, world)
        local = full[lo:hi].clone()
        got = gather_rows(local, total)
        q.put((rank, bool(torch.equal(got, full)), tuple(got.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7])
def test_gather_rows_world2_gloo(total):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in results) == [0, 1]
    assert all(ok and shape == (total, 5) for _, ok, shape in results)
