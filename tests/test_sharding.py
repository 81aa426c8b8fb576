"""Window sharding + the single score all-gather, on CPU with gloo and world_size 2."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mocodad_b200.sharding import gather_scores, shard_bounds


@pytest.mark.parametrize("n,world", [(0, 2), (1, 2), (7, 2), (1024, 8), (1000003, 8), (5, 8)])
def test_shard_bounds_partition_the_range(n, world):
    spans = [shard_bounds(n, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [hi - lo for lo, hi in spans]
    assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(n, world, world)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_bounds(n_total, rank, world)
        local = torch.arange(lo, hi, dtype=torch.float32) * 0.5  # "score" of global window i is i/2
        full = gather_scores(local, n_total)
        q.put((rank, full.tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [7, 10, 1])
def test_gather_scores_world_size_2(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [0.5 * i for i in range(n_total)]
    assert got[0] == want and got[1] == want


def test_single_process_passthrough():
    x = torch.arange(5.0)
    assert gather_scores(x, 5) is x
