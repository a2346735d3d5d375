"""CPU, world_size 2 over gloo: the host-side multi-rank logic (read-set sharding and whole-job aggregation)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from harc_b200.shard import aggregate_throughput, split_fastq_ranges


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = split_fastq_ranges(1001, world)[rank]
    ms = 100.0 if rank == 0 else 250.0  # rank 1 is the slow one
    val, t, n = aggregate_throughput(e - b, ms, dist)
    q.put((rank, b, e, val, t, n))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_aggregate_as_sum_of_reads_over_max_time():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 501), (501, 1001)]
    for r in res:
        assert r[4] == 250.0 and r[5] == 1001.0
        assert r[3] == pytest.approx(1001 / 0.25 / 1e6)


def test_split_ranges_cover_everything():
    for n in (0, 1, 7, 35_000_000):
        for w in (1, 2, 4, 8):
            rs = split_fastq_ranges(n, w)
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            assert max(e - b for b, e in rs) - min(e - b for b, e in rs) <= 1
