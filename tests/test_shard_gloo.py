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


def _gather_worker(rank, world, port, q):
    import numpy as np
    from harc_b200 import multi
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L = 8
    if rank == 0:  # 3 aligned + 2 unaligned singletons (2*8 bases = 4 packed bytes), 1 aligned N + 1 unaligned N
        part = dict(order=np.array([10, 11, 12, 90, 91], np.uint32), order_N=np.array([5, 7], np.uint32),
                    singleton=np.zeros(4, np.uint8), singleton_tail=np.zeros(0, np.uint8), input_N=np.zeros(9, np.uint8))
    else:
        part = dict(order=np.array([20, 21], np.uint32), order_N=np.array([6], np.uint32),
                    singleton=np.zeros(0, np.uint8), singleton_tail=np.zeros(0, np.uint8), input_N=np.zeros(0, np.uint8))
    parts = [None] * world
    dist.all_gather_object(parts, part)
    g = multi.assemble_globals(parts, L)
    q.put((rank, g["order"].tolist(), g["order_N"].tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_assemble_the_global_order_streams():
    """One job on two ranks: file sets in rank order, then rank 0's unaligned tail (decoder.cpp:141-169)."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_gather_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in res:
        assert r[1] == [10, 11, 12, 20, 21, 90, 91]
        assert r[2] == [5, 6, 7]
