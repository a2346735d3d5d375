"""CPU, world_size 2 over gloo: the host-side logic of one job on several ranks (harc_b200/multi.py): slice split,
whole-job aggregation (reads of all ranks / slowest rank), assembly of the global order streams."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from harc_b200 import multi


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _spawn(fn, world=2):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=fn, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def _init(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    return multi.DistComm(dist)


def _agg_worker(rank, world, port, q):
    comm = _init(rank, world, port)
    b, e = multi.slice_ranges(1001, world)[rank]
    ms = 100.0 if rank == 0 else 250.0  # rank 1 is the slow one
    val, t, n = multi.whole_job_throughput(e - b, ms, comm)
    counts = comm.all_gather_object(e - b)  # what multi.Job does to find its base
    q.put((rank, b, e, val, t, n, sum(counts[:rank])))
    comm.barrier()
    dist.destroy_process_group()


def test_two_ranks_aggregate_as_sum_of_reads_over_max_time():
    res = _spawn(_agg_worker)
    assert [(r[1], r[2]) for r in res] == [(0, 501), (501, 1001)]
    for r in res:
        assert r[4] == 250.0 and r[5] == 1001.0
        assert r[3] == pytest.approx(1001 / 0.25 / 1e6)
        assert r[6] == r[1]   # base of the slice = reads of the ranks before


def test_slice_ranges_cover_everything():
    for n in (0, 1, 7, 35_000_000):
        for w in (1, 2, 4, 8):
            rs = multi.slice_ranges(n, w)
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            assert max(e - b for b, e in rs) - min(e - b for b, e in rs) <= 1


def _gather_worker(rank, world, port, q):
    comm = _init(rank, world, port)
    L = 8
    if rank == 0:  # 3 aligned + 2 unaligned singletons (2*8 bases = 4 packed bytes), 1 aligned N + 1 unaligned N
        part = dict(order=np.array([10, 11, 12, 90, 91], np.uint32), order_N=np.array([5, 7], np.uint32),
                    singleton=np.zeros(4, np.uint8), singleton_tail=np.zeros(0, np.uint8), input_N=np.zeros(9, np.uint8))
    else:
        part = dict(order=np.array([20, 21], np.uint32), order_N=np.array([6], np.uint32),
                    singleton=np.zeros(0, np.uint8), singleton_tail=np.zeros(0, np.uint8), input_N=np.zeros(0, np.uint8))
    g = multi.assemble_globals(comm.all_gather_object(part), L)
    q.put((rank, g["order"].tolist(), g["order_N"].tolist()))
    comm.barrier()
    dist.destroy_process_group()


def test_two_ranks_assemble_the_global_order_streams():
    """One job on two ranks: file sets in rank order, then rank 0's unaligned tail (decoder.cpp:141-169)."""
    for r in _spawn(_gather_worker):
        assert r[1] == [10, 11, 12, 20, 21, 90, 91]
        assert r[2] == [5, 6, 7]
