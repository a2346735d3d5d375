"""GPU: ONE read set compressed by several ranks (harc_b200/multi.py, csrc/job.cu): every rank uploads a slice, the packed
reads are replicated and the (key, id) pairs exchanged by the library's own kernels over peer memory, the claim bitmap is
shared, pool claims are settled by all-reduce(min).

* `local` cases: the ranks are contexts of this process that share ONE GPU (one host thread per rank), so the whole
  multi-rank path runs on a one-GPU box;
* `nccl` case: one process per GPU over NCCL (skipped on a box with fewer than two GPUs).
The result must hold every read exactly once, decode losslessly with the reference's decoder (one file set per rank) and
stay within 2 % of the single-GPU archive; the union of the dictionary shards must be the single-GPU dictionary."""
import os
import socket
import threading

import numpy as np
import pytest

import harness as H
import refrun as R

pytestmark = pytest.mark.gpu
CASE = ("mg100", 120000, 100, 600000, True, True)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _slices(src, L, world, rank):
    from harc_b200 import multi
    out = os.path.join(src, "output")
    clean = np.fromfile(os.path.join(out, "input_clean.dna"), dtype=np.uint8).reshape(-1, L + 1)
    withN = np.fromfile(os.path.join(out, "input_N.dna"), dtype=np.uint8).reshape(-1, L + 1)
    a, b = multi.slice_ranges(len(clean), world)[rank]
    c, d = multi.slice_ranges(len(withN), world)[rank]
    return np.ascontiguousarray(clean[a:b]).reshape(-1), np.ascontiguousarray(withN[c:d]).reshape(-1)


def _run_rank(comm, ctx, src, dst, L, torch, dump=None):
    from harc_b200 import multi
    rank, world = comm.rank, comm.world
    clean, withN = _slices(src, L, world, rank)
    job = multi.Job(ctx, comm, len(clean) // (L + 1), torch)
    res = job.run(clean, withN)
    if dump is not None:  # this rank's shards of both dictionaries
        dump[rank] = [ctx.dump_dict(1, l) for l in range(2)]
    res = job.run(clean, withN)  # a second pass on the connected contexts must work too (bench loop)
    res = multi.fetch(ctx, res)
    multi.write_outputs(dst, rank, world, res, L, comm)
    m, s, u = res["counts"]
    stats = comm.all_gather_object((m, s, int(res["sizes"].aligned_singletons), int(res["sizes"].aligned_N)))
    if rank == 0:
        np.save(os.path.join(dst, "stats.npy"), np.array(stats, dtype=np.int64))
    comm.barrier()


def _check(d, dst, L, tag):
    import harc_b200
    stats = np.load(os.path.join(dst, "stats.npy"))
    n_clean = os.path.getsize(os.path.join(d, "output", "input_clean.dna")) // (L + 1)
    n_N = os.path.getsize(os.path.join(d, "output", "input_N.dna")) // (L + 1)
    assert stats[:, 0].sum() + stats[:, 1].sum() == n_clean          # the chains of all ranks partition the read set
    assert (stats[:, 0] > 0).all()                                   # every rank walked
    order = np.fromfile(os.path.join(dst, "output", "read_order.bin"), dtype=np.uint32)
    assert np.array_equal(np.sort(order), np.arange(n_clean, dtype=np.uint32))
    order_N = np.fromfile(os.path.join(dst, "output", "read_order_N_pe.bin"), dtype=np.uint32)
    assert np.array_equal(np.sort(order_N), np.arange(n_N, dtype=np.uint32))
    # single-GPU archive of the same input for the size comparison
    one = H.clone(d, d + ".one" + tag)
    ctx = harc_b200.HarcGpu(L, file_sets=1)
    ctx.reorder_dir(one)
    ctx.encode_dir(one)
    ctx.close()
    s1, s2 = R.standin_size(one)[0], R.standin_size(dst)[0]
    print("%s: several ranks / one GPU archive size: %.4f" % (tag, s2 / s1))
    assert s2 <= 1.02 * s1
    # lossless through the reference decoder (one file set per rank)
    R.decoder(dst)
    fq = np.fromfile(os.path.join(d, "r.fastq"), dtype=np.uint8).tobytes().split(b"\n")[1::4]
    want = os.path.join(dst, "all.dna")
    with open(want, "wb") as f:
        f.write(b"\n".join(fq) + b"\n")
    assert R.sorted_lines_digest(os.path.join(dst, "output", "output.dna"), L) == R.sorted_lines_digest(want, L)


@pytest.mark.parametrize("world,shard_dicts", [(2, 1), (4, 1), (2, 0), (8, 1)], ids=["w2_sharded", "w4_sharded", "w2_replicated", "w8_sharded"])
def test_one_job_local_ranks(workroot, world, shard_dicts):
    import torch
    import harc_b200
    from harc_b200 import multi
    name, n, L, G, rc, err = CASE
    d = H.make_dataset(workroot, name, n, L, G, rc, err, seed=17)
    tag = "local%d_%d" % (world, shard_dicts)
    dst = d + "." + tag
    os.makedirs(os.path.join(dst, "output"), exist_ok=True)
    group = multi.LocalGroup(world)
    ctxs = [harc_b200.HarcGpu(L, file_sets=1, shard_dicts=shard_dicts) for _ in range(world)]
    dump = [None] * world
    errs = []

    def body(rank):
        try:
            torch.cuda.set_device(0)
            _run_rank(multi.LocalComm(group, rank, torch), ctxs[rank], d, dst, L, torch, dump)
        except BaseException as e:  # a rank that dies must not leave the others waiting at a host barrier
            errs.append(e)
            group.bar.abort()

    ts = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for c in ctxs:
        c.close()
    assert not errs, errs
    if shard_dicts:
        # the shards of all ranks together are the dictionary of one GPU (keys, bin sizes, ids inside the bins)
        ref = harc_b200.HarcGpu(L)
        ref.load_reads(np.fromfile(os.path.join(d, "output", "input_clean.dna"), dtype=np.uint8))
        ref.build_dicts()
        for l in range(2):
            rk, rc_, ri = ref.dump_dict(1, l)
            keys = np.concatenate([dump[r][l][0] for r in range(world)])
            cnts = np.concatenate([dump[r][l][1] for r in range(world)])
            ids = np.concatenate([dump[r][l][2][: int(dump[r][l][1].sum())] for r in range(world)])
            o = np.argsort(keys, kind="stable")
            assert np.array_equal(keys[o], rk) and np.array_equal(cnts[o], rc_)
            cn = cnts.astype(np.int64)
            starts = np.concatenate([np.zeros(1, np.int64), np.cumsum(cn)])[:-1]
            got = np.concatenate([ids[starts[i]: starts[i] + cn[i]] for i in o]) if len(o) else ids
            assert np.array_equal(got, ri[: int(rc_.sum())])
        ref.close()
    _check(d, dst, L, tag)


def _worker(rank, world, port, src, dst, L, shard_dicts):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import harc_b200
    from harc_b200 import multi
    ctx = harc_b200.HarcGpu(L, device=rank, file_sets=1, shard_dicts=shard_dicts)
    _run_rank(multi.DistComm(dist, torch), ctx, src, dst, L, torch)
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("shard_dicts", [1, 0], ids=["dicts_sharded", "dicts_replicated"])
def test_one_job_on_two_gpus(workroot, shard_dicts):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    name, n, L, G, rc, err = CASE
    d = H.make_dataset(workroot, name, n, L, G, rc, err, seed=17)
    dst = d + ".two%d" % shard_dicts
    os.makedirs(os.path.join(dst, "output"), exist_ok=True)
    mp.spawn(_worker, args=(2, _free_port(), d, dst, L, shard_dicts), nprocs=2, join=True)
    _check(d, dst, L, "nccl%d" % shard_dicts)
