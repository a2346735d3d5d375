"""GPU, 2 ranks over NCCL: ONE read set compressed by two GPUs (shared claim bitmap over NVLink peer memory, pool claims by
all-reduce(min)).  Skipped on a box with fewer than two GPUs.  The result must decode losslessly with the reference's
decoder and stay within 2 % of the single-GPU archive."""
import os
import socket
import sys

import numpy as np
import pytest

import harness as H
import refrun as R

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, src, dst, L, shard_dicts):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import harc_b200
    from harc_b200 import multi
    out = os.path.join(src, "output")
    clean = np.fromfile(os.path.join(out, "input_clean.dna"), dtype=np.uint8)
    withN = np.fromfile(os.path.join(out, "input_N.dna"), dtype=np.uint8)
    ctx = harc_b200.HarcGpu(L, device=rank, file_sets=1, shard_dicts=shard_dicts)
    res = multi.compress_sharded(ctx, dist, clean, withN)
    # a second pass on the connected context must work too (bench loop)
    ctx.load_reads(clean)  # as the bench loop does: reload, rebuild (the shard tables are rebuilt in place), rerun
    ctx.build_dicts()
    res = multi.fetch(ctx, multi.run_pass(ctx, dist, withN, rank, world, torch))
    multi.write_outputs(dst, rank, world, res, L, dist)
    m, s, u = res["counts"]
    stats = [None] * world
    dist.all_gather_object(stats, (m, s, int(res["sizes"].aligned_singletons), int(res["sizes"].aligned_N)))
    if rank == 0:
        np.save(os.path.join(dst, "stats.npy"), np.array(stats, dtype=np.int64))
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("shard_dicts", [0, 1], ids=["dicts_replicated", "dicts_sharded"])
@pytest.mark.parametrize("case", [("mg100", 120000, 100, 600000, True, True)], ids=["L100_rc_err"])
def test_one_job_on_two_gpus(workroot, case, shard_dicts):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    name, n, L, G, rc, err = case
    d = H.make_dataset(workroot, name, n, L, G, rc, err, seed=17)
    dst = d + ".two%d" % shard_dicts
    os.makedirs(os.path.join(dst, "output"), exist_ok=True)
    mp.spawn(_worker, args=(2, _free_port(), d, dst, L, shard_dicts), nprocs=2, join=True)
    stats = np.load(os.path.join(dst, "stats.npy"))
    n_clean = os.path.getsize(os.path.join(d, "output", "input_clean.dna")) // (L + 1)
    n_N = os.path.getsize(os.path.join(d, "output", "input_N.dna")) // (L + 1)
    assert stats[:, 0].sum() + stats[:, 1].sum() == n_clean          # the chains of both GPUs partition the read set
    assert (stats[:, 0] > 0).all()                                   # both GPUs walked
    order = np.fromfile(os.path.join(dst, "output", "read_order.bin"), dtype=np.uint32)
    assert np.array_equal(np.sort(order), np.arange(n_clean, dtype=np.uint32))
    order_N = np.fromfile(os.path.join(dst, "output", "read_order_N_pe.bin"), dtype=np.uint32)
    assert np.array_equal(np.sort(order_N), np.arange(n_N, dtype=np.uint32))
    # single-GPU archive of the same input for the size comparison
    one = H.clone(d, d + ".one%d" % shard_dicts)
    import harc_b200
    ctx = harc_b200.HarcGpu(L, file_sets=1)
    ctx.reorder_dir(one)
    ctx.encode_dir(one)
    ctx.close()
    s1, s2 = R.standin_size(one)[0], R.standin_size(dst)[0]
    print("two GPUs / one GPU archive size: %.4f" % (s2 / s1))
    assert s2 <= 1.02 * s1
    # lossless through the reference decoder (two file sets)
    R.decoder(dst)
    fq = np.fromfile(os.path.join(d, "r.fastq"), dtype=np.uint8).tobytes().split(b"\n")[1::4]
    want = os.path.join(dst, "all.dna")
    with open(want, "wb") as f:
        f.write(b"\n".join(fq) + b"\n")
    assert R.sorted_lines_digest(os.path.join(dst, "output", "output.dna"), L) == R.sorted_lines_digest(want, L)
