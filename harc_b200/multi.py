"""One compression job on several GPUs of one box: the host side (one process per GPU, torch.distributed / NCCL for the
plumbing).  See include/harcgpu.h "one job on several GPUs" for the device side.

What crosses GPUs, and how:
  * the claimed-read bitmap of stage I -- peer memory over NVLink (CUDA IPC handles exchanged here), read and claimed
    from inside the walk kernel;
  * the singleton ids of all ranks -- all-gather (they form the common pool of stage II);
  * the pool priorities -- all-reduce(min) over an int64 device array, inside harcgpu_encode through a hook;
  * the order streams -- gathered to rank 0 only when files are written.
Everything else (packed reads, dictionaries) is replicated; rank r's chains become file set r.
"""
import os

import numpy as np


def _dev_tensor(ptr, count, torch):
    """int64 view of `count` words of device memory owned by libharcgpu."""
    class _Arr:
        __cuda_array_interface__ = {"shape": (int(count),), "typestr": "<i8", "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(_Arr(), device="cuda")


def compress_sharded(ctx, dist, clean_ascii, N_ascii, rank=None, world=None):
    """Stage I + II of one read set on all ranks of `dist` (every rank passes the same inputs).  Returns
    dict(set=file set of this rank, glob=its share of the global streams, sizes, counts)."""
    import torch
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    n = ctx.load_reads(clean_ascii)
    handle = ctx.shard_init(rank, world, n)
    handles = [None] * world
    dist.all_gather_object(handles, handle)
    ctx.shard_connect(handles)
    ctx.build_dicts()
    return fetch(ctx, run_pass(ctx, dist, N_ascii, rank, world, torch))


def gather_ids(mine, dist, world, torch):
    """Concatenation, in rank order, of every rank's uint32 id list (NCCL all-gather of padded device tensors)."""
    cnt = torch.tensor([len(mine)], dtype=torch.int64, device="cuda")
    cnts = torch.empty(world, dtype=torch.int64, device="cuda")
    dist.all_gather_into_tensor(cnts, cnt)
    cnts = cnts.cpu().numpy()
    cap = int(cnts.max()) if world else 0
    if cap == 0:
        return np.empty(0, dtype=np.uint32)
    buf = torch.zeros(cap, dtype=torch.int32, device="cuda")
    buf[: len(mine)] = torch.from_numpy(mine.view(np.int32)).cuda()
    allb = torch.empty(world * cap, dtype=torch.int32, device="cuda")
    dist.all_gather_into_tensor(allb, buf)
    h = allb.cpu().numpy().view(np.uint32).reshape(world, cap)
    return np.concatenate([h[r, : int(cnts[r])] for r in range(world)])


def run_pass(ctx, dist, N_ascii, rank, world, torch, n_N=None):
    """One timed pass on a connected context (reads loaded, dictionaries built)."""
    ctx.shard_reset()
    dist.barrier()                      # every range of the bitmap is armed before any walker claims
    m, s, u = ctx.reorder()
    dist.barrier()                      # nobody re-arms or frees its range while a peer still walks
    pool_ids = gather_ids(ctx.get_singleton_ids(), dist, world, torch)

    def exchange(ptr, count):
        t = _dev_tensor(ptr, count, torch)
        assert t.data_ptr() == ptr, "the exchange must work in place on the library's array"
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        torch.cuda.synchronize()
    ctx.set_pool_exchange(exchange)
    ctx.load_pool_ids(pool_ids, N_ascii, n_N)
    es = ctx.encode()
    return dict(sizes=es, counts=(m, s, u), pool=len(pool_ids))


def fetch(ctx, res, empty=np.empty):
    """Copy this rank's file set and its share of the global streams to the host."""
    res["set"] = ctx.get_set(0, empty)
    res["glob"] = ctx.get_globals(empty)
    return res


def assemble_globals(parts, L):
    """parts[r] = dict(order, order_N, singleton, singleton_tail, input_N) of rank r (harcgpu_get_globals).  The decoder
    reads file sets 0..K-1, then the unaligned singletons, then the unaligned N reads (decoder.cpp:141-169), so the
    order streams are: every rank's aligned part in rank order, then rank 0's unaligned tail."""
    p0 = parts[0]
    u_s = (4 * len(p0["singleton"]) + len(p0["singleton_tail"])) // L
    u_n = len(p0["input_N"]) // (L + 1)
    for p in parts[1:]:
        assert len(p["singleton"]) == 0 and len(p["input_N"]) == 0, "only rank 0 writes unaligned pool reads"
    cut = lambda a, k: (a[: len(a) - k], a[len(a) - k:])
    o0, o_tail = cut(p0["order"], u_s)
    n0, n_tail = cut(p0["order_N"], u_n)
    order = np.concatenate([o0] + [p["order"] for p in parts[1:]] + [o_tail])
    order_N = np.concatenate([n0] + [p["order_N"] for p in parts[1:]] + [n_tail])
    return dict(order=order, order_N=order_N, singleton=p0["singleton"], singleton_tail=p0["singleton_tail"], input_N=p0["input_N"])


def write_outputs(basedir, rank, world, res, L, dist):
    """Write the stage II files of SURVEY Appendix A under <basedir>/output/: file set `rank` by every rank, the global
    streams by rank 0."""
    out = os.path.join(basedir, "output")
    os.makedirs(out, exist_ok=True)
    s = res["set"]
    for key, stem in (("seq", "read_seq.txt"), ("pos", "read_pos.txt"), ("noise", "read_noise.txt"),
                      ("noisepos", "read_noisepos.txt"), ("rev", "read_rev.txt")):
        s[key].tofile(os.path.join(out, "%s.%d" % (stem, rank)))
    s["seq_tail"].tofile(os.path.join(out, "read_seq.txt.%d.tail" % rank))
    s["rev_tail"].tofile(os.path.join(out, "read_rev.txt.%d.tail" % rank))
    parts = [None] * world
    dist.all_gather_object(parts, res["glob"])
    if rank == 0:
        g = assemble_globals(parts, L)
        g["order"].tofile(os.path.join(out, "read_order.bin"))
        g["order_N"].tofile(os.path.join(out, "read_order_N_pe.bin"))
        g["singleton"].tofile(os.path.join(out, "read_singleton.txt"))
        g["singleton_tail"].tofile(os.path.join(out, "read_singleton.txt.tail"))
        g["input_N"].tofile(os.path.join(out, "input_N.dna"))
        with open(os.path.join(out, "read_meta.txt"), "w") as f:
            f.write("%d\n" % L)
    dist.barrier()
