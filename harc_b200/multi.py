"""One compression job on several GPUs of one box: the host side (one process per GPU, torch.distributed / NCCL for the
plumbing).  See include/harcgpu.h "one job on several GPUs" and csrc/job.cu for the device side.

What crosses GPUs, and how:
  * packed reads, (key, id) pairs of the dictionary build, Bloom filter segments, dictionary probes, claims, barriers
    -- kernels of libharcgpu over NVLink peer memory (the arenas whose CUDA IPC handles are exchanged here);
  * the singleton ids of all ranks and the reads with N of all slices -- all-gather of device tensors (the common pool of
    stage II);
  * the pool priorities -- all-reduce(min) over an int64 device array, inside harcgpu_encode through a hook;
  * the order streams -- gathered to rank 0 only when files are written.
Rank r uploads slice r of the clean reads and its chains become file set r (encoder.cpp:169-196).

`DistComm` is the real thing (one process per GPU).  `LocalComm` lets `world` contexts of ONE process share one GPU, one
host thread per rank, so that the whole multi-rank path (exchange kernels, barriers, sharded probes) is also exercised
by the GPU tests of a one-GPU box.
"""
import os
import threading

import numpy as np


def _dev_tensor(ptr, count, torch, typestr="<i8"):
    """View of `count` elements of device memory owned by libharcgpu."""
    if count == 0:
        return torch.empty(0, dtype={"<i8": torch.int64, "<i4": torch.int32, "|u1": torch.uint8}[typestr], device="cuda")

    class _Arr:
        __cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr, "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(_Arr(), device="cuda")


def slice_ranges(n_reads, world):
    """Contiguous [begin, end) read ranges of one input, one per rank, sizes differing by at most one (the reference's
    static split of reorder.cpp:242-261)."""
    base, extra = divmod(int(n_reads), int(world))
    out, b = [], 0
    for r in range(world):
        e = b + base + (1 if r < extra else 0)
        out.append((b, e))
        b = e
    return out


def whole_job_throughput(local_reads, local_ms, comm=None):
    """Whole-job Mreads/s = reads of all ranks / slowest rank's time (never a sum of per-rank rates)."""
    if comm is None:
        return local_reads / (local_ms / 1000.0) / 1e6, float(local_ms), float(local_reads)
    parts = comm.all_gather_object((float(local_reads), float(local_ms)))
    n = sum(p[0] for p in parts)
    t = max(p[1] for p in parts)
    return n / (t / 1000.0) / 1e6, t, n


class DistComm:
    """torch.distributed process group (NCCL on the GPUs; gloo works for the host-only calls)."""
    local = False

    def __init__(self, dist, torch=None):
        self.dist, self.torch = dist, torch
        self.rank, self.world = dist.get_rank(), dist.get_world_size()

    def all_gather_object(self, obj):
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def barrier(self):
        self.dist.barrier()

    def _all_gather_var(self, t):
        """Concatenation in rank order of every rank's 1-D device tensor (lengths differ): only the lengths visit the host."""
        torch = self.torch
        cnt = torch.tensor([t.numel()], dtype=torch.int64, device="cuda")
        cnts = torch.empty(self.world, dtype=torch.int64, device="cuda")
        self.dist.all_gather_into_tensor(cnts, cnt)
        cnts = cnts.tolist()
        cap = max(cnts)
        if cap == 0:
            return t[:0].clone()
        buf = torch.zeros(cap, dtype=t.dtype, device="cuda")
        buf[: t.numel()] = t
        allb = torch.empty(self.world * cap, dtype=t.dtype, device="cuda")
        self.dist.all_gather_into_tensor(allb, buf)
        return torch.cat([allb[r * cap: r * cap + cnts[r]] for r in range(self.world)])

    def all_gather_u32(self, ptr, count):
        return self._all_gather_var(_dev_tensor(ptr, count, self.torch, "<i4"))

    def all_gather_bytes(self, t):
        return self._all_gather_var(t)

    def all_reduce_min_i64(self, ptr, count):
        t = _dev_tensor(ptr, count, self.torch)
        assert t.data_ptr() == ptr, "the exchange must work in place on the library's array"
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        self.torch.cuda.synchronize()


class LocalGroup:
    def __init__(self, world):
        self.world = world
        self.bar = threading.Barrier(world)
        self.slots = [None] * world


class LocalComm:
    """`world` contexts of this process on one GPU, one host thread per rank."""
    local = True

    def __init__(self, group, rank, torch):
        self.g, self.rank, self.world, self.torch = group, rank, group.world, torch

    def all_gather_object(self, obj):
        self.g.slots[self.rank] = obj
        self.g.bar.wait()
        out = list(self.g.slots)
        self.g.bar.wait()
        return out

    def barrier(self):
        self.torch.cuda.synchronize()
        self.g.bar.wait()

    def _all_gather_var(self, t):
        self.torch.cuda.synchronize()
        parts = self.all_gather_object(t)
        out = self.torch.cat(parts) if parts else t
        self.torch.cuda.synchronize()
        self.g.bar.wait()  # nobody's tensor goes away before everybody has copied it
        return out

    def all_gather_u32(self, ptr, count):
        return self._all_gather_var(_dev_tensor(ptr, count, self.torch, "<i4"))

    def all_gather_bytes(self, t):
        return self._all_gather_var(t)

    def all_reduce_min_i64(self, ptr, count):
        torch = self.torch
        t = _dev_tensor(ptr, count, torch)
        parts = self.all_gather_object(t)
        if self.rank == 0:
            m = parts[0].clone()
            for p in parts[1:]:
                torch.minimum(m, p, out=m)
            for p in parts:
                p.copy_(m)
            torch.cuda.synchronize()
        self.g.bar.wait()


class Job:
    """Rank-local handle of one job on `comm.world` GPUs: arena set-up once, then any number of passes."""

    def __init__(self, ctx, comm, n_local, torch):
        self.ctx, self.comm, self.torch = ctx, comm, torch
        self.rank, self.world = comm.rank, comm.world
        counts = comm.all_gather_object(int(n_local))
        self.n_local = int(n_local)
        self.base = int(sum(counts[: self.rank]))
        self.n_total = int(sum(counts))
        handle, ptr = ctx.job_init(self.rank, self.world, self.n_total, self.base, self.n_local)
        if comm.local:
            ctx.job_connect(local_ptrs=comm.all_gather_object(ptr))
            ctx.job_set_barrier(comm.g.bar.wait)
        else:
            ctx.job_connect(handles=comm.all_gather_object(handle))
        ctx.set_pool_exchange(comm.all_reduce_min_i64)

    def stage1(self, clean, device=False):
        """Pack + replicate this rank's slice, build the dictionaries, walk.  `clean`: uint8 array (host) or device pointer."""
        if device:
            self.ctx.job_load_reads_device(clean, self.n_local)
        else:
            self.ctx.job_load_reads(clean, self.n_local)
        self.ctx.build_dicts()
        return self.ctx.reorder()

    def stage2(self, N_local):
        """Pool = singletons of all ranks ++ reads with N of all slices (N_local: this slice's lines, uint8 device tensor or
        host array); encode this rank's chains."""
        torch = self.torch
        ptr, cnt = self.ctx.device_result("singleton_ids")
        ids = self.comm.all_gather_u32(ptr, cnt)
        if not torch.is_tensor(N_local):
            N_local = torch.from_numpy(np.ascontiguousarray(N_local)).cuda() if len(N_local) else torch.empty(0, dtype=torch.uint8, device="cuda")
        allN = self.comm.all_gather_bytes(N_local) if self.world > 1 else N_local
        n_N = allN.numel() // (self.ctx.L + 1)
        pad = torch.zeros(allN.numel() + 16, dtype=torch.uint8, device="cuda")  # the pack kernel reads whole 16-byte words
        pad[: allN.numel()] = allN
        torch.cuda.synchronize()
        self._keep = (ids, pad)
        self.ctx.load_pool_ids(ids.data_ptr() if ids.numel() else None, pad.data_ptr() if n_N else None, n_N, n_s=int(ids.numel()))
        es = self.ctx.encode()
        return dict(sizes=es, pool=int(ids.numel()) + n_N)

    def run(self, clean, N_local, device=False):
        m, s, u = self.stage1(clean, device)
        res = self.stage2(N_local)
        res["counts"] = (m, s, u)
        return res


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


def write_set(out, k, s):
    """File set k (harcgpu_get_set) as the files of SURVEY Appendix A."""
    for key, stem in (("seq", "read_seq.txt"), ("pos", "read_pos.txt"), ("noise", "read_noise.txt"),
                      ("noisepos", "read_noisepos.txt"), ("rev", "read_rev.txt")):
        s[key].tofile(os.path.join(out, "%s.%d" % (stem, k)))
    s["seq_tail"].tofile(os.path.join(out, "read_seq.txt.%d.tail" % k))
    s["rev_tail"].tofile(os.path.join(out, "read_rev.txt.%d.tail" % k))


def write_globals(out, g, L):
    g["order"].tofile(os.path.join(out, "read_order.bin"))
    g["order_N"].tofile(os.path.join(out, "read_order_N_pe.bin"))
    g["singleton"].tofile(os.path.join(out, "read_singleton.txt"))
    g["singleton_tail"].tofile(os.path.join(out, "read_singleton.txt.tail"))
    g["input_N"].tofile(os.path.join(out, "input_N.dna"))
    with open(os.path.join(out, "read_meta.txt"), "w") as f:
        f.write("%d\n" % L)


def write_outputs(basedir, rank, world, res, L, comm):
    """Write the stage II files of SURVEY Appendix A under <basedir>/output/: file set `rank` by every rank, the global
    streams by rank 0."""
    out = os.path.join(basedir, "output")
    os.makedirs(out, exist_ok=True)
    write_set(out, rank, res["set"])
    parts = comm.all_gather_object(res["glob"])
    if rank == 0:
        write_globals(out, assemble_globals(parts, L), L)
    comm.barrier()
