#!/usr/bin/env python
"""bench.py -- Mreads/s of HARC's reorder+encode hot path (stage I + stage II) on B200.

A "step" is one full pass of the hot path over one synthetic FASTQ's worth of reads: 2-bit pack -> dictionary build
-> chain walk -> finalize -> stage II (pool dictionaries, consensus, re-alignment, emission, packbits).
Workload at N=1: BASELINE.json configs[1] -- 35 M x 100 bp reads, 1 % substitutions incl. N (gen_fastq_noRC -e read
model), from a synthetic 50 Mbp genome.  At N>1 every rank owns an independent read set of that size (weak scaling,
no data-path collective); `value` = reads of all ranks / max-over-ranks time.

  value : device-timed, inputs (ASCII reads) already resident in HBM.
  e2e   : the same pass through the C ABI with HOST buffers: H2D of the reads and D2H of every stage II stream
          are inside the timed region.
  --impl reference : the reference's own reorder.out + encoder.out (oracle/_ref, built unmodified from the
          reference sources) on the host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import numpy as np

METRIC = "Mreads/s reorder+encode (100bp)"
L = 100


def algorithmic_bytes_per_clean_read(genome, n_clean):
    """walk-kernel share of B_alg = 32 B per dictionary probe x P + candidate fetch R + claim RMW 4 B +
    8 B record, with P = 2*numdict*(g+1) + numdict, g = min(maxmatch, genome/N_clean)."""
    g = min(L // 2, genome / max(1, n_clean))
    P = 2 * 2 * (g + 1) + 2
    return 32.0 * P + 32 + 4 + 8, P


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop = threading.Event()
        self.sm = []
        self.smmax = 0
        self.reasons = set()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop.is_set():
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                   stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, timeout=5).stdout.decode().strip()
                f = [x.strip() for x in o.split(",")]
                self.sm.append(float(f[0]))
                self.smmax = float(f[1])
                for nme, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nme)
            except Exception:
                pass
            self.stop.wait(0.2)

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.smmax or None,
                "reasons": sorted(self.reasons)}


def run_reference(args):
    """The reference's own CPU implementation of the path, all host threads it can use, bounded sample."""
    import refrun as R
    import workload as W
    avail = R.ref_threads_available(L)
    if not avail:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built (needs /root/reference at build time)"}))
        return
    ncpu = os.cpu_count() or 1
    T = max([t for t in avail if t <= ncpu] or [min(avail)])
    n = int(args.ref_reads)
    genome = int(args.genome * (n / args.reads))
    w = W.make(n, L, genome, rc=False, errors=True, seed=args.seed)
    times = []
    for it in range(args.warmup + args.steps):
        tmp = tempfile.mkdtemp(prefix="harcref")
        try:
            W.write_dir(w, tmp)
            t1, _ = R.reorder(tmp, L, T)
            t2, _ = R.encoder(tmp, L, T)
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
        if it >= args.warmup:
            times.append(t1 + t2)
    ms = 1000.0 * sum(times) / len(times)
    val = n / (ms / 1000.0) / 1e6
    sample = "%d reads x %d bp, %d bp genome (same coverage and error model as the workload), reorder.out + encoder.out wall time incl. their file I/O" % (n, L, genome)
    print(json.dumps({
        "metric": METRIC, "value": val, "unit": "Mreads/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "impl": "reference",
        "config": {"workload": "configs[1]: 35M x 100bp, 1% substitutions incl. N (gen_fastq_noRC -e model), 50 Mbp genome; timed on a bounded sample", "sample": sample},
        "cpu_baseline": {"value": val, "unit": "Mreads/s", "cores": T, "kind": "reference", "sample": sample},
        "e2e": {"value": val, "unit": "Mreads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cpu_baseline(args):
    import refrun as R
    import workload as W
    avail = R.ref_threads_available(L)
    if not avail:
        return None
    ncpu = os.cpu_count() or 1
    T = max([t for t in avail if t <= ncpu] or [min(avail)])
    n = int(args.ref_reads)
    genome = int(args.genome * (n / args.reads))
    w = W.make(n, L, genome, rc=False, errors=True, seed=args.seed)
    tmp = tempfile.mkdtemp(prefix="harcref")
    try:
        W.write_dir(w, tmp)
        t1, _ = R.reorder(tmp, L, T)
        t2, _ = R.encoder(tmp, L, T)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return {"value": n / (t1 + t2) / 1e6, "unit": "Mreads/s", "cores": T, "kind": "reference",
            "sample": "%d reads x %d bp, %d bp genome (same coverage/error model), reorder.out %.1fs + encoder.out %.1fs incl. file I/O"
                      % (n, L, genome, t1, t2)}


def time_ingest(torch, harc_b200, all_lines, n, device):
    """FASTQ of n reads (fixed-width ids, constant qualities) resident in HBM -> packed clean reads + N reads."""
    rec = 11 + (L + 1) + 2 + (L + 1)
    fq = np.empty((n, rec), dtype=np.uint8)
    num = np.arange(n, dtype=np.int64)
    fq[:, 0] = ord("@")
    for k in range(9):
        fq[:, 9 - k] = (num // 10 ** k) % 10 + 48
    fq[:, 10] = 10
    fq[:, 11:11 + L + 1] = all_lines[: n * (L + 1)].reshape(n, L + 1)
    fq[:, 11 + L + 1] = ord("+")
    fq[:, 11 + L + 2] = 10
    fq[:, 11 + L + 3: rec - 1] = ord("I")
    fq[:, rec - 1] = 10
    d = torch.empty(fq.size + 16, dtype=torch.uint8, device="cuda")
    d[: fq.size].copy_(torch.from_numpy(fq.reshape(-1)))
    ctx = harc_b200.HarcGpu(L, device=device)
    for _ in range(2):
        info = ctx.ingest_fastq_device(d.data_ptr(), fq.size)
    ms = []
    for _ in range(3):
        info = ctx.ingest_fastq_device(d.data_ptr(), fq.size)
        ms.append(ctx.last_ms("ingest"))
    want_N = int((all_lines[: n * (L + 1)].reshape(n, L + 1) == ord("N")).any(axis=1).sum())
    assert info["total_reads"] == n and info["n_N"] == want_N, (info, want_N)
    ctx.close()
    t = float(np.median(ms))
    return {"what": "fused FASTQ ingest (preprocess.cpp + readDnaFile), FASTQ resident in HBM", "reads": n, "fastq_bytes": int(fq.size),
            "ms": t, "Mreads_per_s": n / t / 1e3, "GB_per_s": fq.size / t / 1e6}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=float, default=35e6)
    ap.add_argument("--genome", type=float, default=50e6)
    ap.add_argument("--ref-reads", type=float, default=3.5e6)
    ap.add_argument("--walkers", type=int, default=0)
    ap.add_argument("--file-sets", type=int, default=1)
    ap.add_argument("--reads-per-walker", type=int, default=0)
    ap.add_argument("--extend", type=int, default=0)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--rc", type=int, default=0, help="1: odd reads reverse-complemented (gen_fastq model, configs[2]); default gen_fastq_noRC")
    ap.add_argument("--errors", type=int, default=1, help="0: error-free reads (configs[0])")
    ap.add_argument("--mode", default="read-sets", choices=["read-sets", "single-job"],
                    help="N>1: read-sets = every rank compresses its own read set (weak scaling, no data-path collective); "
                         "single-job = ONE read set of --reads on all ranks (strong scaling: shared claim bitmap over NVLink peer "
                         "memory, all-gather of singleton ids, all-reduce(min) of pool claims)")
    ap.add_argument("--shard-dicts", type=int, default=0,
                    help="single-job mode: 1 = dictionaries sharded by key hash over the GPUs (probed through NVLink peer memory)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the host-buffer pass")
    ap.add_argument("--ingest-reads", type=float, default=8e6,
                    help="reads of the workload that are also laid out as a FASTQ file to time the fused ingest (0 = skip)")
    args = ap.parse_args()
    args.reads = int(args.reads)
    args.genome = int(args.genome)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank == 0:
            run_reference(args)
        return

    # Rank 0 prints exactly ONE line on stdout, the JSON.  Everything else that writes to file descriptor 1 (NCCL's version
    # banner, library chatter) is sent to stderr: fd 1 is pointed at stderr for the run and the JSON goes to the saved fd.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import harc_b200
    import workload as W
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    # ---- workload: every rank owns an independent read set (seed differs per rank)
    single_job = args.mode == "single-job" and world > 1
    w = W.make(args.reads, L, args.genome, rc=bool(args.rc), errors=bool(args.errors), seed=args.seed + (0 if single_job else 7919 * rank))
    n_clean, n_N = w["n_clean"], w["n_N"]
    h_clean = torch.from_numpy(w["clean"]).pin_memory()
    h_N = torch.from_numpy(w["withN"]).pin_memory()
    d_clean = torch.empty(h_clean.numel() + 16, dtype=torch.uint8, device="cuda")
    d_N = torch.empty(h_N.numel() + 16, dtype=torch.uint8, device="cuda")
    d_clean[: h_clean.numel()].copy_(h_clean)
    d_N[: h_N.numel()].copy_(h_N)
    torch.cuda.synchronize()
    # auxiliary figure: the fused FASTQ ingest (SURVEY f-1) on the first --ingest-reads reads laid out as a FASTQ file
    ingest = None
    n_ing = min(int(args.ingest_reads), args.reads) if world == 1 else 0
    if n_ing:
        ingest = time_ingest(torch, harc_b200, w["all"], n_ing, local)
    del w["all"]

    ctx = harc_b200.HarcGpu(L, device=local, walkers=args.walkers, file_sets=args.file_sets,
                            reads_per_walker=args.reads_per_walker, extend=args.extend,
                            shard_dicts=args.shard_dicts if single_job else 0)
    stream = torch.cuda.ExternalStream(ctx.stream())
    phases = ["pack", "dict", "walk", "finalize", "pooldict", "encode"]

    if single_job:
        from harc_b200 import multi
        ctx.load_reads_device(d_clean.data_ptr(), n_clean)
        handles = [None] * world
        dist.all_gather_object(handles, ctx.shard_init(rank, world, n_clean))
        ctx.shard_connect(handles)

    def step_device():
        ctx.load_reads_device(d_clean.data_ptr(), n_clean)
        ctx.build_dicts()
        if single_job:
            return multi.run_pass(ctx, dist, d_N.data_ptr(), rank, world, torch, n_N)["sizes"]
        ctx.reorder()
        ctx.load_pool_device(d_N.data_ptr(), n_N)
        return ctx.encode()

    d2h = [0]

    host_ms = {}
    # pinned landing area for the stage II streams (the caller owns every host buffer of the C ABI)
    h_out = torch.empty(int((n_clean + n_N) * 16 + (64 << 20)), dtype=torch.uint8).pin_memory().numpy()
    cur = [0]

    def pinned_empty(count, dtype):
        nb = int(count) * np.dtype(dtype).itemsize
        a = (cur[0] + 63) // 64 * 64
        if a + nb > h_out.size:
            return np.empty(int(count), dtype)
        cur[0] = a + nb
        return h_out[a:a + nb].view(dtype)

    hN_np = h_N.numpy()

    def step_host():
        t = [time.perf_counter()]

        def lap(name):
            t.append(time.perf_counter())
            host_ms[name] = host_ms.get(name, 0.0) + 1000.0 * (t[-1] - t[-2])
        ctx.load_reads(h_clean.numpy(), n_clean)
        lap("load_reads(H2D+pack)")
        if single_job:
            ctx.build_dicts()
            multi.run_pass(ctx, dist, hN_np, rank, world, torch)
            lap("reorder+exchange+encode")
        else:
            ctx.stage_nreads(hN_np)  # upload of the reads with N overlaps stage I
            ctx.reorder()
            lap("reorder")
            ctx.load_pool(None, None, hN_np)
            lap("load_pool(H2D+dict)")
            ctx.encode()
            lap("encode")
        nbytes = 0
        cur[0] = 0
        for k in range(args.file_sets):
            s = ctx.get_set(k, pinned_empty)
            nbytes += sum(v.nbytes for v in s.values())
        g = ctx.get_globals(pinned_empty)
        nbytes += sum(v.nbytes for v in g.values())
        lap("get outputs(D2H)")
        d2h[0] = nbytes

    alloc_stats = {}

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, sampler=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ph = {p: 0.0 for p in phases}
        if sampler:
            sampler.start()
        l0 = harc_b200.launch_count()
        m0 = ctx.last_ms("cudaMalloc_calls")
        e0.record(stream)
        for _ in range(steps):
            fn()
            for p in phases:
                ph[p] += max(0.0, ctx.last_ms(p))
        e1.record(stream)
        barrier()
        if sampler:
            sampler.stop.set()
        ms = e0.elapsed_time(e1) / steps
        if dist is not None:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        alloc_stats["cudaMalloc_calls_in_timed_region"] = int(ctx.last_ms("cudaMalloc_calls") - m0)
        alloc_stats["device_peak_MB"] = ctx.last_ms("peak_MB")
        alloc_stats["device_cached_MB"] = ctx.last_ms("cached_MB")
        return ms, {p: ph[p] / steps for p in phases}, (harc_b200.launch_count() - l0) // steps

    for _ in range(args.warmup):
        es = step_device()
    sampler = ClockSampler(local)
    ms_dev, ph, launches = timed(step_device, args.steps, sampler)
    alloc_dev = dict(alloc_stats)
    cnt = ctx.counters()
    m, s, u = ctx.reorder_counts()
    if args.no_e2e:
        ms_e2e = float("nan")
    else:
        step_host()  # warm the host path
        step_host()
        host_ms.clear()
        ms_e2e, _, _ = timed(step_host, args.steps)

    total_reads = (n_clean + n_N) * (1 if single_job else world)
    value = total_reads / (ms_dev / 1000.0) / 1e6
    e2e = total_reads / (ms_e2e / 1000.0) / 1e6
    if rank != 0:
        ctx.close()
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    balg, P = algorithmic_bytes_per_clean_read(args.genome, n_clean)
    walk_ms = ph["walk"]
    achieved = n_clean * balg / (walk_ms / 1000.0) / 1e9 if walk_ms > 0 else 0.0
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "walk_traffic.json"))).get("dram_bytes_per_launch")
    except Exception:
        pass
    out = {
        "metric": METRIC, "value": value, "unit": "Mreads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "strong" if single_job else "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "%s: %d x %dbp reads, %s (%s model), %d bp synthetic genome, %s"
                               % ("configs[1]" if (args.reads, args.genome, args.rc, args.errors) == (35000000, 50000000, 0, 1) else "custom",
                                  args.reads, L, "1% substitutions incl. N" if args.errors else "error-free",
                                  ("gen_fastq" if args.rc else "gen_fastq_noRC") + (" -e" if args.errors else ""), args.genome,
                                  "ONE read set on all GPUs" if single_job else "per GPU"),
                   "reads_per_gpu": args.reads, "clean_reads": n_clean, "reads_with_N": n_N, "walkers": ctx.p.walkers or "auto",
                   "file_sets": args.file_sets, "parallelism": ("single GPU" if world == 1 else ("one job: claim bitmap%s in NVLink peer memory + all-gather/all-reduce(min) of pool claims" % (" and dictionary shards" if args.shard_dicts else ""))
                                   if single_job else "1 independent read set per GPU"),
                   "l2": "inputs (%.1f GB ASCII, %.1f GB packed) exceed the 126 MB L2; no explicit flush" % ((n_clean + n_N) * 101 / 1e9, n_clean * 32 / 1e9)},
        "phases_ms": ph,
        "stage1": {"matched": m, "singletons": s, "chain_heads": u, "probes_per_read": cnt["probes"] / max(1, n_clean),
                   "compares_per_read": cnt["compares"] / max(1, n_clean), "claim_fails": cnt["claim_fails"],
                   "search_rounds_from_shift_0": cnt["steps"], "harvested": cnt["harvested"]},
        "stage2": {"aligned_singletons": es.aligned_singletons, "aligned_N": es.aligned_N},
        "roofline": {"bound": "hbm", "kernel": "walk_kernel<4, 32> (one walker per warp)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak if peak else None, "traffic": traffic,
                     "algorithmic_bytes_per_clean_read": balg, "model_probes_per_read": P,
                     # the stricter whole-step figure, compulsory once-through bytes only (B_stream = 3R + 99 = 195 B
                     # per read at L = 100), over the whole device-timed step of this rank
                     "stream_only": {"bytes_per_read": 3 * 32 + 99, "achieved": (n_clean + n_N) * (3 * 32 + 99) / (ms_dev / 1000.0) / 1e9,
                                     "frac": (n_clean + n_N) * (3 * 32 + 99) / (ms_dev / 1000.0) / 1e9 / peak if peak else None},
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s"},
        "e2e": {"value": e2e, "unit": "Mreads/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(h_clean.numel() + h_N.numel()),
                "d2h_bytes_per_step": int(d2h[0]),
                "host_wall_ms": {k: v / args.steps for k, v in host_ms.items()}},
        "gpu_launches": int(launches), "allocator": dict(alloc_dev),
        "clocks": sampler.summary(),
    }
    if ingest:
        out["ingest"] = ingest
    if not args.no_cpu_baseline and world == 1:
        try:
            out["cpu_baseline"] = cpu_baseline(args)
        except Exception as ex:  # the reference binaries are a reported baseline, never a dependency of the GPU number
            out["cpu_baseline"] = {"error": str(ex)[:200]}
    os.write(json_fd, (json.dumps(out) + "\n").encode())
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
