#!/usr/bin/env python
"""bench.py -- Mreads/s of HARC's reorder+encode hot path (stage I + stage II) on B200.

A "step" is one full pass of the hot path over one synthetic FASTQ's worth of reads: 2-bit pack -> dictionary build
-> chain walk -> finalize -> stage II (pool dictionaries, consensus, re-alignment, emission, packbits).

  N = 1 : BASELINE.json configs[1] -- 35 M x 100 bp reads, 1 % substitutions incl. N, 50 Mbp genome (--config picks another).
  N > 1 : ONE job on N GPUs (strong scaling), BASELINE.json configs[2] -- 200 M x 100 bp reads with reverse complements,
          1 Gbp genome: every rank uploads 1/N of the reads, the packed reads are replicated and the (key, id) pairs of the
          dictionary build exchanged over NVLink by the library's own kernels, the dictionaries are sharded by key, the
          claim bitmap is shared, stage II runs per rank on its own chains (harc_b200/multi.py, csrc/job.cu).
          `--mode read-sets` is the other split: one independent read set per GPU (weak scaling, no data-path collective).

  value : device-timed, inputs (ASCII reads) already resident in HBM.
  e2e   : the same pass through the C ABI with HOST buffers: H2D of the reads and D2H of every stage II stream are inside
          the timed region.  `e2e.value` is the throughput of a pipeline of two jobs in flight (two contexts, so the copies of
          one job overlap the kernels of the other); `e2e.single_job` is one job at a time (latency).
  bits_per_base : stage III stand-in (bz2/xz, oracle/refrun.py) over the streams of the benched run, ours and the reference's.
  --impl reference : the reference's own reorder.out + encoder.out (oracle/_ref, built unmodified from the reference
          sources) on the host cores, on the SAME workload, full size.
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


def algorithmic_bytes_per_clean_read(L, genome, n_clean):
    """walk-kernel share of B_alg = 32 B per dictionary probe x P + candidate fetch R + claim RMW 4 B +
    8 B record, with P = 2*numdict*(g+1) + numdict, g = min(maxmatch, genome/N_clean), R = 8 * ceil(2L/64)."""
    g = min(L // 2, genome / max(1, n_clean))
    P = 2 * 2 * (g + 1) + 2
    R = 8 * ((2 * L + 63) // 64)
    return 32.0 * P + R + 4 + 8, P


def ncpu():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


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


def pick_config(args, world):
    import workload as W
    one_job = world > 1 and args.mode == "one-job"
    k = args.config if args.config is not None else (2 if one_job else 1)
    cfg = dict(W.CONFIGS[k])
    cfg["index"] = k
    custom = False
    for key, val in (("reads", args.reads), ("genome", args.genome), ("rc", args.rc), ("errors", args.errors)):
        if val is not None:
            cfg[key] = type(W.CONFIGS[k][key])(val)
            custom = True
    if custom:
        cfg["name"] = "custom (from configs[%d]): %d x %dbp reads, %s, %s, %d bp genome" % (
            k, cfg["reads"], cfg["L"], "1% substitutions incl. N" if cfg["errors"] else "error-free",
            "reverse complements" if cfg["rc"] else "no reverse complements", cfg["genome"])
    return cfg


def workload_signature(cfg, seed):
    return "reads=%d,L=%d,genome=%d,rc=%d,errors=%d,seed=%d" % (cfg["reads"], cfg["L"], cfg["genome"], int(cfg["rc"]), int(cfg["errors"]), seed)


def scratch_dir(need_bytes):
    """A directory with room for the reference's intermediate files (RAM-backed if it fits)."""
    for base in ("/dev/shm", tempfile.gettempdir(), ROOT):
        try:
            if shutil.disk_usage(base).free > need_bytes:
                return tempfile.mkdtemp(prefix="harcref", dir=base)
        except Exception:
            pass
    return None


def reference_threads(L):
    import refrun as R
    avail = R.ref_threads_available(L)
    if not avail:
        return None
    n = ncpu()
    return max([t for t in avail if t <= n] or [min(avail)])


def reference_pass(w, L, T, keep=False):
    """One full-size run of the reference's reorder.out + encoder.out on workload w.  Returns (seconds, dir or None)."""
    import refrun as R
    import workload as W
    need = 3 * (w["clean"].nbytes + w["withN"].nbytes) + (1 << 30)
    tmp = scratch_dir(need)
    if tmp is None:
        raise RuntimeError("no scratch directory with %.1f GB free for the reference's files" % (need / 1e9))
    try:
        W.write_dir(w, tmp)
        t1, _ = R.reorder(tmp, L, T, timeout=3000)
        t2, _ = R.encoder(tmp, L, T, timeout=3000)
    except BaseException:
        shutil.rmtree(tmp, ignore_errors=True)
        raise
    if not keep:
        shutil.rmtree(tmp, ignore_errors=True)
        tmp = None
    return t1, t2, tmp


def run_reference(args, world):
    """--impl reference: the reference's own CPU implementation of the path, all host threads it can use, on the SAME
    workload as our arm, full size.  A step of configs[1] takes ~25 s on 32 cores, so the run is bounded by wall time:
    at most --steps timed steps, fewer when the budget (--ref-budget-s) runs out -- never a smaller sample."""
    import workload as W
    cfg = pick_config(args, world)
    L = cfg["L"]
    T = reference_threads(L)
    if T is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built (needs /root/reference at build time)"}))
        return
    t_start = time.time()
    try:
        w = W.make(cfg["reads"], L, cfg["genome"], rc=cfg["rc"], errors=cfg["errors"], seed=args.seed, threads=ncpu(), keep_all=False)
        times, warm = [], 0
        want_warm = min(args.warmup, 1)
        est = None
        while len(times) < args.steps:
            if est is not None and time.time() - t_start + est > args.ref_budget_s:
                break
            t1, t2, _ = reference_pass(w, L, T)
            est = t1 + t2 + 5.0
            # a warm-up step only if at least one timed step still fits afterwards
            if warm < want_warm and not times and time.time() - t_start + est <= args.ref_budget_s:
                warm += 1
                continue
            times.append(t1 + t2)
    except Exception as ex:
        print(json.dumps({"impl": "reference", "unavailable": "reference run failed: %s" % str(ex)[:300]}))
        return
    ms = 1000.0 * sum(times) / len(times)
    val = cfg["reads"] / (ms / 1000.0) / 1e6
    sample = ("full workload, %d timed step(s) of %d asked for (+%d warm-up): a step takes %.0f s on %d threads and the run is bounded to "
              "%d s of wall time; reorder.out + encoder.out wall time incl. their file I/O" % (len(times), args.steps, warm, ms / 1000.0, T, args.ref_budget_s))
    print(json.dumps({
        "metric": METRIC, "value": val, "unit": "Mreads/s", "n_gpus": args.gpus, "steps": len(times), "warmup": warm,
        "steps_requested": args.steps, "warmup_requested": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if (world > 1 and args.mode == "one-job") else "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic", "impl": "reference",
        "config": {"workload": cfg["name"] + " (the whole workload, same as the GPU arm; the CPU run does not use the GPUs)", "sample": sample,
                   "reference_build": "unmodified reorder.cpp / encoder.cpp, g++ -O3 -march=x86-64-v3 -fopenmp (built where /root/reference is mounted; "
                                      "harc:65 uses -march=native, which a binary that travels cannot)"},
        "cpu_baseline": {"value": val, "unit": "Mreads/s", "cores": T, "kind": "reference", "sample": sample},
        "e2e": {"value": val, "unit": "Mreads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def time_ingest(torch, harc_b200, all_lines, n, L, device):
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


def git_head():
    try:
        return subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                              timeout=5).stdout.decode().strip() or None
    except Exception:
        return None


def roofline_block(cfg, n_clean_per_gpu, n_clean_total, n_all_per_gpu, walk_ms, ms_dev, kernel):
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    L = cfg["L"]
    balg, P = algorithmic_bytes_per_clean_read(L, cfg["genome"], n_clean_total)
    achieved = n_clean_per_gpu * balg / (walk_ms / 1000.0) / 1e9 if walk_ms > 0 else 0.0
    traffic, tnote = None, None
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "walk_traffic.json")))
        if t.get("workload_signature", "").startswith("reads=%d,L=%d,genome=%d" % (cfg["reads"], L, cfg["genome"])):
            traffic = t.get("dram_bytes_per_launch")
            tnote = "ncu --set full capture of commit %s (%s); not re-measured by this run" % (t.get("commit"), t.get("source"))
    except Exception:
        pass
    bstream = 3 * 8 * ((2 * L + 63) // 64) + 99
    return {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None,
            "traffic": traffic, "traffic_note": tnote,
            "algorithmic_bytes_per_clean_read": balg, "model_probes_per_read": P, "clean_reads_per_launch": n_clean_per_gpu,
            # the stricter whole-step figure, compulsory once-through bytes only (B_stream = 3R + 99), over the
            # whole device-timed step of this GPU
            "stream_only": {"bytes_per_read": bstream, "achieved": n_all_per_gpu * bstream / (ms_dev / 1000.0) / 1e9,
                            "frac": n_all_per_gpu * bstream / (ms_dev / 1000.0) / 1e9 / peak if peak else None},
            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s"}


def write_archive(dst, sets, glob, L):
    from harc_b200 import multi
    out = os.path.join(dst, "output")
    os.makedirs(out, exist_ok=True)
    for k, s in enumerate(sets):
        multi.write_set(out, k, s)
    multi.write_globals(out, glob, L)


def bits_block(cfg, seed, gpu_bytes, ref_bytes, ref_T):
    nb = float(cfg["reads"]) * cfg["L"]
    bpb = lambda b: None if b is None else 8.0 * b / nb
    t1 = None
    try:
        fx = json.load(open(os.path.join(ROOT, "tests", "golden", "bits_ref_t1.json")))
        t1 = fx.get(workload_signature(cfg, seed))
    except Exception:
        pass
    t1_bytes = t1["standin_bytes"] if t1 else None
    return {"gpu": bpb(gpu_bytes), "ref_tN": bpb(ref_bytes), "ref_t1": bpb(t1_bytes), "ref_tN_threads": ref_T,
            "gpu_over_ref_tN": gpu_bytes / ref_bytes if ref_bytes else None,
            "gpu_over_ref_t1": gpu_bytes / t1_bytes if t1_bytes else None,
            "bytes": {"gpu": gpu_bytes, "ref_tN": ref_bytes, "ref_t1": t1_bytes},
            "how": "stage III stand-in (bz2 -9 for the bsc streams, xz for the 7z streams; oracle/refrun.py standin_size) over the streams of the "
                   "benched configuration, order-free mode; ref_tN = the reference run of cpu_baseline in this run; ref_t1 = the reference at "
                   "num_thr=1 (deterministic), %s" % ("from tests/golden/bits_ref_t1.json (tools/ref_t1_bits.py, %s)" % t1.get("made") if t1
                                                      else "not available for this workload (run tools/ref_t1_bits.py)")}


# ------------------------------------------------------------------------------------------------ one GPU / read sets
def run_single(args, rank, world, local, dist, json_fd):
    import torch
    import harc_b200
    import workload as W
    cfg = pick_config(args, world)
    L = cfg["L"]
    w = W.make(cfg["reads"], L, cfg["genome"], rc=cfg["rc"], errors=cfg["errors"], seed=args.seed + 7919 * rank, threads=max(1, ncpu() // world))
    n_clean, n_N = w["n_clean"], w["n_N"]
    h_clean = torch.from_numpy(w["clean"]).pin_memory()
    h_N = torch.from_numpy(w["withN"]).pin_memory()
    d_clean = torch.empty(h_clean.numel() + 16, dtype=torch.uint8, device="cuda")
    d_N = torch.empty(h_N.numel() + 16, dtype=torch.uint8, device="cuda")
    d_clean[: h_clean.numel()].copy_(h_clean)
    d_N[: h_N.numel()].copy_(h_N)
    torch.cuda.synchronize()
    ingest = None
    n_ing = min(int(args.ingest_reads), cfg["reads"]) if world == 1 else 0
    if n_ing:
        ingest = time_ingest(torch, harc_b200, w["all"], n_ing, L, local)
    w["all"] = None
    preserve = bool(cfg["preserve"])

    def new_ctx():
        return harc_b200.HarcGpu(L, device=local, walkers=args.walkers, file_sets=args.file_sets,
                                 reads_per_walker=args.reads_per_walker, extend=args.extend)
    ctx = new_ctx()
    stream = torch.cuda.ExternalStream(ctx.stream())
    phases = ["pack", "dict", "walk", "finalize", "pooldict", "encode"]

    def step_device():
        ctx.load_reads_device(d_clean.data_ptr(), n_clean)
        ctx.build_dicts()
        ctx.reorder()
        ctx.load_pool_device(d_N.data_ptr(), n_N)
        return ctx.encode()

    hN_np, hC_np = h_N.numpy(), h_clean.numpy()

    class HostJob:
        """One context + its pinned landing area for the stage II streams (the caller owns every host buffer of the C ABI)."""

        def __init__(self, c):
            self.c = c
            self.buf = torch.empty(int((n_clean + n_N) * 16 + (64 << 20)), dtype=torch.uint8).pin_memory().numpy()
            self.cur = 0
            self.d2h = 0
            self.ms = {}
            self.sets, self.glob = None, None

        def empty(self, count, dtype):
            nb = int(count) * np.dtype(dtype).itemsize
            a = (self.cur + 63) // 64 * 64
            if a + nb > self.buf.size:
                return np.empty(int(count), dtype)
            self.cur = a + nb
            return self.buf[a:a + nb].view(dtype)

        def step(self, n_first=False):
            """One job.  n_first: queue the upload of the reads with N in front of the clean reads instead of behind them.
            One job at a time, behind is better (it runs under stage I); with several jobs in flight on one link it would
            land behind the other jobs' 50 ms uploads and stage II would wait for it."""
            c = self.c
            t = [time.perf_counter()]

            def lap(name):
                t.append(time.perf_counter())
                self.ms[name] = self.ms.get(name, 0.0) + 1000.0 * (t[-1] - t[-2])
            if n_first:
                c.stage_nreads(hN_np)
            c.load_reads(hC_np, n_clean)
            lap("load_reads(H2D+pack)")
            if not n_first:
                c.stage_nreads(hN_np)  # upload of the reads with N overlaps stage I
            c.reorder()
            lap("reorder")
            c.load_pool(None, None, hN_np)
            lap("load_pool(H2D+dict)")
            c.encode()
            lap("encode")
            self.cur = 0
            self.sets = [c.get_set(k, self.empty) for k in range(args.file_sets)]
            self.glob = c.get_globals(self.empty)
            nbytes = sum(v.nbytes for s in self.sets for v in s.values()) + sum(v.nbytes for v in self.glob.values())
            if preserve:  # -p: pack_order.cpp on the device (harc:111-112), the packed order stream is part of the result
                self.packed = c.get_packed_order()
                nbytes += sum(v.nbytes for v in self.packed)
            lap("get outputs(D2H)")
            self.d2h = nbytes

    alloc_stats = {}

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, c, sampler=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ph = {p: 0.0 for p in phases}
        if sampler:
            sampler.start()
        l0 = harc_b200.launch_count()
        m0 = c.last_ms("cudaMalloc_calls")
        e0.record(stream)
        for _ in range(steps):
            fn()
            for p in phases:
                ph[p] += max(0.0, c.last_ms(p))
        e1.record(stream)
        barrier()
        if sampler:
            sampler.stop.set()
        ms = e0.elapsed_time(e1) / steps
        if dist is not None:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        alloc_stats["cudaMalloc_calls_in_timed_region"] = int(c.last_ms("cudaMalloc_calls") - m0)
        alloc_stats["device_peak_MB"] = c.last_ms("peak_MB")
        alloc_stats["device_cached_MB"] = c.last_ms("cached_MB")
        return ms, {p: ph[p] / steps for p in phases}, (harc_b200.launch_count() - l0) // steps

    es = None
    for _ in range(args.warmup):
        es = step_device()
    sampler = ClockSampler(local)
    ms_dev, ph, launches = timed(step_device, args.steps, ctx, sampler)
    if es is None:
        es = step_device()
    alloc_dev = dict(alloc_stats)
    cnt = ctx.counters()
    m, s, u = ctx.reorder_counts()

    # ---- e2e: one job at a time, then two jobs in flight (two contexts, one host thread each)
    ms_e2e = ms_pipe = float("nan")
    job = None
    pipe_note = pipe_host = None
    if not args.no_e2e:
        job = HostJob(ctx)
        job.step()  # warm the host path
        job.step()
        job.ms.clear()
        ms_e2e, _, _ = timed(job.step, args.steps, ctx)
        host_ms = {k: v / args.steps for k, v in job.ms.items()}
        if args.pipeline > 1 and dist is None:
            jobs = [job] + [HostJob(new_ctx()) for _ in range(args.pipeline - 1)]
            streams = [torch.cuda.ExternalStream(j.c.stream()) for j in jobs]
            for j in jobs[1:]:
                j.step(True)
                j.step(True)
            for j in jobs:
                j.ms.clear()
            per = max(2, args.steps)

            def worker(j):
                torch.cuda.set_device(local)
                for _ in range(per):
                    j.step(True)
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True)
            ends = [torch.cuda.Event(enable_timing=True) for _ in jobs]
            e0.record(streams[0])
            ts = [threading.Thread(target=worker, args=(j,)) for j in jobs]
            for t in ts:
                t.start()
            for t in ts:
                t.join()
            for e, st in zip(ends, streams):
                e.record(st)
            torch.cuda.synchronize()
            ms_pipe = max(e0.elapsed_time(e) for e in ends) / (per * len(jobs))
            pipe_note = "%d jobs in flight (one context and one host thread each), %d jobs in all" % (len(jobs), per * len(jobs))
            pipe_host = [{k: v / per for k, v in j.ms.items()} for j in jobs]
            for j in jobs[1:]:
                j.c.close()

    n_reads = n_clean + n_N
    total_reads = n_reads * world
    value = total_reads / (ms_dev / 1000.0) / 1e6
    e2e_single = total_reads / (ms_e2e / 1000.0) / 1e6
    e2e_pipe = total_reads / (ms_pipe / 1000.0) / 1e6 if ms_pipe == ms_pipe else None
    if rank != 0:
        ctx.close()
        return

    out = {
        "metric": METRIC if L == 100 else METRIC.replace("100bp", "%dbp" % L), "value": value, "unit": "Mreads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": cfg["name"] + (", per GPU" if world > 1 else ""), "workload_signature": workload_signature(cfg, args.seed),
                   "reads_per_gpu": cfg["reads"], "clean_reads": n_clean, "reads_with_N": n_N, "walkers": ctx.p.walkers or "auto",
                   "file_sets": args.file_sets, "parallelism": "single GPU" if world == 1 else "1 independent read set per GPU (no data-path collective)",
                   "l2": "inputs (%.1f GB ASCII, %.1f GB packed) exceed the 126 MB L2; no explicit flush" % (n_reads * (L + 1) / 1e9, n_clean * 8 * ctx.NW() / 1e9)},
        "phases_ms": ph,
        "stage1": {"matched": m, "singletons": s, "chain_heads": u, "probes_per_read": cnt["probes"] / max(1, n_clean),
                   "compares_per_read": cnt["compares"] / max(1, n_clean), "claim_fails": cnt["claim_fails"],
                   "search_rounds_from_shift_0": cnt["steps"], "harvested": cnt["harvested"]},
        "stage2": {"aligned_singletons": es.aligned_singletons, "aligned_N": es.aligned_N},
        "roofline": roofline_block(cfg, n_clean, n_clean, n_reads, ph["walk"], ms_dev, "walk_kernel<%d, 32> (one walker per warp)" % ctx.NW()),
        "gpu_launches": int(launches), "allocator": dict(alloc_dev),
        "clocks": sampler.summary(), "commit": git_head(),
    }
    if not args.no_e2e:
        out["e2e"] = {"value": e2e_pipe if e2e_pipe else e2e_single, "unit": "Mreads/s",
                      "ms_per_step": ms_pipe if e2e_pipe else ms_e2e,
                      "what": ("pipelined: " + pipe_note) if e2e_pipe else "one job at a time",
                      "single_job": {"value": e2e_single, "ms_per_step": ms_e2e, "host_wall_ms": host_ms},
                      "pipelined_host_wall_ms_per_job": pipe_host,
                      "h2d_bytes_per_step": int(h_clean.numel() + h_N.numel()), "d2h_bytes_per_step": int(job.d2h),
                      "pcie_note": "H2D of %.2f GB per job: at the ~55 GB/s a Gen5 x16 link delivers that alone is %.0f ms, the floor of a step"
                                   % ((h_clean.numel() + h_N.numel()) / 1e9, (h_clean.numel() + h_N.numel()) / 55e9 * 1e3)}
    if ingest:
        out["ingest"] = ingest
    # ---- reference on the same workload (full size, one step) + bits/base of both
    if not args.no_cpu_baseline and world == 1 and cfg["reads"] * L > 6e9 and not args.force_cpu_baseline:
        # the reference needs ~10 s per Gbase on 16 threads: the default bench run stays within minutes
        out["cpu_baseline"] = {"skipped": "%d reads x %d bp would take the reference several minutes to hours; run with --force-cpu-baseline "
                                          "or `--impl reference --config %d`" % (cfg["reads"], L, cfg["index"])}
    elif not args.no_cpu_baseline and world == 1:
        try:
            import refrun as R
            T = reference_threads(L)
            if T is None:
                raise RuntimeError("oracle/_ref not built")
            t1, t2, rdir = reference_pass(w, L, T, keep=not args.no_bits)
            out["cpu_baseline"] = {"value": cfg["reads"] / (t1 + t2) / 1e6, "unit": "Mreads/s", "cores": T, "kind": "reference",
                                   "sample": "the whole workload (%d reads), one run: reorder.out %.1f s + encoder.out %.1f s wall time incl. their file I/O"
                                             % (cfg["reads"], t1, t2)}
            if rdir and job is not None:
                try:
                    ref_bytes = R.standin_size(rdir)[0]
                    gdir = tempfile.mkdtemp(prefix="harcgpu", dir=os.path.dirname(rdir))
                    write_archive(gdir, job.sets, job.glob, L)
                    gpu_bytes = R.standin_size(gdir)[0]
                    shutil.rmtree(gdir, ignore_errors=True)
                    out["bits_per_base"] = bits_block(cfg, args.seed, gpu_bytes, ref_bytes, T)
                finally:
                    shutil.rmtree(rdir, ignore_errors=True)
        except Exception as ex:  # the reference binaries are a reported baseline, never a dependency of the GPU number
            out["cpu_baseline"] = {"error": str(ex)[:300]}
    os.write(json_fd, (json.dumps(out) + "\n").encode())
    ctx.close()


# ------------------------------------------------------------------------------------------------ one job on N GPUs
def run_one_job(args, rank, world, local, dist, json_fd):
    import torch
    import harc_b200
    import workload as W
    from harc_b200 import multi
    cfg = pick_config(args, world)
    L = cfg["L"]
    thr = max(1, ncpu() // world)
    genome = W.make_genome(cfg["genome"], args.seed, thr)
    a, b = W.slice_bounds(cfg["reads"], world)[rank]
    w = W.make(cfg["reads"], L, cfg["genome"], rc=cfg["rc"], errors=cfg["errors"], seed=args.seed, first=a, count=b - a, genome=genome,
               threads=thr, keep_all=False)
    n_clean, n_N = w["n_clean"], w["n_N"]
    h_clean = torch.from_numpy(w["clean"]).pin_memory()
    h_N = torch.from_numpy(w["withN"]).pin_memory()
    d_clean = torch.empty(h_clean.numel() + 16, dtype=torch.uint8, device="cuda")
    d_clean[: h_clean.numel()].copy_(h_clean)
    d_N = h_N.cuda()
    torch.cuda.synchronize()
    comm = multi.DistComm(dist, torch)
    ctx = harc_b200.HarcGpu(L, device=local, walkers=args.walkers, file_sets=1, reads_per_walker=args.reads_per_walker, extend=args.extend,
                            shard_dicts=args.shard_dicts, lanes_per_walker=args.lanes)
    job = multi.Job(ctx, comm, n_clean, torch)
    stream = torch.cuda.ExternalStream(ctx.stream())
    phases = ["pack", "dict", "walk", "finalize", "pooldict", "encode"]
    hC_np, hN_np = h_clean.numpy(), h_N.numpy()
    buf = torch.empty(int((n_clean + n_N) * 24 + (64 << 20)), dtype=torch.uint8).pin_memory().numpy()
    cur = [0]
    last = {}

    def pinned_empty(count, dtype):
        nb = int(count) * np.dtype(dtype).itemsize
        p = (cur[0] + 63) // 64 * 64
        if p + nb > buf.size:
            return np.empty(int(count), dtype)
        cur[0] = p + nb
        return buf[p:p + nb].view(dtype)

    def step_device():
        last["res"] = job.run(d_clean.data_ptr(), d_N, device=True)

    def step_host():
        res = job.run(hC_np, hN_np)
        cur[0] = 0
        s = ctx.get_set(0, pinned_empty)
        g = ctx.get_globals(pinned_empty)
        last["d2h"] = sum(v.nbytes for v in s.values()) + sum(v.nbytes for v in g.values())
        last["set"], last["glob"], last["res"] = s, g, res

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    detail = {}

    def timed(fn, steps, sampler=None, tag=None):
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        phl = {p: [] for p in phases}
        if sampler:
            sampler.start()
        l0 = harc_b200.launch_count()
        m0 = ctx.last_ms("cudaMalloc_calls")
        ev[0].record(stream)
        if os.environ.get("HARCGPU_ALLOC_LOG"):
            sys.stderr.write("=== rank %d: timed region (%s) starts\n" % (rank, tag))
            sys.stderr.flush()
        for k in range(steps):
            fn()
            ev[k + 1].record(stream)
            for p in phases:
                phl[p].append(max(0.0, ctx.last_ms(p)))
        barrier()
        if sampler:
            sampler.stop.set()
        ms = ev[0].elapsed_time(ev[-1]) / steps
        per_step = [ev[k].elapsed_time(ev[k + 1]) for k in range(steps)]
        t = torch.tensor([ms] + [sum(phl[p]) / steps for p in phases], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)  # the slowest rank, phase by phase
        t = t.tolist()
        if tag:
            detail[tag] = {"per_step_ms_rank0": per_step, "phases_ms_per_step_rank0": phl,
                           "cudaMalloc_calls_in_timed_region_rank0": int(ctx.last_ms("cudaMalloc_calls") - m0)}
        return t[0], dict(zip(phases, t[1:])), (harc_b200.launch_count() - l0) // steps

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local)
    ms_dev, ph, launches = timed(step_device, args.steps, sampler, "device")
    ms_e2e = float("nan")
    if not args.no_e2e:
        step_host()
        ms_e2e, _, _ = timed(step_host, args.steps, None, "e2e")
    else:
        step_host()
    cnt = ctx.counters()       # the verify block and the counters are those of the LAST pass (the walk is not deterministic)
    m, s, u = ctx.reorder_counts()
    es = last["res"]["sizes"]
    laps = {k: ctx.last_ms("lap:" + k) for k in ("dict_keys_partition", "dict_barrier1", "dict_push", "dict_barrier2", "dict_shard_build", "dict_barrier3",
                                                "dict_bloom_pull", "s2_layout", "s2_consensus", "s2_pool", "s2_merge_emit", "s2_unaligned", "s2_sets")}
    laps = {k: v for k, v in laps.items() if v >= 0}

    # ---- verify (outside the timed region): every clean read exactly once over the order streams of all ranks, every read
    # in exactly one output, the parts add up
    n_total_clean = job.n_total
    ctx.trim()  # the library's cached blocks go back to the driver: the check below needs a few bytes per read of its own
    ptr, cnt_o = ctx.device_result("out_order")
    seen = torch.zeros(n_total_clean, dtype=torch.uint8, device="cuda")
    step = 1 << 26
    for o in range(0, cnt_o, step):
        k = min(step, cnt_o - o)
        ids = multi._dev_tensor(ptr + 4 * o, k, torch, "<i4").to(torch.int64) & 0xffffffff
        seen.index_add_(0, ids, torch.ones(k, dtype=torch.uint8, device="cuda"))
        del ids
    dist.all_reduce(seen, op=dist.ReduceOp.SUM)
    once = bool((seen == 1).all().item())
    del seen
    g = last["glob"]
    u_s = (4 * len(g["singleton"]) + len(g["singleton_tail"])) // L
    u_N = len(g["input_N"]) // (L + 1)
    stats = torch.tensor([m, s, u, len(last["set"]["pos"]), u_s, u_N, n_clean, n_N, int(es.aligned_singletons), int(es.aligned_N),
                          cnt["probes"], cnt["compares"], cnt["claim_fails"], cnt["harvested"], len(g["order"]), len(g["order_N"])],
                         dtype=torch.int64, device="cuda")
    allst = torch.empty(world * stats.numel(), dtype=torch.int64, device="cuda")
    dist.all_gather_into_tensor(allst, stats)
    allst = allst.view(world, -1).cpu().numpy()
    tot = allst.sum(axis=0)
    total_reads = int(tot[6] + tot[7])
    verify = {"every_clean_read_exactly_once": once,
              "matched_plus_singletons_equals_clean_reads": bool(tot[0] + tot[1] == n_total_clean),
              "reads_in_file_sets_plus_unaligned_equals_all_reads": bool(tot[3] + tot[4] + tot[5] == total_reads),
              "order_entries_equal_clean_reads": bool(tot[14] == n_total_clean), "order_N_entries_equal_N_reads": bool(tot[15] == tot[7]),
              "aligned_singletons_plus_unaligned": bool(tot[8] + tot[4] == tot[1]), "aligned_N_plus_unaligned": bool(tot[9] + tot[5] == tot[7]),
              "per_rank_matched": allst[:, 0].tolist()}
    verify["ok"] = all(v for k, v in verify.items() if isinstance(v, bool))

    # ---- the same workload on ONE GPU (rank 0, the others wait), same build, same run: the strong-scaling reference
    one_gpu = None
    if args.t1 and cfg["reads"] * 330 < 150e9:
        if rank == 0:
            try:
                wf = W.make(cfg["reads"], L, cfg["genome"], rc=cfg["rc"], errors=cfg["errors"], seed=args.seed, genome=genome, threads=ncpu(), keep_all=False)
                c1 = harc_b200.HarcGpu(L, device=local, walkers=args.walkers, file_sets=1, reads_per_walker=args.reads_per_walker, extend=args.extend)
                dC = torch.empty(wf["clean"].size + 16, dtype=torch.uint8, device="cuda")
                step = 1 << 30
                for o in range(0, wf["clean"].size, step):
                    chunk = wf["clean"][o:o + step]
                    dC[o:o + chunk.size].copy_(torch.from_numpy(chunk))
                dN = torch.empty(wf["withN"].size + 16, dtype=torch.uint8, device="cuda")
                dN[: wf["withN"].size].copy_(torch.from_numpy(wf["withN"]))
                st1 = torch.cuda.ExternalStream(c1.stream())
                ms1, ph1 = [], []
                for it in range(6):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    torch.cuda.synchronize()
                    e0.record(st1)
                    c1.load_reads_device(dC.data_ptr(), wf["n_clean"])
                    c1.build_dicts()
                    c1.reorder()
                    c1.load_pool_device(dN.data_ptr(), wf["n_N"])
                    c1.encode()
                    e1.record(st1)
                    torch.cuda.synchronize()
                    ms1.append(e0.elapsed_time(e1))
                    ph1.append({p: c1.last_ms(p) for p in phases})
                m1, s1, u1 = c1.reorder_counts()
                best = int(np.argmin(ms1[3:])) + 3
                one_gpu = {"ms_per_step": ms1[best], "Mreads_per_s": cfg["reads"] / ms1[best] / 1e3, "all_steps_ms": ms1,
                           "phases_ms": ph1[best], "chain_heads": u1, "singletons": s1,
                           "what": "the whole workload on rank 0's GPU alone (plain single-GPU path), best of 3 after 3 warm-ups, device-timed, "
                                   "measured in this run while the other ranks wait"}
                c1.close()
                del dC, dN, wf
            except Exception as ex:
                one_gpu = {"error": str(ex)[:300]}
        dist.barrier()

    # ---- secondary figure: the other split, one INDEPENDENT configs[1] read set per GPU (weak scaling, no data-path
    # collective), a few steps on a context of its own
    read_sets = None
    if args.read_sets_steps > 0:
        try:
            c1cfg = dict(W.CONFIGS[1])
            w1 = W.make(c1cfg["reads"], c1cfg["L"], c1cfg["genome"], rc=c1cfg["rc"], errors=c1cfg["errors"], seed=args.seed + 7919 * rank,
                        threads=thr, keep_all=False)
            dC1 = torch.empty(w1["clean"].size + 16, dtype=torch.uint8, device="cuda")
            dC1[: w1["clean"].size].copy_(torch.from_numpy(w1["clean"]))
            dN1 = torch.empty(w1["withN"].size + 16, dtype=torch.uint8, device="cuda")
            dN1[: w1["withN"].size].copy_(torch.from_numpy(w1["withN"]))
            c2 = harc_b200.HarcGpu(c1cfg["L"], device=local, walkers=args.walkers, file_sets=1, reads_per_walker=args.reads_per_walker, extend=args.extend)
            st2 = torch.cuda.ExternalStream(c2.stream())

            def rs_step():
                c2.load_reads_device(dC1.data_ptr(), w1["n_clean"])
                c2.build_dicts()
                c2.reorder()
                c2.load_pool_device(dN1.data_ptr(), w1["n_N"])
                c2.encode()
            for _ in range(3):
                rs_step()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st2)
            for _ in range(args.read_sets_steps):
                rs_step()
            e1.record(st2)
            barrier()
            t = torch.tensor([e0.elapsed_time(e1) / args.read_sets_steps], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            rs_ms = float(t.item())
            read_sets = {"value": world * c1cfg["reads"] / rs_ms / 1e3, "unit": "Mreads/s", "ms_per_step": rs_ms, "steps": args.read_sets_steps,
                         "scaling": "weak", "what": "one independent configs[1] read set (35 M reads) per GPU, no data-path collective; "
                                                    "reads of all ranks / slowest rank, device-timed"}
            c2.close()
            del dC1, dN1, w1
        except Exception as ex:
            read_sets = {"error": str(ex)[:300]}
        dist.barrier()

    value = total_reads / (ms_dev / 1000.0) / 1e6
    e2e = total_reads / (ms_e2e / 1000.0) / 1e6
    if rank == 0:
        n_loc_max = int(allst[:, 6].max())
        NWb = 8 * ctx.NW()
        comm_bytes = {
            "packed reads, stored into every replica by the pack kernel (per GPU, sent)": int(n_clean * NWb * (world - 1)),
            "(key, id) pairs of both dictionaries, pushed to the owners (per GPU, sent)": int(2 * n_clean * 12 * (world - 1) / world) if args.shard_dicts else 0,
            "Bloom filter segments copied from their owners (per GPU, received)": int(ctx.last_ms("job_bloom_bytes")) if args.shard_dicts else 0,
            "singleton ids all-gather (NCCL)": int(4 * tot[1]), "pool priorities all-reduce(min) (NCCL)": int(8 * (tot[1] + tot[7])),
            "walk": "dictionary probes (32 B loads) and claims (atomicAnd) over NVLink peer memory from inside the walk kernel",
        }
        out = {
            "metric": METRIC, "value": value, "unit": "Mreads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": cfg["name"] + ", ONE job on all GPUs", "workload_signature": workload_signature(cfg, args.seed),
                       "reads": total_reads, "clean_reads": n_total_clean, "reads_with_N": int(tot[7]), "reads_uploaded_per_gpu_max": n_loc_max,
                       "walkers": ctx.p.walkers or "auto", "file_sets": world,
                       "parallelism": "one job on %d GPUs: slice upload + pack with replication over NVLink, dictionaries %s, claim bitmap in NVLink "
                                      "peer memory, contigs by rank; NCCL for singleton ids / pool priorities"
                                      % (world, "sharded by key (all-to-all of (key, id) pairs, Bloom-filtered remote probes)" if args.shard_dicts else "replicated"),
                       "l2": "inputs (%.1f GB ASCII per GPU, %.1f GB packed reads in all) exceed the 126 MB L2; no explicit flush"
                             % (n_loc_max * (L + 1) / 1e9, n_total_clean * NWb / 1e9)},
            "phases_ms": ph, "phases_note": "max over ranks, phase by phase; waiting at the barriers between the GPUs is inside the phase that waits",
            "stage1": {"matched": int(tot[0]), "singletons": int(tot[1]), "chain_heads": int(tot[2]), "probes_per_read": float(tot[10]) / max(1, n_total_clean),
                       "compares_per_read": float(tot[11]) / max(1, n_total_clean), "claim_fails": int(tot[12]), "harvested": int(tot[13])},
            "stage2": {"aligned_singletons": int(tot[8]), "aligned_N": int(tot[9])},
            "roofline": roofline_block(cfg, n_total_clean / world, n_total_clean, total_reads / world, ph["walk"], ms_dev,
                                       "walk_kernel<%d, %d>, per GPU" % (ctx.NW(), args.lanes or 32)),
            "e2e": {"value": e2e, "unit": "Mreads/s", "ms_per_step": ms_e2e, "what": "one job at a time; every rank uploads its slice from pinned host "
                    "memory and reads back its file set", "h2d_bytes_per_step": int(h_clean.numel() + h_N.numel()) * world, "d2h_bytes_per_step": int(last.get("d2h", 0)) * world},
            "exchange_bytes_per_step": comm_bytes, "verify": verify, "one_gpu_same_workload": one_gpu, "read_sets": read_sets,
            "scaling_note": "strong scaling of ONE %d-read job; `bench.py --gpus 1` runs configs[1] (the metric's configuration, a different "
                            "workload), so the one-GPU time of THIS workload is measured here, on rank 0's GPU in the same run "
                            "(one_gpu_same_workload)" % total_reads, "laps_ms_rank0": laps or None, "detail": detail,
            "gpu_launches": int(launches), "allocator": {"device_peak_MB": ctx.last_ms("peak_MB"), "cudaMalloc_calls": ctx.last_ms("cudaMalloc_calls")},
            "clocks": sampler.summary(), "commit": git_head(),
        }
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    dist.barrier()
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=None, choices=[0, 1, 2, 3, 4],
                    help="BASELINE.json configs[k]; default 1 on one GPU / read sets, 2 for one job on several GPUs")
    ap.add_argument("--reads", type=float, default=None, help="override the config's read count")
    ap.add_argument("--genome", type=float, default=None)
    ap.add_argument("--rc", type=int, default=None)
    ap.add_argument("--errors", type=int, default=None)
    ap.add_argument("--walkers", type=int, default=0)
    ap.add_argument("--file-sets", type=int, default=1)
    ap.add_argument("--reads-per-walker", type=int, default=0)
    ap.add_argument("--extend", type=int, default=0)
    ap.add_argument("--lanes", type=int, default=0, help="lanes per walker (16 or 32; 0 = default)")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--mode", default="one-job", choices=["one-job", "read-sets", "single-job"],
                    help="N>1: one-job = ONE read set on all ranks (strong scaling, default); read-sets = every rank compresses its own "
                         "read set (weak scaling, no data-path collective)")
    ap.add_argument("--shard-dicts", type=int, default=1, help="one-job mode: 1 = dictionaries sharded by key over the GPUs, 0 = replicated")
    ap.add_argument("--read-sets-steps", type=int, default=5,
                    help="one-job mode: timed steps of the secondary read-sets figure (one independent configs[1] set per GPU); 0 = skip")
    ap.add_argument("--t1", type=int, default=1, help="one-job mode: also time the whole workload on rank 0's GPU alone (outside the timed region)")
    ap.add_argument("--pipeline", type=int, default=3, help="e2e: jobs in flight (contexts) for the pipelined figure; 1 = off")
    ap.add_argument("--ref-budget-s", type=float, default=600.0, help="--impl reference: wall-time bound of the whole run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--force-cpu-baseline", action="store_true", help="run the reference on the host even for the large configs")
    ap.add_argument("--no-bits", action="store_true", help="skip the bits/base block (stage III stand-in over both archives)")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the host-buffer pass")
    ap.add_argument("--ingest-reads", type=float, default=8e6,
                    help="reads of the workload that are also laid out as a FASTQ file to time the fused ingest (0 = skip)")
    args = ap.parse_args()
    if args.mode == "single-job":
        args.mode = "one-job"

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank == 0:
            run_reference(args, world)
        return

    # Rank 0 prints exactly ONE line on stdout, the JSON.  Everything else that writes to file descriptor 1 (NCCL's version
    # banner, library chatter) is sent to stderr: fd 1 is pointed at stderr for the run and the JSON goes to the saved fd.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if world > 1 and args.mode == "one-job":
        run_one_job(args, rank, world, local, dist, json_fd)
    else:
        run_single(args, rank, world, local, dist, json_fd)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
