#!/usr/bin/env python
"""One pass of the hot path over a synthetic read set of any shape (read length, coverage, errors, reverse complements),
with the invariants that do not need the reference: every clean read exactly once in the order streams, counts add up.
Usage: run_shape.py <reads> <readlen> <genome> [rc] [errors] [file_sets]
<genome> = a length (uniform i.i.d. bases) or `repeats:<scale>`: oracle/refrun.repeat_rich_genome scaled up <scale> times
(poly-A run, tandem repeat, exact and diverged duplications: dictionary bins of 10^3..10^5 reads)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import harc_b200
import workload as W


def main():
    n, L = int(float(sys.argv[1])), int(sys.argv[2])
    genome = None
    if sys.argv[3].startswith("repeats:"):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import refrun as R
        k = int(sys.argv[3].split(":")[1])
        g = R.repeat_rich_genome(3, unique=300000 * k, polyA=80000 * k, tandem_unit=7, tandem_copies=2000 * k, dup_len=300, dup_copies=60 * k,
                                 div_len=1000, div_copies=20 * k)
        genome = np.concatenate([g, np.full(1, ord("A"), np.uint8)])
        G = len(g)
    else:
        G = int(float(sys.argv[3]))
    rc = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    err = int(sys.argv[5]) if len(sys.argv) > 5 else 1
    K = int(sys.argv[6]) if len(sys.argv) > 6 else 1
    w = W.make(n, L, G, rc=bool(rc), errors=bool(err), seed=3, genome=genome)
    ctx = harc_b200.HarcGpu(L, file_sets=K)
    for it in range(3):
        m0 = ctx.last_ms("cudaMalloc_calls")
        t0 = time.perf_counter()
        ctx.load_reads(w["clean"], w["n_clean"])
        ctx.build_dicts()
        m, s, u = ctx.reorder()
        ctx.load_pool(None, None, w["withN"])
        es = ctx.encode()
        t1 = time.perf_counter()
    ph = {p: round(ctx.last_ms(p), 2) for p in ("pack", "dict", "walk", "finalize", "pooldict", "encode")}
    g = ctx.get_globals()
    order = np.sort(g["order"])
    assert m + s == w["n_clean"], (m, s, w["n_clean"])
    assert np.array_equal(order, np.arange(w["n_clean"], dtype=np.uint32)), "every clean read exactly once"
    assert np.array_equal(np.sort(g["order_N"]), np.arange(w["n_N"], dtype=np.uint32)), "every read with N exactly once"
    dev = sum(ph.values())
    print({"reads": n, "L": L, "genome": G, "rc": rc, "errors": err, "clean": w["n_clean"], "with_N": w["n_N"], "matched": m, "singletons": s,
           "chain_heads": u, "aligned_singletons": int(es.aligned_singletons), "aligned_N": int(es.aligned_N), "phases_ms": ph,
           "device_ms": round(dev, 1), "Mreads_per_s_device": round(n / dev / 1e3, 1), "host_wall_ms_incl_copies": round(1000 * (t1 - t0), 1),
           "counters": ctx.counters(), "pool_passes": ctx.last_ms("pool_passes"),
           "laps": {k: round(ctx.last_ms("lap:" + k), 2) for k in ("bd1_keys", "bd1_sort", "bd1_csr", "bd1_place", "s2_layout", "s2_consensus", "s2_pool",
                                                                    "s2_merge_emit", "s2_unaligned", "s2_sets") if ctx.last_ms("lap:" + k) >= 0},
           "walkers": "auto", "driver_allocations_last_pass": int(ctx.last_ms("cudaMalloc_calls") - m0)})
    ctx.close()


if __name__ == "__main__":
    main()
