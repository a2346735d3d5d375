"""CPU: the synthetic workload generator (tools/simreads.c through tools/workload.py) names a workload by (seed, sizes)
alone: the same reads whatever the thread count, and a slice made on its own is that slice of the whole (one job on
several GPUs: every rank makes only its part)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import workload as W


def test_reads_do_not_depend_on_the_thread_count():
    n, L, G = 100_000, 100, 500_000
    a = W.make(n, L, G, rc=True, errors=True, seed=5, threads=1)
    b = W.make(n, L, G, rc=True, errors=True, seed=5, threads=7)
    assert np.array_equal(a["all"], b["all"]) and np.array_equal(a["clean"], b["clean"]) and np.array_equal(a["order_N"], b["order_N"])
    c = W.make(n, L, G, rc=True, errors=True, seed=6, threads=7)
    assert not np.array_equal(a["all"], c["all"])
    lines = a["all"].reshape(n, L + 1)
    assert np.all(lines[:, L] == 10) and set(np.unique(lines[:, :L]).tolist()) <= set(b"ACGTN")
    assert a["n_clean"] + a["n_N"] == n and 0.15 < a["n_N"] / n < 0.30   # ~22 % of 100-bp reads carry an N


def test_slices_are_slices_of_the_whole():
    n, L, G, world = 150_000, 100, 400_000, 4
    whole = W.make(n, L, G, rc=True, errors=True, seed=9)
    genome = W.make_genome(G, 9)
    cuts = W.slice_bounds(n, world)
    assert cuts[0][0] == 0 and cuts[-1][1] == n and all(x[1] == y[0] for x, y in zip(cuts, cuts[1:]))
    assert all(a % W.read_block() == 0 for a, _ in cuts)
    parts = [W.make(n, L, G, rc=True, errors=True, seed=9, first=a, count=b - a, genome=genome) for a, b in cuts]
    assert np.array_equal(np.concatenate([p["all"] for p in parts]), whole["all"])
    assert np.array_equal(np.concatenate([p["clean"] for p in parts]), whole["clean"])
    assert np.array_equal(np.concatenate([p["order_N"] for p in parts]), whole["order_N"])   # global read numbers
    for k, cfg in W.CONFIGS.items():
        assert cfg["reads"] > 0 and cfg["L"] in (100, 250) and cfg["genome"] > cfg["L"]
