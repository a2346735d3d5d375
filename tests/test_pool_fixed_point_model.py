"""CPU: a small model of stage II's pool re-alignment for bins beyond maxsearch.

The reference (encoder.cpp:231-418) works through its windows one after the other; a window scans its bin from the tail over
at most `maxsearch` reads that are still LIVE and takes every one that matches, and a taken read leaves both of its bins
at once (encoder.cpp:293, 1010-1031).  The GPU probes all windows at the same time: a read counts as live for a window
unless a window of higher priority (= earlier in the reference's order) holds it at that moment, every match lowers the
read's priority word with atomicMin, and the probe is repeated until nothing moves.  This test checks on random instances
that the fixed point of that iteration is exactly the sequential result -- in any order of the windows inside a pass, which
is what a parallel launch amounts to.  (The kernel itself is checked against the reference on a repeat-rich input by
tests/test_stage2_gpu.py; this model is the argument behind it, made executable.)"""
import numpy as np
import pytest

INF = 1 << 60


def _instance(rng, n_reads, n_bins, n_windows):
    # every read sits in one bin of dictionary 0 and one of dictionary 1; a bin lists its reads in ascending id (= scan from the tail)
    bins = [[] for _ in range(2 * n_bins)]
    for r in range(n_reads):
        bins[int(rng.integers(n_bins))].append(r)
        bins[n_bins + int(rng.integers(n_bins))].append(r)
    # a window = one probe of one bin; priorities are the window numbers (the reference's order)
    win_bin = rng.integers(0, 2 * n_bins, size=n_windows)
    match = rng.random((n_windows, n_reads)) < 0.35
    return bins, win_bin, match


def _sequential(bins, win_bin, match, maxsearch):
    n_reads = match.shape[1]
    owner = [INF] * n_reads
    for w, b in enumerate(win_bin):
        live = 0
        for r in reversed(bins[b]):
            if owner[r] != INF:
                continue            # removed from the bin when it was taken
            if live >= maxsearch:
                break
            live += 1
            if match[w, r]:
                owner[r] = w
    return owner


def _parallel(bins, win_bin, match, maxsearch, rng):
    n_reads = match.shape[1]
    best = [INF] * n_reads
    passes = 0
    while True:
        passes += 1
        changed = False
        for w in rng.permutation(len(win_bin)):     # any order: the windows of a pass run side by side
            w = int(w)
            live = 0
            for r in reversed(bins[win_bin[w]]):
                if best[r] < w:
                    continue        # held by an earlier window: not in the bin any more for this one
                if live >= maxsearch:
                    break
                live += 1
                if match[w, r] and w < best[r]:
                    best[r] = w
                    changed = True
        if not changed:
            return best, passes
        assert passes < 1000


@pytest.mark.parametrize("seed", range(12))
def test_fixed_point_of_the_parallel_probe_is_the_sequential_result(seed):
    rng = np.random.default_rng(seed)
    n_reads, n_bins, n_windows = int(rng.integers(20, 120)), int(rng.integers(1, 5)), int(rng.integers(5, 60))
    maxsearch = int(rng.integers(1, 6))             # small, so that most bins are "beyond maxsearch"
    bins, win_bin, match = _instance(rng, n_reads, n_bins, n_windows)
    want = _sequential(bins, win_bin, match, maxsearch)
    got, passes = _parallel(bins, win_bin, match, maxsearch, rng)
    assert got == want, (seed, passes)


def test_one_pass_is_enough_when_no_bin_exceeds_maxsearch():
    rng = np.random.default_rng(99)
    bins, win_bin, match = _instance(rng, 60, 8, 40)
    maxsearch = max(len(b) for b in bins)           # nothing is ever cut short
    want = _sequential(bins, win_bin, match, maxsearch)
    got, passes = _parallel(bins, win_bin, match, maxsearch, rng)
    assert got == want and passes == 2              # the second pass only confirms that nothing moves
