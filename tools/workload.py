"""Synthetic workloads for bench.py and the large-size tests (ctypes over tools/simreads.c).  Not the product, not
the oracle: it only manufactures inputs with the read model of the reference's util/gen_fastq*.  The random stream is
seeded per fixed-size block, so (seed, sizes) names the workload on any machine, and a slice of it can be made alone."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libsimreads.so")

# BASELINE.json configs: reads, read length, genome, reverse complements, errors, order-preserving
CONFIGS = {
    0: dict(reads=1_000_000, L=100, genome=10_000_000, rc=False, errors=False, preserve=False,
            name="configs[0]: 1M error-free 100-bp reads (gen_fastq_noRC model), 10 Mbp genome"),
    1: dict(reads=35_000_000, L=100, genome=50_000_000, rc=False, errors=True, preserve=False,
            name="configs[1]: 35M x 100bp reads, 1% substitutions incl. N (gen_fastq_noRC -e model), 50 Mbp genome"),
    2: dict(reads=200_000_000, L=100, genome=1_000_000_000, rc=True, errors=False, preserve=False,
            name="configs[2]: 200M x 100bp error-free reads with reverse complements (gen_fastq model), 1 Gbp genome"),
    3: dict(reads=100_000_000, L=250, genome=1_000_000_000, rc=False, errors=True, preserve=True,
            name="configs[3]: 100M x 250bp reads, 1% errors incl. N (gen_fastq_noRC -e model), 1 Gbp genome (25x; BASELINE.json "
                 "leaves the genome open), order-preserving -p"),
    4: dict(reads=1_000_000_000, L=100, genome=3_000_000_000, rc=True, errors=False, preserve=False,
            name="configs[4]: 1B x 100bp reads (gen_fastq model), 3 Gbp genome"),
}


def build():
    src = os.path.join(HERE, "simreads.c")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        subprocess.check_call([cc, "-O2", "-fopenmp", "-shared", "-fPIC", "-o", LIB, src, "-lm"])
    return LIB


_lib = None


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB)
        u64, vp, ci = ctypes.c_uint64, ctypes.c_void_p, ctypes.c_int
        _lib.sim_reads.restype = u64
        _lib.sim_read_block.restype = ctypes.c_uint32
        _lib.sim_genome.argtypes = [vp, u64, u64, ci]
        _lib.sim_reads.argtypes = [vp, u64, u64, u64, ci, ci, ci, u64, vp, vp, ci]
        _lib.sim_split.argtypes = [vp, vp, u64, ci, vp, vp, vp, u64]
    return _lib


def read_block():
    """Granularity of slices: `first` of make() must be a multiple of this."""
    return int(_load().sim_read_block())


def slice_bounds(nreads, parts):
    """Cut [0, nreads) into `parts` contiguous ranges on block boundaries (the last one takes the remainder)."""
    b = read_block()
    per = (nreads // parts + b - 1) // b * b
    cuts = [min(nreads, k * per) for k in range(parts)] + [nreads]
    return [(cuts[k], cuts[k + 1]) for k in range(parts)]


def make_genome(genome_len, seed=1, threads=0):
    g = np.empty(genome_len + 1, dtype=np.uint8)
    _load().sim_genome(g.ctypes.data, genome_len, seed, threads)
    return g


def make(nreads, L, genome_len, rc=False, errors=False, seed=1, first=0, count=None, genome=None, threads=0, keep_all=True):
    """Reads [first, first + count) of the workload (nreads, L, genome_len, rc, errors, seed).
    Returns dict(clean=uint8[n_clean*(L+1)], withN=uint8[n_N*(L+1)], order_N=uint32[n_N], all=uint8[count*(L+1)], ...)."""
    lib = _load()
    if count is None:
        count = nreads - first
    if threads <= 0:
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    g = genome if genome is not None else make_genome(genome_len, seed, threads)
    allr = np.empty(count * (L + 1), dtype=np.uint8)
    hasN = np.empty(count, dtype=np.uint8)
    nN = int(lib.sim_reads(g.ctypes.data, genome_len, first, count, L, int(rc), int(errors), seed + 1000003, allr.ctypes.data,
                           hasN.ctypes.data, threads))
    assert nN <= count, "first must be a multiple of read_block()"
    clean = np.empty((count - nN) * (L + 1), dtype=np.uint8)
    withN = np.empty(nN * (L + 1), dtype=np.uint8)
    order_N = np.empty(nN, dtype=np.uint32)
    lib.sim_split(allr.ctypes.data, hasN.ctypes.data, count, L, clean.ctypes.data, withN.ctypes.data, order_N.ctypes.data, first)
    return dict(clean=clean, withN=withN, order_N=order_N, all=allr if keep_all else None, n=count, n_clean=count - nN, n_N=nN, L=L,
                first=first)


def write_dir(w, basedir):
    """Lay the workload out as preprocess.cpp would have (harc:50): <basedir>/output/{input_clean.dna,...}."""
    out = os.path.join(basedir, "output")
    os.makedirs(out, exist_ok=True)
    w["clean"].tofile(os.path.join(out, "input_clean.dna"))
    w["withN"].tofile(os.path.join(out, "input_N.dna"))
    w["order_N"].tofile(os.path.join(out, "read_order_N.bin"))
    np.array([w["n_clean"]], dtype=np.uint32).tofile(os.path.join(out, "numreads.bin"))
    return basedir
