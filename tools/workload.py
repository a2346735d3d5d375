"""Synthetic workloads for bench.py and the large-size tests (ctypes over tools/simreads.c).  Not the product, not
the oracle: it only manufactures inputs with the read model of the reference's util/gen_fastq*."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libsimreads.so")


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
        _lib.sim_reads.restype = ctypes.c_uint64
        _lib.sim_genome.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64]
        _lib.sim_reads.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                   ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p]
        _lib.sim_split.argtypes = [ctypes.c_void_p] * 2 + [ctypes.c_uint64, ctypes.c_int] + [ctypes.c_void_p] * 3
    return _lib


def make(nreads, L, genome_len, rc=False, errors=False, seed=1):
    """Returns dict(clean=uint8[n_clean*(L+1)], withN=uint8[n_N*(L+1)], order_N=uint32[n_N], all=uint8[n*(L+1)])."""
    lib = _load()
    g = np.empty(genome_len + 1, dtype=np.uint8)
    lib.sim_genome(g.ctypes.data, genome_len, seed)
    allr = np.empty(nreads * (L + 1), dtype=np.uint8)
    hasN = np.empty(nreads, dtype=np.uint8)
    nN = int(lib.sim_reads(g.ctypes.data, genome_len, nreads, L, int(rc), int(errors), seed + 1000003, allr.ctypes.data, hasN.ctypes.data))
    clean = np.empty((nreads - nN) * (L + 1), dtype=np.uint8)
    withN = np.empty(nN * (L + 1), dtype=np.uint8)
    order_N = np.empty(nN, dtype=np.uint32)
    lib.sim_split(allr.ctypes.data, hasN.ctypes.data, nreads, L, clean.ctypes.data, withN.ctypes.data, order_N.ctypes.data)
    return dict(clean=clean, withN=withN, order_N=order_N, all=allr, n=nreads, n_clean=nreads - nN, n_N=nN, L=L)


def write_dir(w, basedir):
    """Lay the workload out as preprocess.cpp would have (harc:50): <basedir>/output/{input_clean.dna,...}."""
    out = os.path.join(basedir, "output")
    os.makedirs(out, exist_ok=True)
    w["clean"].tofile(os.path.join(out, "input_clean.dna"))
    w["withN"].tofile(os.path.join(out, "input_N.dna"))
    w["order_N"].tofile(os.path.join(out, "read_order_N.bin"))
    np.array([w["n_clean"]], dtype=np.uint32).tofile(os.path.join(out, "numreads.bin"))
    return basedir
