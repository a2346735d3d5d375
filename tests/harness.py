"""Shared helpers of the test suite: synthetic datasets and the oracle (checker) bindings."""
import ctypes
import filecmp
import os
import shutil

import numpy as np

import refrun as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGE1_FILES = ["temp.dna", "temp.dna.singleton", "read_rev.txt", "tempflag.txt", "temppos.txt", "read_order.bin",
                "read_order.bin.singleton"]


class OParams(ctypes.Structure):
    _fields_ = [("readlen", ctypes.c_int), ("maxmatch", ctypes.c_int), ("thresh", ctypes.c_int), ("thresh_s", ctypes.c_int),
                ("numdict", ctypes.c_int), ("maxsearch", ctypes.c_int), ("dict_start", ctypes.c_int * 2),
                ("dict_end", ctypes.c_int * 2)]


_olib = None


def oracle():
    global _olib
    if _olib is None:
        _olib = ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
        _olib.oracle_reorder_dir.restype = ctypes.c_int64
        _olib.oracle_reorder_dir.argtypes = [ctypes.c_char_p, ctypes.POINTER(OParams), ctypes.c_int]
        _olib.oracle_encode_dir.argtypes = [ctypes.c_char_p, ctypes.POINTER(OParams), ctypes.c_int, ctypes.c_void_p]
    return _olib


def oparams(L):
    p = OParams()
    oracle().oracle_default_params(L, ctypes.byref(p))
    return p


def oracle_reorder(basedir, L, walkers=1):
    r = oracle().oracle_reorder_dir(basedir.encode(), ctypes.byref(oparams(L)), walkers)
    assert r >= 0, r
    return r


def oracle_encode(basedir, L, sets=1):
    al = (ctypes.c_uint32 * 2)()
    r = oracle().oracle_encode_dir(basedir.encode(), ctypes.byref(oparams(L)), sets, al)
    assert r == 0, r
    return al[0], al[1]


def make_dataset(root, name, n, L, genome, rc=False, errors=False, seed=1):
    """genome FASTA -> gen_fastq[_noRC] -> preprocess: <root>/<name>/output/{input_clean.dna,input_N.dna,...}."""
    d = os.path.join(root, name)
    if os.path.exists(os.path.join(d, "output", "numreads.bin")):
        return d
    os.makedirs(d, exist_ok=True)
    R.make_genome(os.path.join(d, "g.fa"), genome, seed=seed)
    R.gen_fastq(os.path.join(d, "g.fa"), os.path.join(d, "r.fastq"), n, L, rc=rc, errors=errors)
    R.preprocess(os.path.join(d, "r.fastq"), d, L)
    return d


def make_repeat_dataset(root, name, n, L, rc=True, errors=True, seed=1):
    """Like make_dataset on refrun.repeat_rich_genome: poly-A, tandem repeats, duplications -> bins beyond maxsearch."""
    d = os.path.join(root, name)
    if os.path.exists(os.path.join(d, "output", "numreads.bin")):
        return d
    os.makedirs(d, exist_ok=True)
    R.make_genome(os.path.join(d, "g.fa"), 0, seed=seed, bases=R.repeat_rich_genome(seed))
    R.gen_fastq(os.path.join(d, "g.fa"), os.path.join(d, "r.fastq"), n, L, rc=rc, errors=errors)
    R.preprocess(os.path.join(d, "r.fastq"), d, L)
    return d


def dataset(root, case, seed):
    """case = (name, reads, L, genome, rc, errors); genome == 'repeats' -> the repeat-rich genome."""
    name, n, L, G, rc, err = case
    if G == "repeats":
        return make_repeat_dataset(root, name, n, L, rc, err, seed)
    return make_dataset(root, name, n, L, G, rc, err, seed)


def clone(src, dst, names=None):
    shutil.rmtree(dst, ignore_errors=True)
    return R.copy_stage(src, dst, names)


def same_files(a, b, names):
    bad = [f for f in names if not filecmp.cmp(os.path.join(a, "output", f), os.path.join(b, "output", f), shallow=False)]
    return bad


def read_lines(path, L):
    a = np.fromfile(path, dtype=np.uint8)
    return a.reshape(-1, L + 1)
