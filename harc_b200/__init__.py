"""harc_b200 -- ctypes binding of libharcgpu.so (include/harcgpu.h) for the tests and bench.py.

The product is the C-ABI library and the two drop-in executables (`reorder.out <basedir>`, `encoder.out <basedir>`,
the reference's process contract of harc:65-69).  This module only loads the library and mirrors its calls; it
fails loudly when the CUDA extension is missing -- there is no CPU path.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.path.join(HERE, "libharcgpu.so")

EXPORTS = [
    "harcgpu_default_params", "harcgpu_create", "harcgpu_destroy", "harcgpu_last_error", "harcgpu_device_count",
    "harcgpu_load_reads", "harcgpu_load_reads_device", "harcgpu_build_dicts", "harcgpu_dump_dict", "harcgpu_reorder",
    "harcgpu_reorder_counts", "harcgpu_get_reorder", "harcgpu_get_reordered_reads", "harcgpu_get_counters",
    "harcgpu_set_stream", "harcgpu_load_pool", "harcgpu_encode", "harcgpu_get_encode_sizes", "harcgpu_get_set_sizes",
    "harcgpu_get_set", "harcgpu_get_globals", "harcgpu_reorder_dir", "harcgpu_encode_dir", "harcgpu_last_ms", "harcgpu_stream",
    "harcgpu_load_pool_device", "harcgpu_launch_count", "harcgpu_stage_nreads",
    "harcgpu_job_init", "harcgpu_job_connect", "harcgpu_job_load_reads", "harcgpu_job_load_reads_device", "harcgpu_job_build_dicts",
    "harcgpu_job_reorder", "harcgpu_job_set_barrier", "harcgpu_set_pool_exchange", "harcgpu_load_pool_ids", "harcgpu_device_result",
    "harcgpu_get_packed_order", "harcgpu_trim",
    "harcgpu_debug_sort", "harcgpu_fastq_readlen", "harcgpu_ingest_fastq", "harcgpu_ingest_fastq_device", "harcgpu_get_ingest", "harcgpu_load_pool_ingested",
]


class Params(ctypes.Structure):
    """harcgpu_params: the macros of the reference's generated src/config.h (harc:52-63) + walkers/file_sets."""
    _fields_ = [("readlen", ctypes.c_int), ("maxmatch", ctypes.c_int), ("thresh", ctypes.c_int), ("thresh_s", ctypes.c_int),
                ("numdict", ctypes.c_int), ("maxsearch", ctypes.c_int), ("dict_start", ctypes.c_int * 2),
                ("dict_end", ctypes.c_int * 2), ("walkers", ctypes.c_int), ("file_sets", ctypes.c_int),
                ("reads_per_walker", ctypes.c_int), ("extend", ctypes.c_int), ("lanes_per_walker", ctypes.c_int),
                ("shard_dicts", ctypes.c_int)]


class EncodeSizes(ctypes.Structure):
    _fields_ = [("n_order", ctypes.c_uint32), ("n_order_N", ctypes.c_uint32), ("singleton_bytes", ctypes.c_uint64),
                ("singleton_tail", ctypes.c_uint64), ("input_N_bytes", ctypes.c_uint64),
                ("aligned_singletons", ctypes.c_uint32), ("aligned_N", ctypes.c_uint32)]


class SetSizes(ctypes.Structure):
    _fields_ = [("seq_bytes", ctypes.c_uint64), ("seq_tail", ctypes.c_uint64), ("pos_bytes", ctypes.c_uint64),
                ("noise_bytes", ctypes.c_uint64), ("noisepos_bytes", ctypes.c_uint64), ("rev_bytes", ctypes.c_uint64),
                ("rev_tail", ctypes.c_uint64)]


class IngestInfo(ctypes.Structure):
    _fields_ = [("readlen", ctypes.c_uint32), ("total_reads", ctypes.c_uint64), ("n_clean", ctypes.c_uint32), ("n_N", ctypes.c_uint32)]


class Counters(ctypes.Structure):
    _fields_ = [("steps", ctypes.c_uint64), ("probes", ctypes.c_uint64), ("key_hits", ctypes.c_uint64),
                ("compares", ctypes.c_uint64), ("claim_fails", ctypes.c_uint64), ("restarts", ctypes.c_uint64),
                ("harvested", ctypes.c_uint64)]


_lib = None
# hook of harcgpu_set_pool_exchange: int fn(void *user, void *d_best, uint64_t count)
POOL_EXCHANGE = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64)
JOB_BARRIER = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p)


def load_library():
    """Load libharcgpu.so from the tree.  Raises if it has not been built: the product has no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIBPATH):
        raise RuntimeError("libharcgpu.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                           "harc_b200 has no CPU fallback")
    lib = ctypes.CDLL(LIBPATH)
    lib.harcgpu_last_error.restype = ctypes.c_char_p
    lib.harcgpu_last_ms.restype = ctypes.c_double
    lib.harcgpu_last_ms.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
    lib.harcgpu_stream.restype = ctypes.c_void_p
    lib.harcgpu_stream.argtypes = [ctypes.c_void_p]
    lib.harcgpu_create.argtypes = [ctypes.c_int, ctypes.POINTER(Params), ctypes.POINTER(ctypes.c_void_p)]
    lib.harcgpu_destroy.argtypes = [ctypes.c_void_p]
    vp, u32, cp = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_char_p
    lib.harcgpu_load_reads.argtypes = [vp, vp, u32]
    lib.harcgpu_load_reads_device.argtypes = [vp, vp, u32]
    lib.harcgpu_build_dicts.argtypes = [vp]
    lib.harcgpu_dump_dict.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, vp, vp, vp, vp]
    lib.harcgpu_reorder.argtypes = [vp]
    lib.harcgpu_reorder_counts.argtypes = [vp, vp, vp, vp]
    lib.harcgpu_get_reorder.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.harcgpu_get_reordered_reads.argtypes = [vp, vp, vp]
    lib.harcgpu_get_counters.argtypes = [vp, ctypes.POINTER(Counters)]
    lib.harcgpu_set_stream.argtypes = [vp, vp, vp, vp, vp, vp, u32]
    lib.harcgpu_load_pool.argtypes = [vp, vp, vp, u32, vp, u32]
    lib.harcgpu_encode.argtypes = [vp]
    lib.harcgpu_load_pool_device.argtypes = [vp, vp, u32]
    lib.harcgpu_stage_nreads.argtypes = [vp, vp, u32]
    lib.harcgpu_job_init.argtypes = [vp, ctypes.c_int, ctypes.c_int, u32, u32, u32, vp, ctypes.POINTER(vp)]
    lib.harcgpu_job_connect.argtypes = [vp, vp, ctypes.POINTER(vp)]
    lib.harcgpu_job_set_barrier.argtypes = [vp, JOB_BARRIER, vp]
    lib.harcgpu_job_load_reads.argtypes = [vp, vp, u32]
    lib.harcgpu_job_load_reads_device.argtypes = [vp, vp, u32]
    lib.harcgpu_job_build_dicts.argtypes = [vp]
    lib.harcgpu_job_reorder.argtypes = [vp]
    lib.harcgpu_device_result.argtypes = [vp, cp, ctypes.POINTER(vp), ctypes.POINTER(ctypes.c_uint64)]
    lib.harcgpu_set_pool_exchange.argtypes = [vp, POOL_EXCHANGE, vp]
    lib.harcgpu_load_pool_ids.argtypes = [vp, vp, u32, vp, u32]
    lib.harcgpu_get_packed_order.argtypes = [vp, vp, vp, vp, vp]
    lib.harcgpu_trim.argtypes = [vp]
    lib.harcgpu_launch_count.restype = ctypes.c_uint64
    lib.harcgpu_debug_sort.argtypes = [vp, vp, vp, ctypes.c_uint64, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    lib.harcgpu_fastq_readlen.argtypes = [vp, ctypes.c_uint64]
    lib.harcgpu_ingest_fastq.argtypes = [vp, vp, ctypes.c_uint64, ctypes.POINTER(IngestInfo)]
    lib.harcgpu_ingest_fastq_device.argtypes = [vp, vp, ctypes.c_uint64, ctypes.POINTER(IngestInfo)]
    lib.harcgpu_get_ingest.argtypes = [vp, vp, vp, vp]
    lib.harcgpu_load_pool_ingested.argtypes = [vp]
    lib.harcgpu_get_encode_sizes.argtypes = [vp, ctypes.POINTER(EncodeSizes)]
    lib.harcgpu_get_set_sizes.argtypes = [vp, ctypes.c_int, ctypes.POINTER(SetSizes)]
    lib.harcgpu_get_set.argtypes = [vp, ctypes.c_int, vp, vp, vp, vp, vp, vp, vp]
    lib.harcgpu_get_globals.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.harcgpu_reorder_dir.argtypes = [vp, cp]
    lib.harcgpu_encode_dir.argtypes = [vp, cp]
    _lib = lib
    return lib


def fastq_readlen(fastq):
    """harc:44: length of the second line."""
    a = np.frombuffer(fastq, dtype=np.uint8) if isinstance(fastq, (bytes, bytearray)) else fastq
    return int(load_library().harcgpu_fastq_readlen(_ptr(a) if a.size else None, a.size))


def launch_count():
    return int(load_library().harcgpu_launch_count())


class HarcError(RuntimeError):
    pass


def default_params(readlen, walkers=0, file_sets=1, reads_per_walker=0, extend=0, lanes_per_walker=0, shard_dicts=0):
    lib = load_library()
    p = Params()
    if lib.harcgpu_default_params(int(readlen), ctypes.byref(p)):
        raise HarcError(lib.harcgpu_last_error().decode())
    p.walkers = walkers
    p.file_sets = file_sets
    p.reads_per_walker = reads_per_walker
    p.extend = extend
    p.lanes_per_walker = lanes_per_walker
    p.shard_dicts = shard_dicts
    return p


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, (bytes, bytearray)):
        return ctypes.cast(ctypes.c_char_p(bytes(a)), ctypes.c_void_p)
    if isinstance(a, int):
        return ctypes.c_void_p(a)
    assert a.flags["C_CONTIGUOUS"]
    return ctypes.c_void_p(a.ctypes.data)


class HarcGpu:
    """One context on one GPU.  Method names follow include/harcgpu.h."""

    def __init__(self, readlen=None, device=0, params=None, walkers=0, file_sets=1, reads_per_walker=0, extend=0,
                 lanes_per_walker=0, shard_dicts=0):
        self.lib = load_library()
        self.p = params if params is not None else default_params(readlen, walkers, file_sets, reads_per_walker, extend, lanes_per_walker,
                                                                  shard_dicts)
        self.L = self.p.readlen
        h = ctypes.c_void_p()
        self._ck(self.lib.harcgpu_create(device, ctypes.byref(self.p), ctypes.byref(h)))
        self.h = h

    def NW(self):
        """64-bit words of a packed read."""
        return (2 * self.L + 63) // 64

    def _ck(self, rc):
        if rc != 0:
            raise HarcError(self.lib.harcgpu_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.harcgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- fused ingest (preprocess.cpp)
    def ingest_fastq(self, fastq):
        """fastq: bytes / uint8 array with the whole FASTQ file.  Returns the harcgpu_ingest_info fields as a dict."""
        a = np.frombuffer(fastq, dtype=np.uint8) if isinstance(fastq, (bytes, bytearray)) else fastq
        self._keep = a
        info = IngestInfo()
        self._ck(self.lib.harcgpu_ingest_fastq(self.h, _ptr(a) if a.size else None, a.size, ctypes.byref(info)))
        self._ingest = {k: int(getattr(info, k)) for k, _ in IngestInfo._fields_}
        return dict(self._ingest)

    def ingest_fastq_device(self, dptr, nbytes):
        info = IngestInfo()
        self._ck(self.lib.harcgpu_ingest_fastq_device(self.h, ctypes.c_void_p(dptr), nbytes, ctypes.byref(info)))
        self._ingest = {k: int(getattr(info, k)) for k, _ in IngestInfo._fields_}
        return dict(self._ingest)

    def get_ingest(self):
        """(input_clean.dna, input_N.dna, read_order_N.bin) as preprocess.cpp would have written them."""
        i = self._ingest
        clean = np.empty(i["n_clean"] * (self.L + 1), dtype=np.uint8)
        dnaN = np.empty(i["n_N"] * (self.L + 1), dtype=np.uint8)
        orderN = np.empty(i["n_N"], dtype=np.uint32)
        self._ck(self.lib.harcgpu_get_ingest(self.h, _ptr(clean), _ptr(dnaN), _ptr(orderN)))
        return clean, dnaN, orderN

    def load_pool_ingested(self):
        self._ck(self.lib.harcgpu_load_pool_ingested(self.h))

    def debug_sort(self, keys, vals, mode=0, begin_bit=0, end_bit=64):
        """Test hook: stable radix sort of (uint64 key, uint32 value) pairs; returns sorted copies."""
        k = np.ascontiguousarray(keys, dtype=np.uint64).copy()
        v = np.ascontiguousarray(vals, dtype=np.uint32).copy()
        self._ck(self.lib.harcgpu_debug_sort(self.h, _ptr(k), _ptr(v), k.size, mode, begin_bit, end_bit))
        return k, v

    # ---- stage I
    def load_reads(self, ascii_lines, n=None):
        """ascii_lines: bytes / uint8 array holding n lines of L bases + newline (input_clean.dna)."""
        a = np.frombuffer(ascii_lines, dtype=np.uint8) if isinstance(ascii_lines, (bytes, bytearray)) else ascii_lines
        if n is None:
            n = a.size // (self.L + 1)
        self._keep = a
        self._ck(self.lib.harcgpu_load_reads(self.h, _ptr(a), n))
        return n

    def load_reads_device(self, dptr, n):
        self._ck(self.lib.harcgpu_load_reads_device(self.h, ctypes.c_void_p(dptr), n))

    def build_dicts(self):
        self._ck(self.lib.harcgpu_build_dicts(self.h))

    def dump_dict(self, stage, l):
        nk, ni = ctypes.c_uint32(), ctypes.c_uint32()
        self._ck(self.lib.harcgpu_dump_dict(self.h, stage, l, None, None, None, ctypes.byref(nk), ctypes.byref(ni)))
        keys = np.empty(nk.value, dtype=np.uint64)
        counts = np.empty(nk.value, dtype=np.uint32)
        ids = np.empty(ni.value, dtype=np.uint32)
        self._ck(self.lib.harcgpu_dump_dict(self.h, stage, l, _ptr(keys), _ptr(counts), _ptr(ids), None, None))
        return keys, counts, ids

    def reorder(self):
        self._ck(self.lib.harcgpu_reorder(self.h))
        return self.reorder_counts()

    def reorder_counts(self):
        a, b, c = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
        self._ck(self.lib.harcgpu_reorder_counts(self.h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return a.value, b.value, c.value

    def get_singleton_ids(self):
        """read_order.bin.singleton only (the ids of this context's singletons)."""
        _, s, _ = self.reorder_counts()
        order_s = np.empty(s, dtype=np.uint32)
        self._ck(self.lib.harcgpu_get_reorder(self.h, None, None, None, None, _ptr(order_s)))
        return order_s

    def get_reorder(self):
        m, s, _ = self.reorder_counts()
        order = np.empty(m, dtype=np.uint32)
        rev = np.empty(m, dtype=np.uint8)
        flag = np.empty(m, dtype=np.uint8)
        pos = np.empty(m, dtype=np.uint8)
        order_s = np.empty(s, dtype=np.uint32)
        self._ck(self.lib.harcgpu_get_reorder(self.h, _ptr(order), _ptr(rev), _ptr(flag), _ptr(pos), _ptr(order_s)))
        return dict(order=order, rev=rev, flag=flag, pos=pos, order_s=order_s)

    def get_reordered_reads(self):
        m, s, _ = self.reorder_counts()
        dna = np.empty(m * (self.L + 1), dtype=np.uint8)
        sdna = np.empty(s * (self.L + 1), dtype=np.uint8)
        self._ck(self.lib.harcgpu_get_reordered_reads(self.h, _ptr(dna), _ptr(sdna)))
        return dna, sdna

    def counters(self):
        c = Counters()
        self._ck(self.lib.harcgpu_get_counters(self.h, ctypes.byref(c)))
        return {k: getattr(c, k) for k, _ in Counters._fields_}

    # ---- stage II
    def set_stream(self, dna, flag, pos, order, rev):
        n = len(order)
        self._keep2 = (dna, flag, pos, order, rev)
        self._ck(self.lib.harcgpu_set_stream(self.h, _ptr(dna), _ptr(flag), _ptr(pos), _ptr(order), _ptr(rev), n))

    def load_pool(self, singleton_ascii=None, order_s=None, N_ascii=None):
        ns = 0 if order_s is None else len(order_s)
        nN = 0 if N_ascii is None else len(N_ascii) // (self.L + 1)
        self._keep3 = (singleton_ascii, order_s, N_ascii)
        self._ck(self.lib.harcgpu_load_pool(self.h, _ptr(singleton_ascii), _ptr(order_s), ns, _ptr(N_ascii), nN))

    def stage_nreads(self, N_ascii):
        """Start the upload of input_N.dna now (overlaps stage I); load_pool(N_ascii=the same array) picks it up."""
        self._keepN = N_ascii
        self._ck(self.lib.harcgpu_stage_nreads(self.h, _ptr(N_ascii), len(N_ascii) // (self.L + 1)))

    # ---- one job on several GPUs (see harc_b200/multi.py for the driver)
    def job_init(self, rank, world, n_total, base, n_local):
        """Returns (CUDA IPC handle of this GPU's arena, its device pointer)."""
        h = ctypes.create_string_buffer(64)
        p = ctypes.c_void_p()
        self._ck(self.lib.harcgpu_job_init(self.h, rank, world, n_total, base, n_local, ctypes.cast(h, ctypes.c_void_p), ctypes.byref(p)))
        return h.raw, int(p.value)

    def job_connect(self, handles=None, local_ptrs=None):
        """handles: the 64-byte IPC handles of all ranks (other processes); local_ptrs: the arenas of contexts of this process."""
        if local_ptrs is not None:
            arr = (ctypes.c_void_p * len(local_ptrs))(*[ctypes.c_void_p(int(q)) for q in local_ptrs])
            self._ck(self.lib.harcgpu_job_connect(self.h, None, arr))
        else:
            buf = ctypes.create_string_buffer(b"".join(handles), 64 * len(handles))
            self._ck(self.lib.harcgpu_job_connect(self.h, ctypes.cast(buf, ctypes.c_void_p), None))

    def job_set_barrier(self, fn):
        """fn() -> None: host barrier over the ranks of the job (ranks that share one GPU; see harcgpu.h)."""
        def tramp(user):
            try:
                fn()
                return 0
            except Exception:
                return -1
        self._bar_hook = JOB_BARRIER(tramp)
        self._ck(self.lib.harcgpu_job_set_barrier(self.h, self._bar_hook, None))

    def job_load_reads(self, ascii_lines, n_local):
        a = np.frombuffer(ascii_lines, dtype=np.uint8) if isinstance(ascii_lines, (bytes, bytearray)) else ascii_lines
        self._keep = a
        self._ck(self.lib.harcgpu_job_load_reads(self.h, _ptr(a) if n_local else None, n_local))

    def job_load_reads_device(self, dptr, n_local):
        self._ck(self.lib.harcgpu_job_load_reads_device(self.h, ctypes.c_void_p(dptr), n_local))

    def device_result(self, name):
        """(device pointer, element count) of a result kept on the GPU: 'singleton_ids', 'order', 'out_order'."""
        p, n = ctypes.c_void_p(), ctypes.c_uint64()
        self._ck(self.lib.harcgpu_device_result(self.h, name.encode(), ctypes.byref(p), ctypes.byref(n)))
        return int(p.value or 0), int(n.value)

    def set_pool_exchange(self, fn):
        """fn(device_pointer, count) -> None must min-reduce the int64 array over all ranks (None removes the hook)."""
        if fn is None:
            self._hook = POOL_EXCHANGE(0)
        else:
            def tramp(user, ptr, count):
                try:
                    fn(ptr, count)
                    return 0
                except Exception:  # never let an exception cross the C boundary
                    import traceback
                    traceback.print_exc()
                    return -1
            self._hook = POOL_EXCHANGE(tramp)
        self._ck(self.lib.harcgpu_set_pool_exchange(self.h, self._hook, None))

    def load_pool_ids(self, singleton_ids, N_ascii=None, n_N=None, n_s=None):
        """singleton_ids: uint32 array, or a raw pointer (int; host or device memory) together with n_s; N_ascii: uint8
        array, or a raw pointer together with n_N."""
        if isinstance(singleton_ids, int) or singleton_ids is None:
            ids, n_s = singleton_ids, int(n_s or 0)
        else:
            ids = np.ascontiguousarray(singleton_ids, dtype=np.uint32)
            n_s = len(ids)
        if n_N is None:
            n_N = 0 if N_ascii is None else len(N_ascii) // (self.L + 1)
        self._keep3 = (ids, N_ascii)
        self._ck(self.lib.harcgpu_load_pool_ids(self.h, _ptr(ids) if n_s else None, n_s, _ptr(N_ascii) if n_N else None, n_N))

    def load_pool_device(self, dptr, n_N):
        self._ck(self.lib.harcgpu_load_pool_device(self.h, ctypes.c_void_p(dptr), n_N))

    def stream(self):
        return self.lib.harcgpu_stream(self.h)

    def encode(self):
        self._ck(self.lib.harcgpu_encode(self.h))
        s = EncodeSizes()
        self._ck(self.lib.harcgpu_get_encode_sizes(self.h, ctypes.byref(s)))
        return s

    def get_set(self, k, empty=np.empty):
        """File set k as host arrays.  `empty(n, dtype)` supplies the host buffers (e.g. views of pinned memory)."""
        z = SetSizes()
        self._ck(self.lib.harcgpu_get_set_sizes(self.h, k, ctypes.byref(z)))
        o = dict(seq=empty(z.seq_bytes, np.uint8), seq_tail=np.zeros(8, np.uint8), pos=empty(z.pos_bytes, np.uint8),
                 noise=empty(z.noise_bytes, np.uint8), noisepos=empty(z.noisepos_bytes, np.uint8),
                 rev=empty(z.rev_bytes, np.uint8), rev_tail=np.zeros(8, np.uint8))
        self._ck(self.lib.harcgpu_get_set(self.h, k, _ptr(o["seq"]), _ptr(o["seq_tail"]), _ptr(o["pos"]), _ptr(o["noise"]),
                                          _ptr(o["noisepos"]), _ptr(o["rev"]), _ptr(o["rev_tail"])))
        o["seq_tail"] = o["seq_tail"][: z.seq_tail]
        o["rev_tail"] = o["rev_tail"][: z.rev_tail]
        return o

    def get_globals(self, empty=np.empty):
        s = EncodeSizes()
        self._ck(self.lib.harcgpu_get_encode_sizes(self.h, ctypes.byref(s)))
        o = dict(order=empty(s.n_order, np.uint32), order_N=empty(s.n_order_N, np.uint32),
                 singleton=empty(s.singleton_bytes, np.uint8), singleton_tail=np.zeros(8, np.uint8),
                 input_N=empty(s.input_N_bytes, np.uint8))
        self._ck(self.lib.harcgpu_get_globals(self.h, _ptr(o["order"]), _ptr(o["order_N"]), _ptr(o["singleton"]),
                                              _ptr(o["singleton_tail"]), _ptr(o["input_N"])))
        o["singleton_tail"] = o["singleton_tail"][: s.singleton_tail]
        return o

    def get_packed_order(self):
        """(read_order.bin as pack_order.cpp would rewrite it, read_order.bin.tail) for the -p mode."""
        nb, nt = ctypes.c_uint64(), ctypes.c_uint32()
        self._ck(self.lib.harcgpu_get_packed_order(self.h, None, None, ctypes.byref(nb), ctypes.byref(nt)))
        packed = np.empty(nb.value, np.uint8)
        tail = np.empty(nt.value, np.uint32)
        self._ck(self.lib.harcgpu_get_packed_order(self.h, _ptr(packed), _ptr(tail), None, None))
        return packed, tail

    # ---- process contract
    def reorder_dir(self, basedir):
        self._ck(self.lib.harcgpu_reorder_dir(self.h, basedir.encode()))

    def encode_dir(self, basedir):
        self._ck(self.lib.harcgpu_encode_dir(self.h, basedir.encode()))

    def trim(self):
        """Give the cached device blocks back to the driver."""
        self._ck(self.lib.harcgpu_trim(self.h))

    def last_ms(self, phase):
        return self.lib.harcgpu_last_ms(self.h, phase.encode())
