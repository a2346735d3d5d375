"""Test infrastructure only (oracle).  Drives the reference's own programs, built UNMODIFIED into
oracle/_ref/ by oracle/Makefile, through the process + file contract of harc:50-69 / harc:171-188.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
Nothing here reads /root/reference at run time: the binaries under oracle/_ref/ are self-contained.
"""
import bz2
import lzma
import os
import shutil
import subprocess
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def have_ref(L=100, T=1):
    return os.path.exists(os.path.join(REF, "L%d_T%d" % (L, T), "reorder.out"))


def ref_threads_available(L=100):
    out = []
    if os.path.isdir(REF):
        for d in os.listdir(REF):
            if d.startswith("L%d_T" % L) and os.path.exists(os.path.join(REF, d, "reorder.out")):
                out.append(int(d.split("_T")[1]))
    return sorted(out)


def _run(cmd, cwd, timeout=3600, quiet=True):
    # BooPHF drops temp_p<pid>_level_<i> scratch files into the CWD (BooPHF.h:1376-1397): always run in a
    # scratch cwd and under a timeout (SURVEY §7 "oracle hazard").
    t0 = time.time()
    r = subprocess.run(cmd, cwd=cwd, timeout=timeout, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    dt = time.time() - t0
    if r.returncode != 0:
        raise RuntimeError("%s failed (%d): %s" % (cmd, r.returncode, r.stdout.decode(errors="replace")[-2000:]))
    return dt, r.stdout.decode(errors="replace")


def repeat_rich_genome(seed=1, unique=300000, polyA=80000, tandem_unit=7, tandem_copies=2000, dup_len=300, dup_copies=60,
                       div_len=1000, div_copies=20):
    """A genome with what real ones have and i.i.d. ones lack: a long poly-A run, a tandem repeat of a short unit, an exact
    dispersed duplication and a family of 1 %-diverged copies, between stretches of unique sequence.  Dictionary bins of
    thousands of reads (beyond maxsearch = 1000) in both stages follow from it."""
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    rnd = lambda n: acgt[rng.integers(0, 4, size=n, dtype=np.uint8)]
    parts = [rnd(unique // 4), np.full(polyA, ord("A"), np.uint8), rnd(unique // 4), np.tile(rnd(tandem_unit), tandem_copies)]
    dup = rnd(dup_len)
    fam = rnd(div_len)
    gap = unique // 2 // (dup_copies + div_copies + 1)
    for k in range(dup_copies + div_copies):
        parts.append(rnd(gap))
        if k % 4 == 3 and k // 4 < div_copies:
            c = fam.copy()
            pos = rng.integers(0, div_len, size=div_len // 100)
            c[pos] = acgt[rng.integers(0, 4, size=pos.size, dtype=np.uint8)]
            parts.append(c)
        else:
            parts.append(dup)
    parts.append(rnd(gap))
    return np.concatenate(parts)


def make_genome(path, nbases, seed=1, line=100, bases=None):
    """Seeded uniform i.i.d. ACGT FASTA (the reference has no genome generator); `bases` writes a given genome."""
    rng = np.random.default_rng(seed)
    if bases is not None:
        g, nbases = bases, len(bases)
    else:
        g = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=nbases, dtype=np.uint8)]
    nl = (nbases + line - 1) // line
    with open(path, "wb") as f:
        f.write(b">synthetic seed=%d len=%d\n" % (seed, nbases))
        pad = np.full(nl * line, ord("A"), dtype=np.uint8)
        pad[:nbases] = g
        rows = np.empty((nl, line + 1), dtype=np.uint8)
        rows[:, :line] = pad.reshape(nl, line)
        rows[:, line] = ord("\n")
        buf = rows.reshape(-1)
        # cut the padding of the last row
        last = nbases - (nl - 1) * line
        f.write(buf[: (nl - 1) * (line + 1) + last].tobytes())
        f.write(b"\n")
    return g


def gen_fastq(genome_fa, out_fastq, nreads, L, rc=False, errors=False, scratch=None):
    """util/gen_fastq (odd reads reverse-complemented) or util/gen_fastq_noRC, default-seeded mt19937."""
    exe = os.path.join(REF, "gen_fastq" if rc else "gen_fastq_noRC")
    cmd = [exe, str(nreads), str(L), genome_fa, out_fastq] + (["-e"] if errors else [])
    return _run(cmd, scratch or os.path.dirname(out_fastq))


def preprocess(fastq, basedir, L, preserve_order=False):
    """harc:42-50.  Creates <basedir>/output/{input_clean.dna,input_N.dna,read_order_N.bin,numreads.bin}."""
    os.makedirs(os.path.join(basedir, "output"), exist_ok=True)
    cmd = [os.path.join(REF, "preprocess.out"), fastq, basedir, "True" if preserve_order else "False", "False", str(L)]
    return _run(cmd, basedir)


def reorder(basedir, L, T=1, timeout=3600):
    """harc:65-67 with num_thr=T."""
    return _run([os.path.join(REF, "L%d_T%d" % (L, T), "reorder.out"), basedir], os.path.join(basedir, "output"), timeout)


def encoder(basedir, L, T=1, timeout=3600):
    """harc:68-69 with num_thr=T."""
    return _run([os.path.join(REF, "L%d_T%d" % (L, T), "encoder.out"), basedir], os.path.join(basedir, "output"), timeout)


def dictdump(basedir, L, out, stage2=False, T=1):
    exe = os.path.join(REF, "L%d_T%d" % (L, T), "dictdump_s.out" if stage2 else "dictdump.out")
    return _run([exe, basedir, out], os.path.join(basedir, "output"))


def count_file_sets(basedir):
    out = os.path.join(basedir, "output")
    return len([f for f in os.listdir(out) if f.startswith("read_pos.txt")])


def decoder(basedir, threads=4):
    """harc:171,188 order-free decode -> output/output.dna (destroys the packed streams: run on a copy)."""
    k = count_file_sets(basedir)
    return _run([os.path.join(REF, "decoder.out"), basedir, str(threads), str(k)], os.path.join(basedir, "output"))


def decoder_preserve(basedir, L):
    """harc:172-185 (-p): unpack_order -> decoder_preserve -> merge_N -> output/output.dna in FASTQ order."""
    k = count_file_sets(basedir)
    cwd = os.path.join(basedir, "output")
    _run([os.path.join(REF, "unpack_order.out"), basedir], cwd)
    _run([os.path.join(REF, "L%d_K%d" % (L, k), "decoder_preserve.out"), basedir], cwd)
    return _run([os.path.join(REF, "merge_N.out"), basedir], cwd)


def pack_order(basedir):
    return _run([os.path.join(REF, "pack_order.out"), basedir], os.path.join(basedir, "output"))


def copy_stage(basedir, dst, names=None):
    """Copy <basedir>/output (or a subset of files) to <dst>/output."""
    os.makedirs(os.path.join(dst, "output"), exist_ok=True)
    src = os.path.join(basedir, "output")
    for f in os.listdir(src):
        if names is None or f in names:
            shutil.copy(os.path.join(src, f), os.path.join(dst, "output", f))
    return dst


BSC_STEMS = ("read_pos.txt", "read_noise.txt", "read_seq.txt")      # bsc -b64p in harc:102-109
LZMA_STEMS = ("read_noisepos.txt", "read_rev.txt")                   # 7z in harc:104-107


def _cat(out, stem):
    """Concatenate <stem>.<k> (+ .tail) for k = 0.. in order (what the per-stem tar holds, harc:89-93)."""
    names = sorted([f for f in os.listdir(out) if f.startswith(stem + ".")],
                   key=lambda s: (int(s.split(".")[2]), s.endswith(".tail")))
    return b"".join(open(os.path.join(out, n), "rb").read() for n in names)


def standin_size(basedir, include_order=False):
    """Stage III stand-in (bsc/7z are absent): bz2 -9 for the bsc streams, xz for the 7z ones.
    Returns (total_bytes, per-stream dict).  Same function is applied to the reference's and to our output."""
    out = os.path.join(basedir, "output")
    jobs = {}  # name -> (function, bytes); the compressors release the GIL, so the streams are squeezed side by side
    for stem in BSC_STEMS:
        jobs[stem] = (lambda b: len(bz2.compress(b, 9)), _cat(out, stem))
    for stem in LZMA_STEMS:
        jobs[stem] = (lambda b: len(lzma.compress(b, preset=6)), _cat(out, stem))
    extra = {}
    for f in ("read_singleton.txt", "input_N.dna"):
        jobs[f] = (lambda b: len(bz2.compress(b, 9)), open(os.path.join(out, f), "rb").read())
        t = os.path.join(out, f + ".tail")
        extra[f] = os.path.getsize(t) if os.path.exists(t) else 0
    jobs["read_meta.txt"] = (lambda b: len(lzma.compress(b)), open(os.path.join(out, "read_meta.txt"), "rb").read())
    if include_order:
        for f in ("read_order.bin", "read_order_N.bin", "read_order_N_pe.bin"):
            p = os.path.join(out, f)
            if os.path.exists(p):
                jobs[f] = (lambda b: len(lzma.compress(b, preset=6)), open(p, "rb").read())
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
        futs = {k: ex.submit(fn, data) for k, (fn, data) in jobs.items()}
        sizes = {k: f.result() + extra.get(k, 0) for k, f in futs.items()}
    return sum(sizes.values()), sizes


def sorted_lines_digest(path, L):
    """Order-free multiset digest of a file of fixed-length lines (for round-trip checks)."""
    import hashlib
    a = np.fromfile(path, dtype=np.uint8)
    assert a.size % (L + 1) == 0, (a.size, L)
    rows = a.reshape(-1, L + 1)
    v = np.ascontiguousarray(rows).view([("", "V%d" % (L + 1))]).reshape(-1)
    v = np.sort(v)
    return rows.shape[0], hashlib.md5(v.tobytes()).hexdigest()
