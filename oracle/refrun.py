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


def make_genome(path, nbases, seed=1, line=100):
    """Seeded uniform i.i.d. ACGT FASTA (the reference has no genome generator)."""
    rng = np.random.default_rng(seed)
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
    sizes = {}
    for stem in BSC_STEMS:
        sizes[stem] = len(bz2.compress(_cat(out, stem), 9))
    for stem in LZMA_STEMS:
        sizes[stem] = len(lzma.compress(_cat(out, stem), preset=6))
    for f in ("read_singleton.txt", "input_N.dna"):
        sizes[f] = len(bz2.compress(open(os.path.join(out, f), "rb").read(), 9))
        t = os.path.join(out, f + ".tail")
        if os.path.exists(t):
            sizes[f] += os.path.getsize(t)
    sizes["read_meta.txt"] = len(lzma.compress(open(os.path.join(out, "read_meta.txt"), "rb").read()))
    if include_order:
        for f in ("read_order.bin", "read_order_N.bin", "read_order_N_pe.bin"):
            p = os.path.join(out, f)
            if os.path.exists(p):
                sizes[f] = len(lzma.compress(open(p, "rb").read(), preset=6))
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
