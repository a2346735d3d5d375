"""GPU: the boundary that ships.  The reference's own `harc` script (harc:35-192) with the two-line change of
INTEGRATION.md §1 applied (oracle/_ref/harc_cli/harc, made from the reference's file by oracle/Makefile) drives
harc_b200/reorder.out and harc_b200/encoder.out as processes: `harc -c x.fastq` then `harc -d x.harc`, order-free and
with -p, next to the reference's unmodified preprocess / pack_order / unpack_order / decoder / decoder_preserve / merge_N
binaries.  bsc and 7z (stage III, not installable here: SURVEY §0.4) are pass-through stubs with the same argv."""
import os
import shutil
import stat
import subprocess

import numpy as np
import pytest

import harness as H
import refrun as R

pytestmark = pytest.mark.gpu
CLI = os.path.join(R.REF, "harc_cli", "harc")


def _script(path, body):
    with open(path, "w") as f:
        f.write("#!/bin/bash\n" + body)
    os.chmod(path, os.stat(path).st_mode | stat.S_IXUSR | stat.S_IXGRP | stat.S_IXOTH)


def _home(root):
    """A checkout-shaped directory: ./harc, ./src/*.out, ./src/libbsc/bsc, plus a 7z on PATH."""
    home = os.path.join(root, "harc_home")
    shutil.rmtree(home, ignore_errors=True)
    os.makedirs(os.path.join(home, "src", "libbsc"))
    os.makedirs(os.path.join(home, "bin"))
    shutil.copy(CLI, os.path.join(home, "harc"))
    for exe in ("preprocess.out", "pack_order.out", "unpack_order.out", "merge_N.out", "decoder.out"):
        shutil.copy(os.path.join(R.REF, exe), os.path.join(home, "src", exe))
    # harc:102-109 `bsc e in out -b64p -tN`, harc:155-162 `bsc d in out -tN`
    _script(os.path.join(home, "src", "libbsc", "bsc"), 'cp "$2" "$3"\n')
    # harc:104-115 `7z a out.7z in -mmt=N`, harc:157-182 `7z e in.7z -oDIR`
    _script(os.path.join(home, "bin", "7z"),
            'if [ "$1" = a ]; then cp "$3" "$2"; elif [ "$1" = e ]; then d="${3#-o}"; cp "$2" "$d/$(basename "${2%.7z}")"; else exit 2; fi\n')
    env = dict(os.environ)
    env["PATH"] = os.path.join(home, "bin") + os.pathsep + env.get("PATH", "")
    env["HARCGPU"] = os.path.join(H.ROOT, "harc_b200")
    env["HARC_PREBUILT"] = R.REF
    return home, env


def _harc(home, env, *args):
    return subprocess.run(["./harc"] + list(args), cwd=home, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=900)


def _fastq(work, n, L, G, rc, seed):
    os.makedirs(work, exist_ok=True)
    R.make_genome(os.path.join(work, "g.fa"), G, seed=seed)
    fq = os.path.join(work, "x.fastq")
    R.gen_fastq(os.path.join(work, "g.fa"), fq, n, L, rc=rc, errors=True)
    os.remove(os.path.join(work, "g.fa"))
    return fq, np.fromfile(fq, dtype=np.uint8).tobytes().split(b"\n")[1::4]


@pytest.mark.skipif(not os.path.exists(CLI), reason="oracle/_ref/harc_cli/harc is made where /root/reference is mounted")
@pytest.mark.parametrize("L,preserve", [(100, False), (100, True), (250, True)], ids=["L100", "L100_p", "L250_p"])
def test_harc_cli_round_trip(workroot, L, preserve):
    home, env = _home(workroot)
    work = os.path.join(workroot, "cli_L%d_%d" % (L, preserve))
    shutil.rmtree(work, ignore_errors=True)
    fq, reads = _fastq(work, 30000, L, 300000, L == 100, seed=31)
    flags = ["-p"] if preserve else []
    r = _harc(home, env, "-c", fq, "-t", "2", *flags)
    out = r.stdout.decode(errors="replace")
    assert r.returncode == 0, out
    # the progress lines of the two programs (reorder.cpp:701, encoder.cpp:507-508) as scripts grep them
    assert "Reordering done, " in out and " were unmatched" in out
    assert " singleton reads were aligned" in out and " reads with N were aligned" in out
    arch = os.path.join(work, "x.harc")
    assert os.path.exists(arch) and not os.path.exists(os.path.join(work, "output"))   # harc:134-135
    r = _harc(home, env, "-d", arch, "-t", "2", *flags)
    assert r.returncode == 0, r.stdout.decode(errors="replace")
    got = open(os.path.join(work, "x.dna.d"), "rb").read()
    if preserve:
        assert got == b"\n".join(reads) + b"\n"                     # the FASTQ's reads in the FASTQ's order
    else:
        assert sorted(got.split(b"\n")[:-1]) == sorted(reads)       # the same multiset of reads


@pytest.mark.skipif(not os.path.exists(CLI), reason="oracle/_ref/harc_cli/harc is made where /root/reference is mounted")
def test_harc_cli_stops_on_errors(workroot):
    """`set -e` (harc:2): a failing stage ends the pipeline with a non-zero exit code and no archive."""
    home, env = _home(workroot)
    work = os.path.join(workroot, "cli_bad")
    shutil.rmtree(work, ignore_errors=True)
    fq, reads = _fastq(work, 2000, 100, 100000, False, seed=5)
    lines = open(fq, "rb").read().split(b"\n")
    lines[4 * 100 + 1] = lines[4 * 100 + 1][:-7]                    # one read of another length (preprocess.cpp:92-97)
    lines[4 * 100 + 3] = lines[4 * 100 + 3][:-7]
    open(fq, "wb").write(b"\n".join(lines))
    r = _harc(home, env, "-c", fq)
    assert r.returncode != 0 and not os.path.exists(os.path.join(work, "x.harc"))
    # our programs on their own: inputs that do not add up -> message on stderr, exit code 1, nothing half-written reported as done
    bad = os.path.join(workroot, "cli_bad2")
    shutil.rmtree(bad, ignore_errors=True)
    os.makedirs(os.path.join(bad, "output"))
    open(os.path.join(bad, "output", "input_clean.dna"), "wb").write((b"ACGT" * 25 + b"\n") * 10)
    np.array([1000], dtype=np.uint32).tofile(os.path.join(bad, "output", "numreads.bin"))   # claims more reads than the file holds
    open(os.path.join(bad, "output", "input_N.dna"), "wb").close()
    for exe in ("reorder.out", "encoder.out"):
        r = subprocess.run([os.path.join(H.ROOT, "harc_b200", exe), bad], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
        assert r.returncode == 1 and r.stderr, (exe, r)
    r = subprocess.run([os.path.join(H.ROOT, "harc_b200", "reorder.out")], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 1 and b"usage" in r.stderr
