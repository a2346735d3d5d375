"""Generates tests/golden/*: small inputs plus the outputs of the reference's OWN binaries (oracle/_ref, num_thr=1,
the only deterministic setting) for them.  Run here, where /root/reference exists; the fixtures travel with the repo.

    python oracle/make_golden.py

Each fixture directory holds in/ (input_clean.dna, input_N.dna, numreads.bin, read_order_N.bin), s1/ (the seven
stage I files written by reorder.out) and s2/ (every file encoder.out leaves behind), plus dict1.bin / dict2.bin
(canonical dictionary dumps from the include-wrapper around the unmodified TUs).
"""
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refrun as R
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tools"))
import workload as W
import numpy as np

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
# name, reads, L, genome, rc, errors, seed
FIXTURES = [
    ("L100_noRC_err", 600, 100, 4000, False, True, 21),
    ("L100_RC", 600, 100, 6000, True, False, 22),
    ("L250_err", 250, 250, 4000, False, True, 23),
    ("L63_RC_err", 700, 63, 3000, True, True, 24),
    ("L36_RC_err", 900, 36, 2500, True, True, 25),
    # repeat-rich (refrun.repeat_rich_genome, scaled down): a poly-A run, a tandem repeat and duplications give dictionary
    # bins beyond maxsearch = 1000 in stage I and bins of hundreds in the stage II pool
    ("L36_repeats", 9000, 36, "repeats", True, True, 26),
]
S1 = ["temp.dna", "temp.dna.singleton", "read_rev.txt", "tempflag.txt", "temppos.txt", "read_order.bin", "read_order.bin.singleton"]


def main():
    for name, n, L, G, rc, err, seed in FIXTURES:
        dst = os.path.join(GOLD, name)
        shutil.rmtree(dst, ignore_errors=True)
        os.makedirs(dst)
        with tempfile.TemporaryDirectory() as tmp:
            # inputs from tools/simreads.c (the reference's gen_fastq overruns its buffer on genomes this small);
            # laid out as preprocess.cpp:98-131 writes them.  Everything under s1/ s2/ comes from the reference binaries.
            if G == "repeats":
                g = R.repeat_rich_genome(seed, unique=3000, polyA=3000, tandem_unit=5, tandem_copies=100, dup_len=60, dup_copies=6,
                                         div_len=80, div_copies=2)
                W.write_dir(W.make(n, L, len(g), rc=rc, errors=err, seed=seed, genome=g), tmp)
                G = len(g)
            else:
                W.write_dir(W.make(n, L, G, rc=rc, errors=err, seed=seed), tmp)
            out = os.path.join(tmp, "output")
            shutil.copytree(out, os.path.join(dst, "in"))
            R.dictdump(tmp, L, os.path.join(dst, "dict1.bin"))
            R.reorder(tmp, L, 1)
            os.makedirs(os.path.join(dst, "s1"))
            for f in S1:
                shutil.copy(os.path.join(out, f), os.path.join(dst, "s1", f))
            R.dictdump(tmp, L, os.path.join(dst, "dict2.bin"), stage2=True)
            R.encoder(tmp, L, 1)
            os.makedirs(os.path.join(dst, "s2"))
            for f in sorted(os.listdir(out)):
                if f.startswith("read_") or f == "input_N.dna":
                    if f in ("read_order_N.bin", "read_rev.txt", "read_order.bin.singleton"):
                        continue
                    shutil.copy(os.path.join(out, f), os.path.join(dst, "s2", f))
            with open(os.path.join(dst, "meta.txt"), "w") as fh:
                fh.write("L=%d reads=%d genome=%d rc=%d errors=%d seed=%d\n" % (L, n, G, rc, err, seed))
        print(name, sum(os.path.getsize(os.path.join(dp, f)) for dp, _, fs in os.walk(dst) for f in fs), "bytes")


if __name__ == "__main__":
    main()
