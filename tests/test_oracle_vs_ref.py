"""CPU: the oracle restatement against the reference's own binaries (oracle/_ref) run here on seeded inputs.
Skipped where oracle/_ref is absent (it needs /root/reference at build time; the golden vectors then stand in)."""
import os

import pytest

import harness as H
import refrun as R
from test_stage2_gpu import stage2_files

pytestmark = pytest.mark.skipif(not R.have_ref(100, 1), reason="oracle/_ref not built")

CASES = [
    ("o100", 12000, 100, 120000, False, True),
    ("o100rc", 12000, 100, 120000, True, False),
    ("o250", 4000, 250, 100000, False, True),
    ("o63", 12000, 63, 80000, True, True),
    ("o36", 12000, 36, 60000, True, True),
    ("orep100", 150000, 100, "repeats", True, True),   # poly-A / tandem repeats / duplications: bins beyond maxsearch in both stages
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_both_stages_byte_identical_at_one_thread(workroot, case):
    name, n, L, G, rc, err = case
    d = H.dataset(workroot, case, seed=31)
    r = H.clone(d, d + ".ref")
    R.reorder(r, L, 1)
    o = H.clone(d, d + ".ora")
    H.oracle_reorder(o, L, 1)
    assert H.same_files(r, o, H.STAGE1_FILES) == []
    _, txt = R.encoder(r, L, 1)
    al = H.oracle_encode(o, L, 1)
    assert "%d singleton reads were aligned" % al[0] in txt and "%d reads with N were aligned" % al[1] in txt
    assert H.same_files(r, o, stage2_files(1)) == []


def test_two_file_sets_match_reference_two_threads_with_empty_pool(workroot):
    """With an empty pool the reference encoder is deterministic at any num_thr: compare K=2."""
    name, n, L, G, rc, err = CASES[1]
    d = H.make_dataset(workroot, name, n, L, G, rc, err, seed=31)
    s1 = H.clone(d, d + ".k2")
    H.oracle_reorder(s1, L, 1)
    for f in ("temp.dna.singleton", "read_order.bin.singleton", "input_N.dna"):
        open(os.path.join(s1, "output", f), "wb").close()
    r = H.clone(s1, s1 + ".ref")
    R.encoder(r, L, 2)
    o = H.clone(s1, s1 + ".ora")
    H.oracle_encode(o, L, 2)
    assert H.same_files(r, o, stage2_files(2)) == []


def test_reference_decoder_round_trip_of_oracle_output(workroot):
    name, n, L, G, rc, err = CASES[0]
    d = H.make_dataset(workroot, name, n, L, G, rc, err, seed=31)
    o = H.clone(d, d + ".rt")
    H.oracle_reorder(o, L, 4)  # four emulated walkers
    H.oracle_encode(o, L, 3)
    R.decoder(o)
    import numpy as np
    fq = np.fromfile(os.path.join(d, "r.fastq"), dtype=np.uint8).tobytes().split(b"\n")[1::4]
    want = os.path.join(o, "all.dna")
    open(want, "wb").write(b"\n".join(fq) + b"\n")
    assert R.sorted_lines_digest(os.path.join(o, "output", "output.dna"), L) == R.sorted_lines_digest(want, L)
