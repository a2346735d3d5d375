"""GPU: fused FASTQ ingest against the reference's own preprocess.out (preprocess.cpp:49-138) on the
same file: input_clean.dna, input_N.dna, read_order_N.bin and numreads.bin byte for byte; the packed reads it leaves on
the device drive stage I and II to the same output as the file route; error and edge cases of the reference."""
import os

import numpy as np
import pytest

import harness as H

pytestmark = pytest.mark.gpu

CASES = [("ing_L100", 30000, 100, 300000, False, True), ("ing_L36_rc", 20000, 36, 100000, True, True),
         ("ing_L250", 8000, 250, 200000, False, True), ("ing_L63_clean", 10000, 63, 100000, True, False)]


def _ref(d, name):
    return np.fromfile(os.path.join(d, "output", name), dtype=np.uint8)


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_ingest_matches_reference_preprocess(workroot, case):
    import harc_b200
    name, n, L, G, rc, err = case
    d = H.make_dataset(workroot, name, n, L, G, rc, err, seed=5)
    fq = np.fromfile(os.path.join(d, "r.fastq"), dtype=np.uint8)
    assert harc_b200.fastq_readlen(fq) == L
    ctx = harc_b200.HarcGpu(L, walkers=1)
    info = ctx.ingest_fastq(fq)
    numreads = int(np.fromfile(os.path.join(d, "output", "numreads.bin"), dtype=np.uint32)[0])
    assert info["total_reads"] == n and info["n_clean"] == numreads and info["n_clean"] + info["n_N"] == n
    clean, dnaN, orderN = ctx.get_ingest()
    assert np.array_equal(clean, _ref(d, "input_clean.dna"))
    assert np.array_equal(dnaN, _ref(d, "input_N.dna"))
    assert np.array_equal(orderN, np.fromfile(os.path.join(d, "output", "read_order_N.bin"), dtype=np.uint32))
    ctx.close()


def test_ingest_drives_both_stages_like_the_file_route(workroot):
    """ingest -> dictionaries -> one-walker reorder -> pool from (device singletons ++ ingested N reads) -> encode equals
    load_reads(input_clean.dna) -> ... -> load_pool(input_N.dna) on every stream."""
    import harc_b200
    L = 100
    d = H.make_dataset(workroot, "ing_pipe", 40000, L, 400000, True, True, seed=6)
    fq = np.fromfile(os.path.join(d, "r.fastq"), dtype=np.uint8)
    a = harc_b200.HarcGpu(L, walkers=1)
    a.ingest_fastq(fq)
    a.build_dicts()
    a.reorder()
    a.load_pool_ingested()
    a.encode()
    b = harc_b200.HarcGpu(L, walkers=1)
    b.load_reads(_ref(d, "input_clean.dna"))
    b.build_dicts()
    b.reorder()
    b.load_pool(N_ascii=_ref(d, "input_N.dna"))
    b.encode()
    ra, rb = a.get_reorder(), b.get_reorder()
    for k in ra:
        assert np.array_equal(ra[k], rb[k]), k
    for oa, ob in ((a.get_set(0), b.get_set(0)), (a.get_globals(), b.get_globals())):
        assert sorted(oa) == sorted(ob)
        for k in oa:
            assert np.array_equal(np.asarray(oa[k]), np.asarray(ob[k])), k
    a.close()
    b.close()


def test_ingest_edge_cases():
    import harc_b200
    L = 8
    ctx = harc_b200.HarcGpu(L, walkers=1)
    # no trailing newline on the last (quality) line; a read with N in the middle; record numbers of the N reads
    fq = b"@a\nACGTACGT\n+\nIIIIIIII\n@b\nACNTACGT\n+\nIIIIIIII\n@c\nTTTTGGGG\n+\nIIIIIIII"
    info = ctx.ingest_fastq(fq)
    assert (info["total_reads"], info["n_clean"], info["n_N"]) == (3, 2, 1)
    clean, dnaN, orderN = ctx.get_ingest()
    assert clean.tobytes() == b"ACGTACGT\nTTTTGGGG\n" and dnaN.tobytes() == b"ACNTACGT\n" and orderN.tolist() == [1]
    # a file that stops after a sequence line: the read is written, the record is not counted (preprocess.cpp:81-121)
    info = ctx.ingest_fastq(b"@a\nACGTACGT\n+\nIIIIIIII\n@b\nGGGGCCCC\n")
    assert (info["total_reads"], info["n_clean"], info["n_N"]) == (1, 2, 0)
    # empty file
    info = ctx.ingest_fastq(b"")
    assert (info["total_reads"], info["n_clean"], info["n_N"]) == (0, 0, 0)
    # preprocess.cpp:92-97: two different read lengths
    with pytest.raises(harc_b200.HarcError, match="Read length not fixed.*8 and 6"):
        ctx.ingest_fastq(b"@a\nACGTACGT\n+\nIIIIIIII\n@b\nACGTAC\n+\nIIIIII\n")
    with pytest.raises(harc_b200.HarcError, match="Read length not fixed.*8 and 10"):
        ctx.ingest_fastq(b"@a\nACGTACGTAC\n+\nIIIIIIIIII\n")
    ctx.close()
