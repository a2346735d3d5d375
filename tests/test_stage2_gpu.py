"""GPU parity of stage II (encoder.cpp) through the C ABI: bit-exact against the oracle for a fixed read order,
lossless through the reference's own decoder, and identical between the file and the in-memory hand-off."""
import os
import shutil

import numpy as np
import pytest

import harness as H
import refrun as R

pytestmark = pytest.mark.gpu

CASES = [
    ("s100", 20000, 100, 200000, False, True),
    ("s100rc", 20000, 100, 200000, True, False),
    ("s250", 8000, 250, 150000, False, True),
    ("s63", 20000, 63, 100000, True, True),
    ("s36", 20000, 36, 60000, True, True),
    ("rep100", 150000, 100, "repeats", True, True),   # poly-A, tandem repeats, duplications: bins beyond maxsearch in both stages
]
STAGE2_GLOBAL = ["read_meta.txt", "read_order.bin", "read_order_N_pe.bin", "read_singleton.txt", "read_singleton.txt.tail", "input_N.dna"]


def stage2_files(k):
    out = list(STAGE2_GLOBAL)
    for t in range(k):
        for stem in ("read_seq.txt", "read_pos.txt", "read_noise.txt", "read_noisepos.txt", "read_rev.txt"):
            out.append("%s.%d" % (stem, t))
        out += ["read_seq.txt.%d.tail" % t, "read_rev.txt.%d.tail" % t]
    return out


@pytest.fixture(scope="module")
def gpu():
    import harc_b200
    return harc_b200


def _stage1(workroot, case):
    name, n, L, G, rc, err = case
    d = H.dataset(workroot, case, seed=11)
    s1 = d + ".s1"
    if not os.path.exists(os.path.join(s1, "output", "temp.dna")):
        H.clone(d, s1)
        H.oracle_reorder(s1, L, 1)
    return d, s1, L


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_pool_dictionary_bit_exact(gpu, workroot, case):
    d, s1, L = _stage1(workroot, case)
    dump = os.path.join(s1, "dict2.bin")
    R.dictdump(s1, L, dump, stage2=True)
    raw = np.fromfile(dump, dtype=np.uint8)
    out = os.path.join(s1, "output")
    g = gpu.HarcGpu(L)
    order_s = np.fromfile(os.path.join(out, "read_order.bin.singleton"), dtype=np.uint32)
    g.load_pool(np.fromfile(os.path.join(out, "temp.dna.singleton"), dtype=np.uint8), order_s,
                np.fromfile(os.path.join(out, "input_N.dna"), dtype=np.uint8))
    off = 4
    for l in range(2):
        nk, nid = raw[off:off + 8].view(np.uint32)
        off += 8
        rec = raw[off:off + 12 * int(nk)].view(np.dtype([("k", "<u8"), ("c", "<u4")]))
        off += 12 * int(nk)
        ids = raw[off:off + 4 * int(nid)].view(np.uint32)
        off += 4 * int(nid)
        keys, counts, gids = g.dump_dict(2, l)
        assert np.array_equal(keys, rec["k"]) and np.array_equal(counts, rec["c"]) and np.array_equal(gids, ids)
    g.close()


@pytest.mark.parametrize("sets", [1, 2, 5])
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_encode_fixed_order_bit_exact(gpu, workroot, case, sets):
    """Same stage I files in, every stage II file byte-identical to the oracle (= reference at num_thr=1 for sets=1)."""
    d, s1, L = _stage1(workroot, case)
    o = H.clone(s1, s1 + ".oe%d" % sets)
    al = H.oracle_encode(o, L, sets)
    g = H.clone(s1, s1 + ".ge%d" % sets)
    ctx = gpu.HarcGpu(L, file_sets=sets)
    ctx.encode_dir(g)
    es = ctx.encode()  # idempotent: encoding the same resident inputs again gives the same sizes
    ctx.close()
    assert (es.aligned_singletons, es.aligned_N) == al
    assert H.same_files(o, g, stage2_files(sets)) == []


def test_encode_empty_pool_and_empty_stream(gpu, workroot):
    """Pool absent (encoder.cpp:147,231 skip) and stream absent (only unaligned pool reads come out)."""
    d, s1, L = _stage1(workroot, CASES[0])
    for variant in ("nopool", "nostream"):
        src = H.clone(s1, s1 + "." + variant)
        out = os.path.join(src, "output")
        if variant == "nopool":
            for f in ("temp.dna.singleton", "read_order.bin.singleton", "input_N.dna"):
                open(os.path.join(out, f), "wb").close()
        else:
            for f in ("temp.dna", "tempflag.txt", "temppos.txt", "read_order.bin", "read_rev.txt"):
                open(os.path.join(out, f), "wb").close()
        o = H.clone(src, src + ".o")
        H.oracle_encode(o, L, 2)
        g = H.clone(src, src + ".g")
        ctx = gpu.HarcGpu(L, file_sets=2)
        ctx.encode_dir(g)
        ctx.close()
        assert H.same_files(o, g, stage2_files(2)) == []


@pytest.mark.parametrize("case", [CASES[0], CASES[2], CASES[4]], ids=["s100", "s250", "s36"])
def test_pipeline_lossless_through_reference_decoder(gpu, workroot, case):
    """GPU reorder (many walkers) + GPU encode, decoded by the reference's unmodified decoder.out: the multiset of
    reads must equal the input (harc -d order-free contract)."""
    name, n, L, G, rc, err = case
    d = H.dataset(workroot, case, seed=11)
    g = H.clone(d, d + ".pipe")
    ctx = gpu.HarcGpu(L, walkers=32, file_sets=3)
    ctx.reorder_dir(g)
    ctx.encode_dir(g)
    ctx.close()
    R.decoder(g)
    fq = np.fromfile(os.path.join(d, "r.fastq"), dtype=np.uint8)
    lines = fq.tobytes().split(b"\n")[1::4]
    want = os.path.join(g, "all.dna")
    with open(want, "wb") as f:
        f.write(b"\n".join(lines) + b"\n")
    assert R.sorted_lines_digest(os.path.join(g, "output", "output.dna"), L) == R.sorted_lines_digest(want, L)


def test_in_memory_handoff_equals_file_path(gpu, workroot):
    """reorder -> encode on one context (streams stay on the device) == reorder_dir -> encode_dir through files."""
    name, n, L, G, rc, err = CASES[0]
    d = H.dataset(workroot, CASES[0], seed=11)
    g = H.clone(d, d + ".mem")
    out = os.path.join(g, "output")
    ctx = gpu.HarcGpu(L, walkers=1, file_sets=2)
    ctx.load_reads(np.fromfile(os.path.join(out, "input_clean.dna"), dtype=np.uint8))
    ctx.reorder()
    ctx.load_pool(None, None, np.fromfile(os.path.join(out, "input_N.dna"), dtype=np.uint8))
    ctx.encode()
    ctx.trim()  # giving the cached blocks back to the driver must leave the results alone
    sets = [ctx.get_set(k) for k in range(2)]
    glob = ctx.get_globals()
    ctx.close()
    f = H.clone(d, d + ".file")
    ctx = gpu.HarcGpu(L, walkers=1, file_sets=2)
    ctx.reorder_dir(f)
    ctx.encode_dir(f)
    ctx.close()
    fo = os.path.join(f, "output")
    rd = lambda nme: np.fromfile(os.path.join(fo, nme), dtype=np.uint8)
    for k in range(2):
        for key, stem in (("seq", "read_seq.txt"), ("pos", "read_pos.txt"), ("noise", "read_noise.txt"),
                          ("noisepos", "read_noisepos.txt"), ("rev", "read_rev.txt")):
            assert np.array_equal(sets[k][key], rd("%s.%d" % (stem, k))), (key, k)
        assert np.array_equal(sets[k]["seq_tail"], rd("read_seq.txt.%d.tail" % k))
        assert np.array_equal(sets[k]["rev_tail"], rd("read_rev.txt.%d.tail" % k))
    assert np.array_equal(glob["order"], np.fromfile(os.path.join(fo, "read_order.bin"), dtype=np.uint32))
    assert np.array_equal(glob["order_N"], np.fromfile(os.path.join(fo, "read_order_N_pe.bin"), dtype=np.uint32))
    assert np.array_equal(glob["singleton"], rd("read_singleton.txt"))
    assert np.array_equal(glob["input_N"], rd("input_N.dna"))
