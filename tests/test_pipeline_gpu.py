"""GPU: whole hot path (many walkers, auto settings) against the reference's own programs on the same input:
compressed size (stage III stand-in) within 2 % of the reference, and the order-preserving (-p) round trip through
the reference's unpack_order / decoder_preserve / merge_N (harc:111-115, 172-185)."""
import json
import os

import numpy as np
import pytest

import harness as H
import refrun as R

pytestmark = pytest.mark.gpu

# name, reads, L, genome, rc, errors   -- scaled shapes of BASELINE.json configs[0..2]
SIZE_CASES = [
    ("c1_10x_clean", 300000, 100, 3000000, False, False),
    ("c2_70x_err", 350000, 100, 500000, False, True),
    ("c3_rc_20x", 300000, 100, 1500000, True, False),
]


def _ref_size(d, L, T):
    r = H.clone(d, d + ".ref%d" % T)
    R.reorder(r, L, T)
    R.encoder(r, L, T)
    return R.standin_size(r)[0]


@pytest.mark.parametrize("case", SIZE_CASES, ids=[c[0] for c in SIZE_CASES])
def test_compressed_size_within_2_percent_of_reference(workroot, case):
    """North star: 'the reorder itself is held to compressed bits/base within 2 % of the reference run with -t 1 and
    -t 8'.  The bound is asserted against BOTH runs (the reference's two runs differ from each other by up to 3.4 % on
    noisy data; ours has to stay within 2 % of the smaller one as well) and both ratios are printed."""
    import harc_b200
    name, n, L, G, rc, err = case
    d = H.make_dataset(workroot, name, n, L, G, rc, err, seed=21)
    s1, s8 = _ref_size(d, L, 1), _ref_size(d, L, 8)
    g = H.clone(d, d + ".gpu")
    ctx = harc_b200.HarcGpu(L, walkers=0, file_sets=1)
    ctx.reorder_dir(g)
    ctx.encode_dir(g)
    ctx.close()
    sg = R.standin_size(g)[0]
    bpb = lambda s: 8.0 * s / (n * L)
    print(json.dumps({"case": name, "bits_per_base": {"gpu": bpb(sg), "ref_t1": bpb(s1), "ref_t8": bpb(s8)},
                      "gpu_over_ref_t1": sg / s1, "gpu_over_ref_t8": sg / s8}))
    assert sg <= 1.02 * s1 and sg <= 1.02 * s8, (sg, s1, s8)
    # and the archive is lossless through the reference decoder
    R.decoder(g)
    fq = np.fromfile(os.path.join(d, "r.fastq"), dtype=np.uint8).tobytes().split(b"\n")[1::4]
    want = os.path.join(g, "all.dna")
    with open(want, "wb") as f:
        f.write(b"\n".join(fq) + b"\n")
    assert R.sorted_lines_digest(os.path.join(g, "output", "output.dna"), L) == R.sorted_lines_digest(want, L)


@pytest.mark.parametrize("L,K,n,G", [(250, 2, 40000, 500000), (100, 1, 60000, 300000)], ids=["L250_K2", "L100_K1"])
def test_order_preserving_round_trip(workroot, L, K, n, G):
    """configs[3] shape (250 bp, 1 % errors incl. N, -p): GPU stage I + II, then the reference's pack_order ->
    unpack_order -> decoder_preserve -> merge_N must reproduce the FASTQ's read sequence byte for byte."""
    import harc_b200
    d = os.path.join(workroot, "p%d" % L)
    os.makedirs(d, exist_ok=True)
    R.make_genome(os.path.join(d, "g.fa"), G, seed=9)
    R.gen_fastq(os.path.join(d, "g.fa"), os.path.join(d, "r.fastq"), n, L, rc=False, errors=True)
    R.preprocess(os.path.join(d, "r.fastq"), d, L, preserve_order=True)
    ctx = harc_b200.HarcGpu(L, walkers=0, file_sets=K)
    ctx.reorder_dir(d)
    ctx.encode_dir(d)
    packed, tail = ctx.get_packed_order()      # pack_order.cpp on the GPU (harc:111-112)
    ctx.close()
    R.pack_order(d)                            # the reference's own pack_order.out, in place
    assert packed.tobytes() == open(os.path.join(d, "output", "read_order.bin"), "rb").read()
    assert tail.tobytes() == open(os.path.join(d, "output", "read_order.bin.tail"), "rb").read()
    R.decoder_preserve(d, L)
    fq = np.fromfile(os.path.join(d, "r.fastq"), dtype=np.uint8).tobytes().split(b"\n")[1::4]
    got = open(os.path.join(d, "output", "output.dna"), "rb").read()
    assert got == b"\n".join(fq) + b"\n"
