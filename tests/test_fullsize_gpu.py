"""GPU: BASELINE.json configs[1] at FULL size (35 M x 100 bp, 1 % substitutions incl. N, 50 Mbp genome) through the C ABI
with host buffers, checked through size-independent properties: every read id exactly once, temp.dna = reads gathered by
order (reverse-complemented where flagged), well-formed streams, and stream sizes that add up.  (Bit-exactness and the
decoder round trip are covered at oracle-sized inputs in the other GPU tests.)"""
import os
import sys

import numpy as np
import pytest

import harness as H

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(H.ROOT, "tools"))


def test_config1_full_size_properties():
    import harc_b200
    import workload as W
    L, n, G = 100, 35_000_000, 50_000_000
    w = W.make(n, L, G, rc=False, errors=True, seed=3)
    n_clean, n_N = w["n_clean"], w["n_N"]
    ctx = harc_b200.HarcGpu(L, file_sets=2)
    ctx.load_reads(w["clean"], n_clean)
    m, s, u = ctx.reorder()
    res = ctx.get_reorder()
    assert m + s == n_clean
    cnt = np.bincount(np.concatenate([res["order"], res["order_s"]]), minlength=n_clean)
    assert cnt.size == n_clean and cnt.min() == 1 and cnt.max() == 1          # every id exactly once
    head = res["flag"] == ord("0")
    assert head[0] and np.all(res["pos"][head] == L) and np.all(res["pos"][~head] < L // 2)
    assert int(head.sum()) + s == u
    # temp.dna on a sample of 200k stream positions: read[order[i]], reverse-complemented where flagged
    dna, _ = ctx.get_reordered_reads()
    dna = dna.reshape(-1, L + 1)
    clean = w["clean"].reshape(-1, L + 1)
    idx = np.random.default_rng(1).integers(0, m, 200_000)
    comp = np.zeros(256, np.uint8)
    for a, b in zip(b"ACGT\n", b"TGCA\n"):
        comp[a] = b
    want = clean[res["order"][idx]].copy()
    r = res["rev"][idx] == ord("r")
    want[r, :L] = comp[want[r, :L][:, ::-1]]
    assert np.array_equal(dna[idx], want)
    # stage II
    ctx.load_pool(None, None, w["withN"])
    es = ctx.encode()
    sets = [ctx.get_set(k) for k in range(2)]
    g = ctx.get_globals()
    ctx.close()
    reads_in_sets = sum(len(x["pos"]) for x in sets)
    u_s = (4 * len(g["singleton"]) + len(g["singleton_tail"])) // L
    u_N = len(g["input_N"]) // (L + 1)
    assert reads_in_sets + u_s + u_N == n                                     # every read is in exactly one output
    assert es.aligned_singletons == s - u_s and es.aligned_N == n_N - u_N
    assert len(g["order"]) == n_clean and len(g["order_N"]) == n_N
    assert np.array_equal(np.sort(g["order_N"]), np.arange(n_N, dtype=np.uint32))
    assert np.bincount(g["order"], minlength=n_clean).max() == 1
    for x in sets:
        k = len(x["pos"])
        assert len(x["rev"]) * 8 + len(x["rev_tail"]) == k                    # one orientation flag per read
        assert int((x["noise"] == ord("\n")).sum()) == k                      # one noise line per read
        assert len(x["noisepos"]) == len(x["noise"]) - k                      # one position byte per noise character
        heads = int((x["pos"] == L).sum())
        assert 4 * len(x["seq"]) + len(x["seq_tail"]) >= heads * L            # at least L consensus bases per contig
