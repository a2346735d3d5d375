"""GPU parity of stage I (reorder.cpp) through the C ABI, against the oracle on the same seeded inputs."""
import os

import numpy as np
import pytest

import harness as H
import refrun as R

pytestmark = pytest.mark.gpu

CASES = [
    # name, reads, L, genome, rc, errors
    ("s100", 20000, 100, 200000, False, True),
    ("s100rc", 20000, 100, 200000, True, False),
    ("s250", 8000, 250, 150000, False, True),
    ("s63", 20000, 63, 100000, True, True),
    ("s36", 20000, 36, 60000, True, True),
    ("rep100", 150000, 100, "repeats", True, True),   # poly-A, tandem repeats, duplications: bins beyond maxsearch in both stages
]


@pytest.fixture(scope="module")
def gpu():
    import harc_b200
    return harc_b200


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_dictionary_bit_exact(gpu, workroot, case):
    """Dictionary contents (keys, bin sizes, ids inside bins) == the reference's own constructdictionary (dictdump
    wrapper around the unmodified reorder.cpp)."""
    name, n, L, G, rc, err = case
    d = H.dataset(workroot, case, seed=11)
    dump = os.path.join(d, "dict1.bin")
    R.dictdump(d, L, dump)
    raw = np.fromfile(dump, dtype=np.uint8)
    g = gpu.HarcGpu(L)
    ascii_ = np.fromfile(os.path.join(d, "output", "input_clean.dna"), dtype=np.uint8)
    g.load_reads(ascii_)
    g.build_dicts()
    off = 4
    assert int(raw[:4].view(np.uint32)[0]) == 2
    for l in range(2):
        nk, nid = raw[off:off + 8].view(np.uint32)
        off += 8
        rec = raw[off:off + 12 * int(nk)].view(np.dtype([("k", "<u8"), ("c", "<u4")]))
        off += 12 * int(nk)
        ids = raw[off:off + 4 * int(nid)].view(np.uint32)
        off += 4 * int(nid)
        keys, counts, gids = g.dump_dict(1, l)
        assert len(keys) == nk
        assert np.array_equal(keys, rec["k"])
        assert np.array_equal(counts, rec["c"])
        assert np.array_equal(gids, ids)
    g.close()


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_reorder_one_walker_bit_exact(gpu, workroot, case):
    """With one walker the GPU chain walk must reproduce the reference at num_thr=1 byte for byte (all seven files)."""
    name, n, L, G, rc, err = case
    d = H.dataset(workroot, case, seed=11)
    o = H.clone(d, d + ".oracle")
    H.oracle_reorder(o, L, 1)
    g = H.clone(d, d + ".gpu")
    ctx = gpu.HarcGpu(L, walkers=1)
    ctx.reorder_dir(g)
    ctx.close()
    assert H.same_files(o, g, H.STAGE1_FILES) == []


def _check_invariants(d, L, res, dna, sdna, heads_direct=True):
    clean = H.read_lines(os.path.join(d, "output", "input_clean.dna"), L)
    n = clean.shape[0]
    allid = np.concatenate([res["order"], res["order_s"]])
    assert allid.size == n and np.array_equal(np.sort(allid), np.arange(n, dtype=np.uint32))
    m = res["order"].size
    assert set(np.unique(res["flag"]).tolist()) <= {ord("0"), ord("1")}
    assert set(np.unique(res["rev"]).tolist()) <= {ord("d"), ord("r")}
    head = res["flag"] == ord("0")
    assert np.all(res["pos"][head] == L) and np.all(res["pos"][~head] < L // 2)
    if heads_direct:  # reference behaviour; with the left extension a chain may start with a reverse-complemented read
        assert np.all(res["rev"][head] == ord("d"))
    if m:
        assert head[0]
    # temp.dna = reads gathered by order, reverse-complemented where flagged
    comp = np.zeros(256, np.uint8)
    for a, b in zip(b"ACGT\n", b"TGCA\n"):
        comp[a] = b
    want = clean[res["order"]].copy()
    r = res["rev"] == ord("r")
    want[r, :L] = comp[want[r, :L][:, ::-1]]
    assert np.array_equal(dna.reshape(-1, L + 1), want)
    assert np.array_equal(sdna.reshape(-1, L + 1), clean[res["order_s"]])


@pytest.mark.parametrize("extend", [-1, 1])
@pytest.mark.parametrize("walkers", [1, 7, 64, 0])
def test_reorder_many_walkers_invariants(gpu, workroot, walkers, extend):
    """Many concurrent walkers, with and without the left extension of new chains: every read exactly once, streams
    well formed, gathered reads consistent."""
    L = 100
    d = H.make_dataset(workroot, "m100", 60000, L, 400000, True, True, seed=5)
    ctx = gpu.HarcGpu(L, walkers=walkers, extend=extend)
    ctx.load_reads(np.fromfile(os.path.join(d, "output", "input_clean.dna"), dtype=np.uint8))
    m, s, u = ctx.reorder()
    res = ctx.get_reorder()
    dna, sdna = ctx.get_reordered_reads()
    cnt = ctx.counters()
    ctx.close()
    _check_invariants(d, L, res, dna, sdna, heads_direct=extend < 0)
    assert cnt["steps"] + cnt["harvested"] >= m and cnt["restarts"] == u
    # chain heads + singletons = restarts
    assert int((res["flag"] == ord("0")).sum()) + s == u


def test_empty_and_tiny_inputs(gpu):
    """Edge cases of reorder.cpp:120,482-497: zero reads, one read, fewer reads than walkers."""
    L = 100
    rng = np.random.default_rng(3)
    for n in (0, 1, 3):
        lines = np.full((n, L + 1), ord("\n"), np.uint8)
        if n:
            lines[:, :L] = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, (n, L))]
        ctx = gpu.HarcGpu(L, walkers=16)
        ctx.load_reads(lines.reshape(-1), n)
        m, s, u = ctx.reorder()
        res = ctx.get_reorder()
        ctx.close()
        assert m + s == n
        assert np.array_equal(np.sort(np.concatenate([res["order"], res["order_s"]])), np.arange(n, dtype=np.uint32))
