"""GPU: the library's own radix sort (sort.cu) against numpy's stable sort: all 64 bits, bit ranges, the dictionary
build's route (top half + run fix-up) incl. crafted runs that share their top half, and the fallback for long runs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _check(ctx, keys, mode=0, begin_bit=0, end_bit=64):
    vals = np.arange(keys.size, dtype=np.uint32)
    k, v = ctx.debug_sort(keys, vals, mode, begin_bit, end_bit)
    if mode == 2:
        mask = np.uint64(((1 << (end_bit - begin_bit)) - 1) << begin_bit) if end_bit - begin_bit < 64 else np.uint64(0xFFFFFFFFFFFFFFFF)
        order = np.argsort(keys & mask, kind="stable")
    else:
        order = np.argsort(keys, kind="stable")
    assert np.array_equal(k, keys[order])
    assert np.array_equal(v, order.astype(np.uint32))  # stable: equal keys keep their input order


def test_radix_sort_matches_stable_argsort():
    import harc_b200
    ctx = harc_b200.HarcGpu(100)
    rng = np.random.default_rng(7)
    for n in (0, 1, 5, 4095, 4096, 4097, 300001):
        keys = rng.integers(0, 2 ** 64, size=n, dtype=np.uint64)
        _check(ctx, keys, 0)
        _check(ctx, keys, 1)
    # few distinct keys (large bins), small keys (only low digits differ)
    _check(ctx, rng.integers(0, 50, size=100000, dtype=np.uint64), 0)
    _check(ctx, rng.integers(0, 50, size=100000, dtype=np.uint64) << np.uint64(40), 1)
    # bit ranges, including one that is not a multiple of eight bits wide
    keys = rng.integers(0, 2 ** 64, size=70000, dtype=np.uint64)
    _check(ctx, keys, 2, 0, 22)
    _check(ctx, keys, 2, 32, 45)
    _check(ctx, keys, 2, 8, 64)
    ctx.close()


def test_top_half_route_orders_runs_that_share_their_top_half():
    import harc_b200
    ctx = harc_b200.HarcGpu(100)
    rng = np.random.default_rng(11)
    base = rng.integers(0, 2 ** 64, size=200000, dtype=np.uint64)
    # runs of 2..32 different keys with one top half, some of them with duplicates inside, scattered through the input
    extra = []
    for r in range(400):
        top = np.uint64(int(rng.integers(0, 2 ** 32))) << np.uint64(32)
        length = int(rng.integers(2, 33))
        low = rng.integers(0, 2 ** 32 if r % 3 else 4, size=length, dtype=np.uint64)
        extra.append(top | low)
    keys = np.concatenate([base] + extra)
    rng.shuffle(keys)
    _check(ctx, keys, 1)
    # a run longer than the fix-up handles: the full sort takes over
    top = np.uint64(0x12345678) << np.uint64(32)
    long_run = top | rng.integers(0, 2 ** 32, size=100, dtype=np.uint64)
    keys = np.concatenate([base[:50000], long_run])
    rng.shuffle(keys)
    _check(ctx, keys, 1)
    ctx.close()


def test_long_bins_of_equal_keys_stay_on_the_fast_path():
    """A dictionary bin of more than 32 reads (equal keys: certain on repetitive genomes) is already in order after the four
    passes over the top half: the fix-up must leave it alone instead of falling back to the full sort."""
    import harc_b200
    ctx = harc_b200.HarcGpu(100)
    rng = np.random.default_rng(12)
    base = rng.integers(0, 2 ** 64, size=100000, dtype=np.uint64)
    bins = [np.full(n, k, dtype=np.uint64) for n, k in ((33, 0xAAAA000011112222), (1000, 0x0123456789ABCDEF), (50000, 0))]
    keys = np.concatenate([base] + bins)
    rng.shuffle(keys)
    l0 = harc_b200.launch_count()
    _check(ctx, keys, 1)
    fast = harc_b200.launch_count() - l0
    l0 = harc_b200.launch_count()
    _check(ctx, keys[: len(base)], 1)
    assert fast == harc_b200.launch_count() - l0, "long bins of equal keys must not trigger the eight-pass fallback"
    # a long bin that shares its top half with a few smaller keys is out of order after the four passes: the fallback sorts it
    mixed = np.concatenate([np.full(400, 0xBEEF000000000007, dtype=np.uint64), np.full(3, 0xBEEF000000000003, dtype=np.uint64)])
    keys = np.concatenate([base, mixed])
    rng.shuffle(keys)
    _check(ctx, keys, 1)
    ctx.close()
