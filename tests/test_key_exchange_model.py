"""CPU: a numpy model of the key exchange of one job on several GPUs (csrc/job.cu): every rank cuts the
(mixed key, id) pairs of its slice by owner with ONE stable pass over the shard bits, the owners receive the ranges in rank
order at the offsets that follow from the count matrix, and a stable sort by key on the owner then gives every shard the
dictionary of one GPU restricted to its key range -- ids ascending inside every bin (reorder.cpp:344-391), which is what
the reference's bin scan order depends on.  (The kernels themselves are checked against the one-GPU dictionary bit for
bit by tests/test_multigpu_gpu.py; this is the ordering argument, made executable.)"""
import numpy as np
import pytest

KEY_MIX = np.uint64(0x9E3779B97F4A7C15)


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_exchange_by_count_matrix_keeps_ids_ascending_inside_bins(world, seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(2000, 6000))
    keys = rng.integers(0, 300, size=n, dtype=np.uint64) * np.uint64(0x0123456789ABCDEF)   # few distinct keys: large bins
    with np.errstate(over="ignore"):
        mixed = keys * KEY_MIX                                                              # common.cuh: key_mix
    ids = np.arange(n, dtype=np.uint32)
    kb = world.bit_length() - 1
    owner_of = lambda t: (t >> np.uint64(64 - kb)).astype(np.int64)
    # slices as multi.slice_ranges cuts them
    base, extra = divmod(n, world)
    cuts = np.cumsum([0] + [base + (1 if r < extra else 0) for r in range(world)])
    # every rank: stable partition of its slice by owner (the one radix pass), its row of the count matrix
    parts, cnt = [], np.zeros((world, world), dtype=np.int64)
    for r in range(world):
        k, v = mixed[cuts[r]:cuts[r + 1]], ids[cuts[r]:cuts[r + 1]]
        o = np.argsort(owner_of(k), kind="stable")
        parts.append((k[o], v[o]))
        cnt[r] = np.bincount(owner_of(k), minlength=world)
    # every owner: the range of rank s lands behind the ranges of the ranks before it (job_push_kernel's offsets)
    for d in range(world):
        rk = np.empty(int(cnt[:, d].sum()), dtype=np.uint64)
        rv = np.empty(len(rk), dtype=np.uint32)
        for s in range(world):
            lo = int(cnt[s, :d].sum())                  # where owner d's range starts in rank s's partitioned pairs
            off = int(cnt[:s, d].sum())                 # pairs of the ranks before s for this owner
            rk[off:off + cnt[s, d]] = parts[s][0][lo:lo + cnt[s, d]]
            rv[off:off + cnt[s, d]] = parts[s][1][lo:lo + cnt[s, d]]
        o = np.argsort(rk, kind="stable")               # the owner's stable radix sort
        sk, sv = rk[o], rv[o]
        # the one-GPU dictionary restricted to this shard
        sel = owner_of(mixed) == d
        o1 = np.argsort(mixed[sel], kind="stable")
        assert np.array_equal(sk, mixed[sel][o1]) and np.array_equal(sv, ids[sel][o1])
        # ids ascending inside every bin
        same = sk[1:] == sk[:-1]
        assert np.all(sv[1:][same] > sv[:-1][same])
