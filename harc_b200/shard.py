"""Host-side multi-GPU logic (one process per GPU, torch.distributed for the plumbing).

Round 1 shards by INDEPENDENT read sets: rank r compresses read set r (its own FASTQ / lane) with no data-path
collective, exactly like running `harc -c` once per file; the only exchange is the max-over-ranks timing and the
sum of the read counts.  `split_fastq_ranges` is the planned single-job split (contiguous ranges of one FASTQ).
"""


def split_fastq_ranges(n_reads, world):
    """Contiguous [begin, end) read ranges, one per rank, sizes differing by at most one."""
    base, extra = divmod(int(n_reads), int(world))
    out, b = [], 0
    for r in range(world):
        e = b + base + (1 if r < extra else 0)
        out.append((b, e))
        b = e
    return out


def aggregate_throughput(local_reads, local_ms, dist=None, device=None):
    """Whole-job Mreads/s = reads of all ranks / slowest rank's time (never a sum of per-rank rates)."""
    import torch
    if dist is None or not dist.is_initialized():
        return local_reads / (local_ms / 1000.0) / 1e6, local_ms, local_reads
    t = torch.tensor([float(local_ms)], dtype=torch.float64, device=device)
    n = torch.tensor([float(local_reads)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    return float(n.item()) / (float(t.item()) / 1000.0) / 1e6, float(t.item()), float(n.item())
