"""Sharding of independent loci across ranks/GPUs by DP work.

The reference shards loci over NUM_OF_CORE processes by locus COUNT (miR_PREFeR.py:1329-1354,
one RNALfold process per shard, MP:3113-3118).  Here shards are balanced by DP cells with greedy
longest-processing-time assignment; there is no data-path collective -- ranks fold their shard
independently and only the (small) results are gathered on the host, in input order.
"""
import numpy as np


def dp_cells(n, L):
    """DP cells RNALfold visits for one locus: sum_{i=1}^{n-4} (min(n, i+min(L,n)) - i - 3)."""
    n = np.asarray(n, dtype=np.int64)
    Ls = np.minimum(L, n)
    full = np.clip(n - Ls, 0, np.maximum(n - 4, 0))          # rows with the full band
    rest = np.maximum(n - 4 - full, 0)                        # rows clipped by the 3' end
    # clipped rows i = full+1 .. n-4 contribute n-i-3 = rest, rest-1, ..., 1
    return np.where(n >= 5, full * (Ls - 3) + rest * (rest + 1) // 2, 0)


def lpt_shards(lengths, L, nshards):
    """Greedy LPT: returns a list of `nshards` sorted index arrays covering range(len(lengths))."""
    work = dp_cells(lengths, L) + 1
    order = np.argsort(-work, kind="stable")
    load = np.zeros(nshards, np.int64)
    out = [[] for _ in range(nshards)]
    for k in order:
        g = int(np.argmin(load))
        out[g].append(int(k))
        load[g] += int(work[k])
    return [np.array(sorted(s), dtype=np.int64) for s in out]


def gather_records(local_idx, local_records, nrec, group=None):
    """All ranks contribute (index, record) pairs; every rank gets the full list in input order.
    Uses torch.distributed.all_gather_object (host-side gather of small results, no NCCL data path)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    bucket = [None] * world
    dist.all_gather_object(bucket, (list(map(int, local_idx)), list(local_records)), group=group)
    out = [None] * nrec
    for idx, recs in bucket:
        for k, r in zip(idx, recs):
            out[k] = r
    return out
