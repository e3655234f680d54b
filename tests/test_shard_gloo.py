"""CPU, world_size 2 over gloo: the N>1 path = shard by DP work, fold independently, gather in input
order.  The oracle stands in for the GPU fold here (tests may use it); the gather/shard code is the
product code of mir_prefer_b200.shard."""
import os
import sys

import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch.distributed as dist
    import oracle as O
    from corpus import synth_loci
    from mir_prefer_b200.shard import gather_records, lpt_shards
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    seqs = synth_loci(9, 10, (30, 90))
    shards = lpt_shards([len(s) for s in seqs], 40, world)
    mine = shards[rank]
    local = [O.fold(seqs[k], 40)["hits"] for k in mine]
    full = gather_records(mine, local, len(seqs))
    if rank == 0:
        q.put(full)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather_matches_single():
    import oracle as O
    from corpus import synth_loci
    O.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    full = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    seqs = synth_loci(9, 10, (30, 90))
    assert full == [O.fold(s, 40)["hits"] for s in seqs]
