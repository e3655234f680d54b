"""CPU, world_size 2 over gloo: the N>1 plan.  Every rank computes the library's own shard plan
(mirfold_plan_shards -- the code mirfold_fold() runs for a context of N devices; it needs no GPU), folds only its
shard, and the per-record results are gathered in input order.  The oracle stands in for the device fold here
(tests may use it); what is under test is that the plan is identical on every rank, covers every record once,
and that shard-wise folding + gather equals folding everything in one place."""
import os
import sys

import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch.distributed as dist
    import oracle as O
    from mir_prefer_b200.corpus import synth_loci
    from mir_prefer_b200.fold import plan_shards
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    seqs = synth_loci(9, 12, (30, 90)) + ["ACG", ""]
    shard_of, load = plan_shards([len(s) for s in seqs], 40, world)
    plans = [None] * world
    dist.all_gather_object(plans, (shard_of.tolist(), load.tolist()))
    assert all(p == plans[0] for p in plans)                       # same plan everywhere
    mine = [k for k, g in enumerate(shard_of.tolist()) if g == rank]
    local = [(k, O.fold(seqs[k], 40)["hits"]) for k in mine]
    parts = [None] * world
    dist.all_gather_object(parts, local)
    if rank == 0:
        full = [None] * len(seqs)
        for part in parts:
            for k, hits in part:
                assert full[k] is None                             # every record exactly once
                full[k] = hits
        q.put((full, load.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather_matches_single():
    import oracle as O
    from mir_prefer_b200.corpus import synth_loci
    O.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    full, load = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    seqs = synth_loci(9, 12, (30, 90)) + ["ACG", ""]
    assert full == [O.fold(s, 40)["hits"] for s in seqs]
    assert max(load) <= 1.35 * (sum(load) / 2)
