#!/usr/bin/env python3
"""BASELINE configs[3] at scale: N loci of 2-10 kb (log-uniform) folded through ONE context that owns
every GPU of the box, results STREAMED chunk by chunk (mirfold_fold_stream) so that the ~20 B/nt of hit text never sits
in host memory; parity on a seeded sample of the loci against the reference's own RNALfold (oracle/_ref), which folds
the sample on the host cores while the GPUs work.

    python tools/long_scale.py [nloci=200000] [sample=200] [seed=1003]
"""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import mir_prefer_b200 as mp  # noqa: E402
from mir_prefer_b200.fold import format_record, plan_shards  # noqa: E402

nloci = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
nsample = int(sys.argv[2]) if len(sys.argv) > 2 else 200
seed = int(sys.argv[3]) if len(sys.argv) > 3 else 1003
SPAN = 300

t0 = time.time()
rng = np.random.Generator(np.random.PCG64(seed))
lens = np.exp(rng.uniform(np.log(2000), np.log(10000), size=nloci)).astype(np.int64)
off = np.zeros(nloci + 1, np.uint64)
off[1:] = np.cumsum(lens, dtype=np.uint64)
nt = int(off[-1])
lut = np.empty(256, np.uint8)           # A 0.30, C 0.20, G 0.20, T 0.30 (GC 0.40), as the corpus generator
lut[:77], lut[77:128], lut[128:179], lut[179:] = ord("A"), ord("C"), ord("G"), ord("T")
buf = np.empty(nt, np.uint8)
step = 1 << 27


def gen(a):      # one generator per block (seeded by the block index): blocks are drawn on all host cores
    b = min(nt, a + step)
    g = np.random.Generator(np.random.PCG64([seed, a // step]))
    buf[a:b] = lut[g.integers(0, 256, size=b - a, dtype=np.uint8)]


from concurrent.futures import ThreadPoolExecutor  # noqa: E402
with ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 1)) as ex:
    list(ex.map(gen, range(0, nt, step)))
t_gen = time.time() - t0
sample = np.sort(np.random.Generator(np.random.PCG64(seed + 1)).choice(nloci, size=min(nsample, nloci), replace=False))

# the reference on the sample, one RNALfold process per host core, started now and collected after the GPU fold
rlf = os.path.join(ROOT, "oracle", "_ref", "RNALfold")
have_ref = os.path.exists(rlf) and os.access(rlf, os.X_OK)
procs, tmp = [], tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
cores = os.cpu_count() or 1
if have_ref:
    P = max(1, min(cores, len(sample)))
    for k in range(P):
        part = sample[k::P]
        fn = os.path.join(tmp, "s%d.fa" % k)
        with open(fn, "w") as f:
            for r in part:
                f.write(">r%d\n%s\n" % (r, buf[int(off[r]):int(off[r + 1])].tobytes().decode()))
        procs.append((part, subprocess.Popen([rlf, "-L", str(SPAN)], stdin=open(fn), stdout=open(fn + ".out", "w")), fn))
t_ref0 = time.time()

ng = torch.cuda.device_count()
_, shard_cells = plan_shards(lens, SPAN, ng)
want = set(int(r) for r in sample)
got, tot = {}, {"chunks": 0, "records": 0, "hits": 0, "ss_bytes": 0, "mfe_sum": 0}


def on_chunk(ch):
    tot["chunks"] += 1
    tot["records"] += len(ch.records)
    tot["hits"] += ch.nhits
    tot["ss_bytes"] += ch.ss_bytes
    tot["mfe_sum"] += int(ch.total_mfe_dcal.astype(np.int64).sum())
    idx = np.flatnonzero(np.isin(ch.records, sample))
    for k in idx.tolist():
        got[int(ch.records[k])] = (ch.hits(k), int(ch.total_mfe_dcal[k]))


with mp.MirFold(devices=list(range(ng))) as mf:
    warm = min(nloci, 2000 * ng)
    mf.fold_stream(buf[:int(off[warm])], off[:warm + 1], SPAN, lambda ch: None)     # warm-up: allocations, module load
    for k in tot:
        tot[k] = 0
    t0 = time.time()
    st = mf.fold_stream(buf, off, SPAN, on_chunk)
    dt = time.time() - t0

res = {"workload": "long-%d (BASELINE configs[3] law): %d loci, 2-10 kb log-uniform, GC 0.40, L=%d, seed %d" % (nloci, nloci, SPAN, seed),
       "n_gpus": ng, "nt": nt, "seconds": dt, "nt_per_s": nt / dt, "dp_cells": int(st["cells"]), "dp_cells_per_s": st["cells"] / dt,
       "chunks": tot["chunks"], "device_chunks": int(st["n_chunks"]), "fill_units": int(st["fill_units"]), "records_seen": tot["records"],
       "hits": tot["hits"], "ss_bytes_streamed": tot["ss_bytes"], "total_mfe_dcal_sum": tot["mfe_sum"],
       "d2h_bytes": int(st["d2h_bytes"]), "h2d_bytes": int(st["h2d_bytes"]), "ms_device_max": st["ms_device"],
       "lpt_imbalance": float(shard_cells.max()) / float(shard_cells.mean()), "generate_seconds": t_gen,
       "result_mode": "streamed per chunk (mirfold_fold_stream); nothing but the sampled records was kept"}
assert tot["records"] == nloci, (tot["records"], nloci)
if have_ref:
    bad = 0
    for part, p, fn in procs:
        assert p.wait() == 0
        text = open(fn + ".out").read()
        mine = "".join(">r%d\n%s" % (r, format_record(buf[int(off[r]):int(off[r + 1])].tobytes().decode(), got[int(r)][0], got[int(r)][1]))
                       for r in part)
        bad += text != mine
    res["parity_sample"] = {"loci": int(len(sample)), "nt": int(lens[sample].sum()), "byte_identical_shards": len(procs) - bad,
                            "shards": len(procs), "reference": "oracle/_ref/RNALfold -L %d on %d host cores" % (SPAN, cores),
                            "reference_seconds": time.time() - t_ref0}
    assert bad == 0, "parity broken on the sample"
print(json.dumps(res))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "r02_long_scale_%d_%dgpu.json" % (nloci, ng)), "w").write(json.dumps(res) + "\n")
