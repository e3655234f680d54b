#!/usr/bin/env python3
"""BASELINE configs[3] shape at reduced size: long 2-10 kb loci folded through ONE context that owns every
GPU of the box (in-process LPT sharding by DP cells, one host thread per device, no collectives)."""
import hashlib
import sys
import time

sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import torch
from mir_prefer_b200.corpus import synth_loci
import mir_prefer_b200 as mp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
seqs = synth_loci(1003, n, "long")
nt = sum(len(s) for s in seqs)
ng = torch.cuda.device_count()


def digest(res):
    h = hashlib.sha256()
    for r in range(0, len(seqs), 97):
        for ss, e, st in res.hits(r):
            h.update(("%s %d %d;" % (ss, e, st)).encode())
        h.update(b"|%d" % res.total(r))
    return h.hexdigest()[:16]


with mp.MirFold(devices=list(range(ng))) as mf:
    buf, off = mf.pack(seqs)
    mf.fold_packed(buf, off, 300).close()
    t0 = time.time()
    with mf.fold_packed(buf, off, 300) as res:
        dt = time.time() - t0
        print("%d GPUs: %d loci, %.1f M nt, %.3f s -> %.1f M nt/s; devices %d, chunks %d, fill units %d, device ms (max over devices) %.1f, digest %s"
              % (ng, n, nt / 1e6, dt, nt / dt / 1e6, res.stats["n_devices"], res.stats["n_chunks"], res.stats["fill_units"],
                 res.stats["ms_device"], digest(res)))
with mp.MirFold(devices=[0]) as mf1:
    sub = seqs[:n // ng]
    b1, o1 = mf1.pack(sub)
    mf1.fold_packed(b1, o1, 300).close()
    t0 = time.time()
    with mf1.fold_packed(b1, o1, 300) as r1:
        dt1 = time.time() - t0
        print("1 GPU, 1/%d of the loci: %.3f s -> %.1f M nt/s" % (ng, dt1, sum(len(s) for s in sub) / dt1 / 1e6))
    if n <= 20000:
        with mf1.fold_packed(buf, off, 300) as rall:
            print("single-device digest %s" % digest(rall))
