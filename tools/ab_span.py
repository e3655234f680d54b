"""Lean same-box A/B of one workload (no torch): resident-batch step time and the single-lane stage split.
usage: [MIRFOLD_BIG_TILE_MIN_SPAN=..] python tools/ab_span.py WORKLOAD LOCI SPAN [STEPS]  ->  one JSON line"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import mir_prefer_b200 as mp  # noqa: E402
from mir_prefer_b200.fold import FLAG_SERIAL  # noqa: E402

name, nloci, span = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
buf, off = bench.workload_packed(0, nloci, name)
with mp.MirFold() as mf:
    batch = mf.upload(buf, off, span)
    for _ in range(2):
        batch.fold().close()
    ms = []
    for _ in range(steps):
        with batch.fold() as r:
            ms.append(r.stats["ms_device"])
            units, cells, nhits = r.stats["fill_units"], r.stats["cells"], r.nhits
    batch.fold(flags=FLAG_SERIAL).close()
    with batch.fold(flags=FLAG_SERIAL) as r:
        serial = {k: round(r.stats[k], 2) for k in ("ms_fill", "ms_f3", "ms_trace", "ms_device")}
print(json.dumps({"workload": name, "loci": nloci, "span": span, "big_tile_min_span": os.environ.get("MIRFOLD_BIG_TILE_MIN_SPAN", "default"),
                  "ms_per_step": round(sum(ms) / len(ms), 2), "nt_per_s": round(float(off[-1]) / (sum(ms) / len(ms)) * 1e3),
                  "fill_units": int(units), "cells": int(cells), "nhits": int(nhits), "serial": serial}))
