#!/usr/bin/env python3
"""Aggregate an ncu source page by named line ranges of fill.cu.
usage: ncu -i rep --page source --csv --print-source cuda,sass | python tools/ncu_ranges.py name:lo-hi[,lo-hi] ..."""
import collections, csv, sys
ranges = []
FILE = "fill.cu"
for a in sys.argv[1:]:
    if a.startswith("file="):
        FILE = a[5:]; continue
    name, spec = a.split(":")
    for part in spec.split(","):
        lo, hi = part.split("-")
        ranges.append((name, int(lo), int(hi)))
rows = list(csv.reader(sys.stdin))
inst = collections.Counter(); stall = collections.Counter()
h = None; fname = ""
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]; ln = -1; continue
    if r and r[0] == "Line No":
        h = r; continue
    if h is None or len(r) < len(h) - 5 or "Instructions Executed" not in h:
        continue
    ii = h.index("Instructions Executed"); ws = h.index("Warp Stall Sampling (All Samples)")
    if r[0]:
        try:
            ln = int(r[0])
        except ValueError:
            continue
    try:
        v = int(r[ii]); s = int(r[ws])
    except ValueError:
        continue
    key = "other:" + fname
    if fname == FILE:
        key = FILE + ":unassigned"
        for name, lo, hi in ranges:
            if lo <= ln <= hi:
                key = name; break
    inst[key] += v; stall[key] += s
T = sum(inst.values()); S = sum(stall.values())
print("total warp-inst %d stall samples %d" % (T, S))
for k, v in inst.most_common():
    print("%5.1f%% inst %5.1f%% stall  %s" % (100 * v / T, 100 * stall[k] / max(S, 1), k))
