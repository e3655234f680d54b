#!/usr/bin/env python3
"""Aggregate an ncu source page (--print-source cuda,sass --csv) by CUDA source line.
usage: ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass | python tools/ncu_lines.py [top]"""
import collections
import csv
import sys

top = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rows = list(csv.reader(sys.stdin))
inst = collections.Counter(); stall = collections.Counter(); tinst = collections.Counter(); text = {}
h = None; cur = None; fname = ""
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        h = r; continue
    if h is None or len(r) < len(h) - 5:
        continue
    ii = h.index("Instructions Executed"); ti = h.index("Thread Instructions Executed"); ws = h.index("Warp Stall Sampling (All Samples)")
    if r[0]:
        cur = (fname, r[0]); text[cur] = r[1]
    try:
        inst[cur] += int(r[ii]); tinst[cur] += int(r[ti]); stall[cur] += int(r[ws])
    except ValueError:
        pass
T = sum(inst.values()); S = sum(stall.values()); TT = sum(tinst.values())
print("total warp-inst %d  thread-inst %d  stall samples %d" % (T, TT, S))
for k, v in inst.most_common(top):
    print("%5.1f%% inst %5.1f%% stall  %s:%s  %s" % (100 * v / T, 100 * stall[k] / max(S, 1), k[0], k[1], text[k].strip()[:100]))
