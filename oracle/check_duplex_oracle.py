#!/usr/bin/env python3
"""Pin oracle/duplex_oracle.py against the reference's own get_maturestar_info (AST-extracted).
Build container only.  python oracle/check_duplex_oracle.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", "tests", "golden"))
import duplex_oracle as D  # noqa: E402
import ref_extract  # noqa: E402

ns = ref_extract.load()
fx = json.load(open(os.path.join(HERE, "..", "tests", "golden", "stage3.json")))
structs = fx["structures"]
bad = 0
for q in fx["queries"]:
    r = D.maturestar(structs[q["ss"]], q["mature"], q["fold_start"], q["region"][0], q["region"][1], q["strand"])
    want = q["result"] if isinstance(q["result"], str) else tuple(q["result"])
    bad += r != want
print("fixture queries", len(fx["queries"]), "bad", bad)
rng = np.random.default_rng(99)
n = 0
for it in range(60000):
    ss = structs[int(rng.integers(len(structs)))]
    fs = int(rng.integers(1, 60)); rs = int(rng.integers(1, 5000)); re_ = rs + fs + len(ss) + int(rng.integers(0, 30))
    strand = "+-"[int(rng.integers(2))]; mlen = int(rng.integers(15, 27)); l0 = int(rng.integers(-8, len(ss) - mlen + 9))
    m0 = l0 + rs + fs - 1 if strand == "+" else re_ - (l0 + mlen) - fs + 1
    try:
        want = ns["get_maturestar_info"](ss, (m0, m0 + mlen), fs, fs + len(ss), rs, re_, strand)
    except Exception as e:
        want = "EXC_" + type(e).__name__
    try:
        got = D.maturestar(ss, (m0, m0 + mlen), fs, rs, re_, strand)
    except Exception as e:
        got = "EXC_" + type(e).__name__
    if got != want:
        bad += 1
        if bad < 5:
            print("MISMATCH", ss, m0, mlen, fs, rs, re_, strand, want, got)
    n += 1
print("random queries", n, "bad", bad)
sys.exit(1 if bad else 0)
