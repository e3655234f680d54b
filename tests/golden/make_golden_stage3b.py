#!/usr/bin/env python3
"""Golden fixture stage3b.json: get_maturestar_info() verdicts of the reference's OWN function (AST-extracted,
ref_extract.py) for the rare branches stage3.json does not reach:
  FAIL_STRUCTURE_MATCHED_BASES            (an unmatched ')' anywhere in the structure, MP:1890-1891)
  FAIL_STRUCTURE_TOO_MANY_BULGE_OR_LOOP   (more than five small loops/bulges in the duplex, MP:1850-1851)
  FAIL_STRUCTURE_TOTAL_LOOP_SIZE_LARGER_THAN_5, FAIL_STRUCTURE_NUM_BULGE_MORE_THAN_2, FAIL_STRUCTURE_MAX_BULGE_LARGE_THAN_2
  "EXCEPTION"                              where the reference raises (unmatched '(' reached through dict_bp[...])
Run in the build container only:  python tests/golden/make_golden_stage3b.py
"""
import json
import os
import sys
from collections import Counter

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_extract  # noqa: E402


def imperfect_hairpin(rng, kinds, stem_lo=3, stem_hi=6):
    """A stem-loop whose 5' arm is cut by the given irregularities: 'L1'/'L2'/'L3' = symmetric loop of that size,
    'B1'/'B2'/'B3' = bulge of that size (side chosen at random); stems of 3..5 pairs in between."""
    left, right = [], []
    for kind in kinds:
        stem = int(rng.integers(stem_lo, stem_hi))
        left.append("(" * stem); right.append(")" * stem)
        k = int(kind[1])
        if kind[0] == "L":
            left.append("." * k); right.append("." * k)
        elif rng.random() < 0.5:
            left.append("." * k); right.append("")
        else:
            left.append(""); right.append("." * k)
    ss = "." * int(rng.integers(0, 3)) + "".join(left) + "((((" + "." * int(rng.integers(4, 9)) + "))))" + "".join(reversed(right)) + "." * int(rng.integers(0, 3))
    return ss


def main():
    ns = ref_extract.load()
    rng = np.random.default_rng(20261017)
    structs, queries = [], []

    def ask(ss, l0, mlen, strand, fold_start):
        rs = int(rng.integers(1, 100000))
        re_ = rs + fold_start + len(ss) + int(rng.integers(0, 40))
        m0 = l0 + rs + fold_start - 1 if strand == "+" else re_ - (l0 + mlen) - fold_start + 1
        try:
            res = ns["get_maturestar_info"](ss, (m0, m0 + mlen), fold_start, fold_start + len(ss), rs, re_, strand)
            res = res if isinstance(res, str) else list(res)
        except (KeyError, IndexError):
            res = "EXCEPTION"
        if ss not in structs:
            structs.append(ss)
        queries.append({"ss": structs.index(ss), "mature": [m0, m0 + mlen], "fold_start": fold_start, "region": [rs, re_],
                        "strand": strand, "result": res})
        return res

    recipes = {
        "too_many": lambda: [["L1", "B1", "L2", "B2"][int(rng.integers(4))] for _ in range(int(rng.integers(6, 9)))],
        "total_loop": lambda: [["L2", "L3", "L1"][int(rng.integers(3))] for _ in range(int(rng.integers(3, 6)))],
        "num_bulge": lambda: [["B1", "B2"][int(rng.integers(2))] for _ in range(int(rng.integers(3, 6)))],
        "max_bulge": lambda: [["B3", "L1", "B1"][int(rng.integers(3))] for _ in range(int(rng.integers(2, 5)))],
    }
    for name, mk in recipes.items():
        for _ in range(60):
            ss = imperfect_hairpin(rng, mk())
            fold_start = int(rng.integers(1, 40))
            for strand in "+-":
                # matures covering most of an arm so that many irregularities fall inside the duplex
                mlen = int(rng.integers(20, 25))
                first_open, last_close = ss.find("("), ss.rfind(")")
                for l0 in (first_open, first_open + 2, last_close - mlen + 1, last_close - mlen - 1):
                    if 0 <= l0 and l0 + mlen <= len(ss):
                        ask(ss, l0, mlen, strand, fold_start)
    # more than five irregularities inside a <= 24-nt mature need two-pair stems
    for _ in range(80):
        ss = imperfect_hairpin(rng, [["L1", "B1"][int(rng.integers(2))] for _ in range(int(rng.integers(6, 9)))], 2, 3)
        fold_start = int(rng.integers(1, 40))
        for strand in "+-":
            mlen = int(rng.integers(22, 25))
            first_open, last_close = ss.find("("), ss.rfind(")")
            for l0 in (first_open, first_open + 1, last_close - mlen + 1, last_close - mlen):
                if 0 <= l0 and l0 + mlen <= len(ss):
                    ask(ss, l0, mlen, strand, fold_start)
    # unbalanced structures: pieces cut out of larger ones (what filter_ss never produces, but the C ABI accepts)
    for _ in range(80):
        ss = imperfect_hairpin(rng, ["L1"] * int(rng.integers(1, 4)))
        cut = int(rng.integers(1, 8))
        variant = [ss[cut:], ss[:-cut], ss + ")" * cut, "(" * cut + ss][int(rng.integers(4))]
        mlen = int(rng.integers(19, 24))
        for strand in "+-":
            l0 = int(rng.integers(0, max(1, len(variant) - mlen)))
            ask(variant, l0, mlen, strand, int(rng.integers(1, 30)))
            ask(variant, max(0, variant.find("(")), mlen, strand, int(rng.integers(1, 30)))
    json.dump({"structures": structs, "queries": queries}, open(os.path.join(HERE, "stage3b.json"), "w"), separators=(",", ":"))
    print(Counter(q["result"] if isinstance(q["result"], str) else "PASS" for q in queries))


if __name__ == "__main__":
    main()
