#!/usr/bin/env python3
"""Pin the CPU oracle against the reference's own RNALfold binary (oracle/_ref/RNALfold).

TEST INFRASTRUCTURE.  Usage: python oracle/check_oracle.py [--quick]
Compares byte-for-byte the stdout of oracle/_build/lfold_oracle and of RNALfold 1.8.5 on
(1) the three sha256 golden pins of SURVEY.md Appendix C, (2) randomized corpora over lengths,
spans and alphabets.
"""
import hashlib
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from mir_prefer_b200.corpus import lcg_records, records_to_fasta, synth_loci  # noqa: E402

ORACLE = os.path.join(HERE, "_build", "lfold_oracle")
RLF = os.path.join(HERE, "_ref", "RNALfold")

PINS = [  # seed, N, lo, span, L, sha256(stdout)
    (1, 1000, 50, 351, 300, "f0b8631476ab6341e1ab5f25b5eeec8314cc46589dc42bd8934f16dc96a3ed8c"),
    (2, 200, 300, 301, 300, "8cedff076c964626787f34072ee1046913a985f6dac9f888534f2d8ad2163e09"),
    (3, 300, 20, 200, 40, "af12c7938fbfcf4f5af72320d21706dfbe1596d6edba38d22f91b29d8d1149c9"),
]


def run(binary, text, L):
    return subprocess.run([binary, "-L", str(L)], input=text.encode(), stdout=subprocess.PIPE, check=True).stdout


def compare(name, text, L, use_rlf=True, sha=None):
    mine = run(ORACLE, text, L)
    ok = True
    if sha is not None:
        ok &= hashlib.sha256(mine).hexdigest() == sha
    if use_rlf and os.path.exists(RLF):
        ref = run(RLF, text, L)
        if ref != mine:
            ok = False
            a, b = ref.decode().split("\n"), mine.decode().split("\n")
            for k, (x, y) in enumerate(zip(a, b)):
                if x != y:
                    print("first diff line", k, "\n ref:", x[:200], "\n ours:", y[:200])
                    break
            else:
                print("length differs", len(a), len(b))
    print("%-40s L=%-4d %s" % (name, L, "OK" if ok else "MISMATCH"))
    return ok


def main():
    quick = "--quick" in sys.argv
    ok = True
    for seed, n, lo, span, L, sha in PINS:
        if quick and seed != 3:
            continue
        ok &= compare("pin seed=%d" % seed, records_to_fasta(lcg_records(seed, n, lo, span)), L, use_rlf=not quick, sha=sha)
    if not quick:
        rng = np.random.default_rng(7)
        for L in (20, 25, 30, 40, 50, 60, 61, 150, 300, 500):
            recs = []
            for k in range(120):
                n = int(rng.integers(1, 4 * L if L < 100 else L + 200))
                fam = k % 5
                if fam == 0:
                    s = "".join(rng.choice(list("ACGU"), size=n))
                elif fam == 1:
                    s = "".join(rng.choice(list("GC"), size=n))
                elif fam == 2:
                    s = "".join(rng.choice(list("AU"), size=n))
                elif fam == 3:
                    s = "".join(rng.choice(list("ACGUTacgutNnKXIRYkxi"), size=n))
                else:
                    s = "".join(rng.choice(list("ACGT"), size=n, p=[.15, .35, .35, .15]))
                recs.append(("q%d" % k, s))
            ok &= compare("random families", records_to_fasta(recs), L)
        loci = synth_loci(1001, 40, "parity")
        ok &= compare("synth parity sample", records_to_fasta([("s%d" % k, s) for k, s in enumerate(loci)]), 300)
        loci = synth_loci(1004, 30, "sweep")
        for L in (150, 500):
            ok &= compare("synth sweep sample", records_to_fasta([("s%d" % k, s) for k, s in enumerate(loci)]), L)
        ok &= compare("long", records_to_fasta([("l0", synth_loci(1003, 1, "long")[0])]), 300)
        edge = ">e\n\n>e2\nACGU\n>e3\nACGUA\n*c\nGGGGAAAACCCC trailing tokens\n>e4\n  GGGGGTTTTCCCCCAAAAGGGGGTTTTCCCCC\n"
        ok &= compare("edge records", edge, 30)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
