#!/usr/bin/env python3
"""Golden fixture faidx.json: stdout of the reference's OWN bundled samtools 0.1.18 (`samtools faidx <fasta> <region>`,
the call of dump_piece, miR_PREFeR.py:1097-1108) and the .fai index it writes, for seeded FASTA files and a list of
region strings covering the clipping / naming edge cases.  The binary imports a few curses symbols for `tview`; the image
has no libncurses.so.5, so it is started against the no-op stub library built from oracle/ncurses_stub.c.
Build container only:  make -C oracle samtools && python tests/golden/make_golden_faidx.py
"""
import json
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SAMTOOLS = "/root/reference/dependency/Linux/x64/samtools"
LIBDIR = os.path.join(ROOT, "oracle", "_ref", "lib")


def fasta_text(width, seed):
    rng = np.random.default_rng(seed)
    seqs = [("Chr1 first contig", "".join(rng.choice(list("ACGTacgtNRYK"), size=777))),
            ("Chr2", "".join(rng.choice(list("ACGT"), size=width))),            # exactly one full line
            ("scaffold:7 x", "".join(rng.choice(list("ACGT"), size=2 * width + 1))),
            ("tiny", "ACG")]
    out = []
    for name, s in seqs:
        out.append(">" + name)
        out.extend(s[k:k + width] for k in range(0, len(s), width))
    return "\n".join(out) + "\n"


REGIONS = ["Chr1:1-10", "Chr1:100-399", "Chr1:770-777", "Chr1:770-900", "Chr1:777-777", "Chr1:778-800", "Chr1:0-5", "Chr1:5-5",
           "Chr1:10-5", "Chr1:700", "Chr1", "Chr2", "Chr2:1-50", "Chr2:50-51", "Chr2:2-1", "scaffold:7", "scaffold:7:10-20",
           "scaffold:7:100-200", "tiny", "tiny:1-3", "tiny:2-9", "tiny:4-9", "nope", "nope:1-5", "Chr1:1,000-1,010", "Chr1:1-1,0",
           "Chr1:-5-10", "Chr1:3-", "chr1:1-5", "Chr1:60-61", "Chr1:50-50", "Chr1:51-51", "Chr1:49-52"]


def run(fa, region):
    env = dict(os.environ, LD_LIBRARY_PATH=LIBDIR)
    p = subprocess.run([SAMTOOLS, "faidx", fa, region], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env)
    return {"region": region, "stdout": p.stdout.decode(), "rc": p.returncode}


def main():
    cases = []
    for width, seed in ((50, 9), (60, 10), (7, 11)):
        text = fasta_text(width, seed)
        tmp = tempfile.mkdtemp()
        fa = os.path.join(tmp, "g.fa")
        open(fa, "w").write(text)
        res = [run(fa, r) for r in REGIONS]
        cases.append({"fasta": text, "fai": open(fa + ".fai").read(), "queries": res})
    json.dump({"samtools": "0.1.18 (reference bundle, dependency/Linux/x64)", "cases": cases},
              open(os.path.join(HERE, "faidx.json"), "w"), separators=(",", ":"))
    for q in cases[0]["queries"]:
        print(repr(q["region"]), q["rc"], repr(q["stdout"][:70]))


if __name__ == "__main__":
    main()
