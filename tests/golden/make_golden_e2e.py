#!/usr/bin/env python3
"""Golden fixture e2e_mirna.json: the final predicted-miRNA list of the reference's OWN pipeline pieces, composed end to end
(miR_PREFeR.py:2344-2432):

    candidate records (reference-format headers with M: matures, MP:1124-1143)
      -> FASTA shard -> the reference's RNALfold binary, `RNALfold -L 300`              (MP:3064)
      -> the reference's get_structures_next_extendregion (AST-extracted)               (MP:1541-1599)
      -> the reference's filter_next_loci / check_loci / get_maturestar_info            (MP:2206-2432, 1876-1999)
         with the samtools-backed expression functions replaced by predict_stub.py

The fixture stores the records and the canonical outputs; tests feed the same records through libmirfold
(fold + fused stage 1 + 3 on the device) and the host consumer and must reproduce the list exactly.
Build container only:  python tests/golden/make_golden_e2e.py
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import predict_stub as PS  # noqa: E402
from make_golden_predict import FakePickle, load_reference  # noqa: E402
from mir_prefer_b200 import records as R  # noqa: E402
from mir_prefer_b200.corpus import synth_loci  # noqa: E402

RLF = "/root/reference/dependency/Linux/x64/RNALfold"
SPAN = 300


def rnalfold(text):
    return subprocess.run([RLF, "-L", str(SPAN)], input=text.encode(), stdout=subprocess.PIPE, check=True).stdout.decode()


def build_records(seed, nloci):
    """Seeded candidate records.  Folding does not depend on the header, so the loci are folded once to know where
    their hairpins are and the matures are then placed (seeded) on and around the arms of the found structures."""
    rng = np.random.default_rng(seed)
    ns = load_reference()
    seqs = synth_loci(seed, nloci, "arabidopsis")
    protos = []
    k = 0
    while k < len(seqs):
        both = rng.random() < 0.3 and k + 1 < len(seqs)
        for tag in (["L", "R"] if both else ["0"]):
            s = seqs[k]
            strand = "+" if rng.random() < 0.5 else "-"
            rs = int(rng.integers(1, 200000))
            protos.append(dict(seqid="Chr%d" % int(rng.integers(1, 6)), region=(rs, rs + len(s)), strand=strand, tag=tag, seq=s))
            k += 1
    # pass 1: structures of every record (reference binary + reference parser)
    tmp = tempfile.mkdtemp()
    recs0 = [R.LocusRecord(p["seqid"], p["region"], p["strand"], (p["region"][0] + 40, p["region"][0] + 61), p["tag"],
                           [(p["region"][0] + 40, p["region"][0] + 61, p["strand"])], [], p["seq"]) for p in protos]
    fa = os.path.join(tmp, "p1.fa")
    R.write_fasta(recs0, fa)
    out = os.path.join(tmp, "p1.out")
    open(out, "w").write(rnalfold(open(fa).read()))
    structs = [rec[2] for rec in ns["get_structures_next_extendregion"](out, 55, 3)]
    assert len(structs) == len(protos)
    # pass 2: matures
    records = []
    for p, st in zip(protos, structs):
        rs, re_ = p["region"]
        matures = []
        for _ in range(int(rng.integers(0, 4))):
            mlen = int(rng.choice([17, 20, 21, 22, 24, 26], p=[.05, .15, .35, .25, .15, .05]))
            if st and rng.random() < 0.9:
                _e, fs, ss, _t = st[int(rng.integers(len(st)))]
                # on an arm: start a few bases into a run of brackets of one kind
                arm = "(" if rng.random() < 0.5 else ")"
                pos = [i for i, ch in enumerate(ss) if ch == arm]
                l0 = (pos[0] if arm == "(" else pos[len(pos) // 2]) + int(rng.integers(-2, 6)) if pos else 0
                l0 = max(-2, min(l0, len(ss) - mlen + 2))
                m0 = l0 + rs + fs - 1 if p["strand"] == "+" else re_ - (l0 + mlen) - fs + 1
            else:
                m0 = rs + int(rng.integers(0, max(1, re_ - rs - mlen)))
            matures.append((m0, m0 + mlen, p["strand"], int(rng.integers(1, 800))))
        locus = (matures[0][0], matures[0][1]) if matures else (rs + 40, rs + 61)
        records.append(R.LocusRecord(p["seqid"], p["region"], p["strand"], locus, p["tag"],
                                     [(locus[0], locus[1], p["strand"])], matures, p["seq"]))
    return records, tmp


def reference_pipeline(ns, records, tmp, allow_no_star, details):
    fa = os.path.join(tmp, "shard.fa")
    R.write_fasta(records, fa)
    out = os.path.join(tmp, "shard_rnalfoldoutput_0")
    open(out, "w").write(rnalfold(open(fa).read()))
    ns["cPickle"] = FakePickle([[[r.seqid, r.region, r.strand], r.tag, {}, list(r.matures)] for r in records])
    res = []
    gen = ns["filter_next_loci"](os.devnull, out, "unused.bam", PS.SAMPLES, True, allow_no_star, details, 18, 24, 20, minlen=55)   # the value run_predict passes (MP:3528)
    try:
        for item in gen:
            res.append(PS.canon(item))
    except RuntimeError as e:          # py3: the reference's `raise StopIteration` at EOF (PEP 479)
        assert "StopIteration" in repr(e.__cause__) or "StopIteration" in str(e), e
    return res


def main():
    ns = load_reference()
    cases = []
    for seed, nloci in ((2002, 140), (2003, 140)):
        records, tmp = build_records(seed, nloci)
        expected = {}
        for allow_no_star in (True, False):
            for details in (True, False):
                expected["%d%d" % (allow_no_star, details)] = reference_pipeline(ns, records, tmp, allow_no_star, details)
        n_mirna = sum(isinstance(x, list) for x in expected["11"])
        print("seed", seed, "records", len(records), "with matures", sum(bool(r.matures) for r in records), "miRNA entries", n_mirna,
              "reason dicts", sum(isinstance(x, dict) for x in expected["11"]))
        cases.append({"seed": seed, "records": [[r.seqid, list(r.region), r.strand, list(r.locus), r.tag, [list(p) for p in r.peaks],
                                                 [list(m) for m in r.matures], r.seq] for r in records], "expected": expected})
    json.dump({"span": SPAN, "cases": cases}, open(os.path.join(HERE, "e2e_mirna.json"), "w"), separators=(",", ":"))


if __name__ == "__main__":
    main()
