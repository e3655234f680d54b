#!/usr/bin/env python3
"""Golden fixture for the predict-stage consumer: the reference's OWN check_loci and
filter_next_loci (AST-extracted from /root/reference/miR_PREFeR.py) driven over seeded synthetic loci,
with the samtools-backed expression functions replaced by the deterministic stubs of predict_stub.py.
Build container only:  python tests/golden/make_golden_predict.py  -> tests/golden/predict.json"""
import ast
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import predict_stub as PS  # noqa: E402
import ref_extract  # noqa: E402

WANTED = ["check_loci", "filter_next_loci"]


def load_reference():
    ns = ref_extract.load()
    tree = ast.parse(open(ref_extract.REF).read())
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in WANTED:
            exec(compile(ast.Module(body=[node], type_ignores=[]), ref_extract.REF, "exec"), ns)
    ns["check_expression_new"] = PS.check_expression_stub
    ns["gen_mapinfo_each_sample"] = lambda bam, samples, seqid, s, e: PS.mapinfo_stub(seqid, s, e)
    return ns


def make_loci(seed, nloci):
    """Seeded loci: (region, which, matures) + structure tuples, from the committed golden structures."""
    rng = np.random.default_rng(seed)
    stage1 = json.load(open(os.path.join(HERE, "stage1.json")))
    pool = [s for rec in stage1["synth8.L300"] for s in rec["structures"]]          # [hex energy, start, ss, type]
    pool = [(float.fromhex(e), st, ss, t) for e, st, ss, t in pool]
    aln, ssrecs = [], []
    k = 0
    while k < nloci:
        both = rng.random() < 0.35
        tags = ["L", "R"] if both else ["0"]
        locus = (int(rng.integers(1000, 90000)),)
        locus = (locus[0], locus[0] + int(rng.integers(18, 25)))
        for tag in tags:
            ns_ = int(rng.integers(0, 5))
            structs = [pool[int(rng.integers(len(pool)))] for _ in range(ns_)]
            if rng.random() < 0.5:
                structs.sort(key=lambda s: s[0])
            rs = int(rng.integers(1, 100000))
            span = max([st + len(ss) for _, st, ss, _ in structs] + [60]) + int(rng.integers(0, 30))
            strand = "+" if rng.random() < 0.5 else "-"
            region = ["Chr%d" % int(rng.integers(1, 6)), (rs, rs + span), strand]
            matures = []
            for _ in range(int(rng.integers(0, 5))):
                mlen = int(rng.choice([15, 18, 20, 21, 22, 24, 27], p=[.05, .1, .15, .3, .2, .15, .05]))
                if structs and rng.random() < 0.85:
                    _, st, ss, _ = structs[int(rng.integers(len(structs)))]
                    l0 = int(rng.integers(-3, max(len(ss) - mlen + 4, -2)))
                    m0 = l0 + rs + st - 1 if strand == "+" else rs + span - (l0 + mlen) - st + 1
                else:
                    m0 = rs + int(rng.integers(0, span))
                matures.append((m0, m0 + mlen, strand, int(rng.integers(1, 500))))
            aln.append([region, tag, {}, matures])
            ssrecs.append((tag, "%d-%d" % locus, structs))
        k += 1
    return aln, ssrecs


class FakePickle:
    def __init__(self, items):
        self.items = list(items)

    def load(self, f):
        if not self.items:
            raise EOFError
        return self.items.pop(0)


def run_reference(ns, aln, ssrecs, allow_no_star, output_details):
    ns["cPickle"] = FakePickle([[list(a[0]), a[1], a[2], list(a[3])] for a in aln])
    ns["get_structures_next_extendregion"] = lambda name, minlen: iter(ssrecs)
    out = []
    gen = ns["filter_next_loci"](os.devnull, "unused", "unused.bam", PS.SAMPLES, True, allow_no_star, output_details, 18, 24, 20)
    try:
        for item in gen:
            out.append(PS.canon(item))
    except RuntimeError as e:          # py3: the reference's `raise StopIteration` at EOF (PEP 479)
        assert "StopIteration" in repr(e.__cause__) or "StopIteration" in str(e), e
    return out


def main():
    ns = load_reference()
    cases = []
    for seed, nloci in ((11, 120), (12, 120)):
        aln, ssrecs = make_loci(seed, nloci)
        for allow_no_star in (True, False):
            for details in (True, False):
                cases.append({"seed": seed, "nloci": nloci, "allow_no_star": allow_no_star, "output_details": details,
                              "expected": run_reference(ns, aln, ssrecs, allow_no_star, details)})
    n_mirna = sum(1 for c in cases for x in c["expected"] if isinstance(x, list))
    n_dict = sum(1 for c in cases for x in c["expected"] if isinstance(x, dict))
    print("cases", len(cases), "miRNA lists", n_mirna, "reason dicts", n_dict)
    json.dump({"cases": cases}, open(os.path.join(HERE, "predict.json"), "w"), separators=(",", ":"))


if __name__ == "__main__":
    main()
