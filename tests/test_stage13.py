"""Stages (1) and (3): host parser/classifier (CPU) and device duplex checks (GPU) against fixtures
produced by the reference's own functions (tests/golden/make_golden_stage13.py)."""
import json
import os

import pytest

from conftest import GOLDEN
from mir_prefer_b200 import structures as S


@pytest.fixture(scope="module")
def stage1():
    return json.load(open(os.path.join(GOLDEN, "stage1.json")))


@pytest.fixture(scope="module")
def stage3():
    return json.load(open(os.path.join(GOLDEN, "stage3.json")))


def test_file_parser_matches_reference_tuples(stage1):
    got = list(S.get_structures_next_extendregion(os.path.join(GOLDEN, "synth8.L300.out"), 55, 3))
    want = stage1["synth8.L300"]
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert g[0] == w["which"] and g[1] == w["peak"]
        assert [[float.hex(e), s, ss, t] for e, s, ss, t in g[2]] == w["structures"]


def test_classifier_matches_reference(stage1):
    n = 0
    for item in stage1["classifier"]:
        ss = item["ss"]
        assert S.is_stem_loop(ss, 3) == item["stem_loop"]
        if "filter_ss" in item:
            sub, tot = S.filter_ss(ss)
            assert [[a, b] for a, b in sub] == item["filter_ss"] and tot == item["totalout"]
            assert [S.has_one_good_bifurcation(b) for a, b in sub if not S.is_stem_loop(b, 3)] == item["bifurcation"]
            n += 1
    assert n > 20


def test_classifier_error_behaviour_matches_reference():
    with pytest.raises(KeyError):
        S.filter_ss("." * 60)            # reference: dict_pair[-1]
    with pytest.raises(IndexError):
        S.filter_ss("())")               # reference: pop from empty list


def test_duplex_oracle_matches_reference_fixture(stage3):
    """CPU: the duplex restatement (oracle) against the reference's own outputs."""
    import duplex_oracle as D
    st = stage3["structures"]
    for q in stage3["queries"]:
        got = D.maturestar(st[q["ss"]], q["mature"], q["fold_start"], q["region"][0], q["region"][1], q["strand"])
        want = q["result"] if isinstance(q["result"], str) else tuple(q["result"])
        assert got == want


def test_live_against_reference_when_present():
    """Build container only: compare with the AST-extracted reference functions directly."""
    if not os.path.exists("/root/reference/miR_PREFeR.py"):
        pytest.skip("reference not present")
    import sys
    sys.path.insert(0, GOLDEN)
    import ref_extract
    ns = ref_extract.load()
    for ss in ("((((...))))....((((....))))" * 3, "(((((...(((....)))..(((....)))...)))))" + "." * 30):
        assert ns["is_stem_loop"](ss, 3) == S.is_stem_loop(ss, 3)
        assert tuple(ns["filter_ss"](ss)) == tuple(S.filter_ss(ss))


@pytest.mark.gpu
def test_structures_from_result_equals_file_parser(mf, tmp_path):
    """Stage 1 without the text round trip == the reference-style parse of the RNALfold text."""
    text = open(os.path.join(GOLDEN, "synth8.in")).read()
    headers = [ln for ln in text.split("\n") if ln.startswith(">")]
    seqs = [ln for ln in text.split("\n") if ln and not ln.startswith(">")]
    out = tmp_path / "x_rnalfoldoutput_0"
    out.write_text(mf.fold_text(text, 300))
    with mf.fold(seqs, 300) as res:
        direct = list(S.structures_from_result(headers, res, 55, 3))
    assert direct == list(S.get_structures_next_extendregion(str(out), 55, 3))


@pytest.mark.gpu
def test_duplex_kernel_matches_reference_fixture(mf, stage3):
    st = stage3["structures"]
    qs = [(st[q["ss"]], q["mature"], q["fold_start"], q["region"][0], q["region"][1], q["strand"]) for q in stage3["queries"]]
    got = mf.duplex(qs)
    for g, q in zip(got, stage3["queries"]):
        want = q["result"] if isinstance(q["result"], str) else tuple(q["result"])
        assert g == want


@pytest.mark.gpu
def test_duplex_kernel_matches_oracle_on_fresh_folds(mf):
    """All (structure x mature) pairs of freshly folded loci: device verdicts == oracle restatement."""
    import numpy as np
    import duplex_oracle as D
    from corpus import synth_loci
    seqs = synth_loci(55, 24, "arabidopsis")
    rng = np.random.default_rng(8)
    qs = []
    with mf.fold(seqs, 300) as res:
        for r, s in enumerate(seqs):
            for ss, e, start in res.hits(r):
                if len(ss) < 55:
                    continue
                for norm, fs, sub, typ in S.classify(ss, e, start):
                    for _ in range(6):
                        strand = "+-"[int(rng.integers(2))]
                        rs = int(rng.integers(1, 10000)); re_ = rs + len(s)
                        mlen = int(rng.integers(18, 24)); l0 = int(rng.integers(0, max(1, len(sub) - mlen)))
                        m0 = l0 + rs + fs - 1 if strand == "+" else re_ - (l0 + mlen) - fs + 1
                        qs.append((sub, (m0, m0 + mlen), fs, rs, re_, strand))
    assert len(qs) > 500
    got = mf.duplex(qs)
    for g, q in zip(got, qs):
        assert g == D.maturestar(*q)
