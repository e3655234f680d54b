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


def _result_from_rnalfold_text(path):
    """A mirfold_result built on the host from an RNALfold output file (no GPU): what mirfold_fold() would have
    returned for that input -- lets the native classifier be checked against the reference fixture on CPU."""
    import ctypes as C
    import re
    import numpy as np
    from mir_prefer_b200 import _lib
    energy = re.compile(r"\(\s*(-?[0-9]+[\.]?[0-9]*)\s*\)")
    recs, cur = [], None
    for line in open(path):
        sp = line.strip().split()
        if line.startswith(">"):
            cur = []
            recs.append(cur)
        elif len(sp) >= 3 and cur is not None:
            cur.append((sp[0], int(round(float(energy.search(line).group(1)) * 100)), int(sp[-1])))
    nseq = len(recs)
    nh = sum(len(r) for r in recs)
    hits = (_lib.Hit * max(nh, 1))()
    hb = np.zeros(nseq + 1, np.uint64)
    hc = np.zeros(nseq + 1, np.uint32)
    arena = bytearray()
    k = 0
    for r, rec in enumerate(recs):
        hb[r], hc[r] = k, len(rec)
        for ss, e, start in rec:
            hits[k].start, hits[k].len, hits[k].mfe_dcal, hits[k].ss_off = start, len(ss), e, len(arena)
            arena += ss.encode() + b"\0"
            k += 1
    arena_arr = np.frombuffer(bytes(arena), np.uint8).copy()
    totals = np.zeros(max(nseq, 1), np.int32)
    res = _lib.Result()
    res.nseq, res.nhits = nseq, nh
    res.hit_begin = hb.ctypes.data_as(C.POINTER(C.c_uint64))
    res.hit_count = hc.ctypes.data_as(C.POINTER(C.c_uint32))
    res.hits = C.cast(hits, C.POINTER(_lib.Hit))
    res.ss_arena = arena_arr.ctypes.data_as(C.c_void_p)
    res.ss_bytes = len(arena)
    res.total_mfe_dcal = totals.ctypes.data_as(C.POINTER(C.c_int32))
    return res, (hits, hb, hc, arena_arr, totals), nseq, arena_arr


def test_native_classifier_matches_reference_tuples(stage1):
    """mirfold_classify() (C++, host) on the golden RNALfold output vs the reference parser's own tuples."""
    import ctypes as C
    from mir_prefer_b200 import _lib
    from mir_prefer_b200.fold import classify_result
    lib = _lib.load()
    res, keep, nseq, arena = _result_from_rnalfold_text(os.path.join(GOLDEN, "synth8.L300.out"))
    got = classify_result(lib, C.pointer(res), nseq, arena, 55, 3)
    want = stage1["synth8.L300"]
    assert len(got) == len(want) and sum(len(g) for g in got) > 50
    for g, w in zip(got, want):
        assert [[float.hex(e), s, ss, t] for e, s, ss, t in g] == w["structures"]
    # the classifier unit vectors (every golden structure): same decisions as the Python rules
    for item in stage1["classifier"]:
        ss = item["ss"]
        assert S.classify(ss, -1234, 7) == _native_classify_one(lib, ss, -1234, 7)
    with pytest.raises(Exception):
        _native_classify_one(lib, "." * 60, -100, 1)       # the reference raises KeyError on a pairless hit


def test_native_formatter_reproduces_rnalfold_text_on_cpu():
    """mirfold_format_records() fed with the hits / totals parsed back from a golden RNALfold output must print
    that output again, byte for byte (no GPU involved: the result struct is built on the host)."""
    import ctypes as C
    import re
    import numpy as np
    from mir_prefer_b200 import _lib
    lib = _lib.load()
    path = os.path.join(GOLDEN, "synth8.L300.out")
    res, keep, nseq, arena = _result_from_rnalfold_text(path)
    # the converted sequence and the total line of every record
    energy = re.compile(r"\(\s*(-?[0-9]+[\.]?[0-9]*)\s*\)")
    blocks, seqs, totals, cur = [], [], [], None
    lines = open(path).read().split("\n")
    for k, line in enumerate(lines):
        if line.startswith(">"):
            cur = []
            blocks.append(cur)
        elif cur is not None and line != "":
            cur.append(line)
            if re.fullmatch(r" \(\s*-?[0-9.]+\)", line):
                totals.append(int(round(float(energy.search(line).group(1)) * 100)))
                seqs.append(cur[-2])
    assert len(seqs) == nseq == len(totals)
    tot = np.array(totals, np.int32)
    res.total_mfe_dcal = tot.ctypes.data_as(C.POINTER(C.c_int32))
    raw = "".join(s.replace("U", "T").lower() if k % 2 else s for k, s in enumerate(seqs)).encode()   # conversion is the formatter's job
    off = np.zeros(nseq + 1, np.uint64)
    off[1:] = np.cumsum([len(s) for s in seqs])
    buf = np.frombuffer(raw, np.uint8).copy()
    text, rec_off = C.c_void_p(), C.POINTER(C.c_uint64)()
    rc = lib.mirfold_format_records(C.pointer(res), buf.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.POINTER(C.c_uint64)), nseq,
                                    C.byref(text), C.byref(rec_off))
    assert rc == 0
    try:
        offs = np.ctypeslib.as_array(rec_off, shape=(nseq + 1,)).copy()
        data = C.string_at(text, int(offs[-1])).decode()
    finally:
        lib.mirfold_free_text(text, rec_off)
    for r in range(nseq):
        assert data[int(offs[r]):int(offs[r + 1])] == "\n".join(blocks[r]) + "\n", r


def _native_classify_one(lib, ss, e, start):
    import ctypes as C
    import numpy as np
    from mir_prefer_b200 import _lib
    from mir_prefer_b200.fold import classify_result
    hits = (_lib.Hit * 1)()
    hits[0].start, hits[0].len, hits[0].mfe_dcal, hits[0].ss_off = start, len(ss), e, 0
    hb, hc = np.zeros(2, np.uint64), np.array([1, 0], np.uint32)
    arena = np.frombuffer(ss.encode() + b"\0", np.uint8).copy()
    tot = np.zeros(1, np.int32)
    res = _lib.Result()
    res.nseq, res.nhits = 1, 1
    res.hit_begin = hb.ctypes.data_as(C.POINTER(C.c_uint64))
    res.hit_count = hc.ctypes.data_as(C.POINTER(C.c_uint32))
    res.hits = C.cast(hits, C.POINTER(_lib.Hit))
    res.ss_arena = arena.ctypes.data_as(C.c_void_p)
    res.ss_bytes = len(arena)
    res.total_mfe_dcal = tot.ctypes.data_as(C.POINTER(C.c_int32))
    return classify_result(lib, C.pointer(res), 1, arena, 1, 3)[0]


def test_classifier_error_behaviour_matches_reference():
    with pytest.raises(KeyError):
        S.filter_ss("." * 60)            # reference: dict_pair[-1]
    with pytest.raises(IndexError):
        S.filter_ss("())")               # reference: pop from empty list


@pytest.fixture(scope="module")
def stage3b():
    return json.load(open(os.path.join(GOLDEN, "stage3b.json")))


def _oracle_verdict(D, q):
    try:
        return D.maturestar(*q)
    except (KeyError, IndexError):
        return "EXCEPTION"


def test_duplex_oracle_matches_rare_branch_fixture(stage3b):
    """CPU: every FAIL_* branch incl. MATCHED_BASES, TOO_MANY_BULGE_OR_LOOP, TOTAL_LOOP_SIZE and the inputs on which
    the reference raises (tests/golden/make_golden_stage3b.py, the reference's own get_maturestar_info)."""
    import duplex_oracle as D
    from collections import Counter
    st = stage3b["structures"]
    seen = Counter()
    for q in stage3b["queries"]:
        got = _oracle_verdict(D, (st[q["ss"]], q["mature"], q["fold_start"], q["region"][0], q["region"][1], q["strand"]))
        want = q["result"] if isinstance(q["result"], str) else tuple(q["result"])
        assert got == want
        seen[want if isinstance(want, str) else "PASS"] += 1
    for name in ("FAIL_STRUCTURE_MATCHED_BASES", "FAIL_STRUCTURE_TOO_MANY_BULGE_OR_LOOP", "FAIL_STRUCTURE_TOTAL_LOOP_SIZE_LARGER_THAN_5",
                 "FAIL_STRUCTURE_MAX_BULGE_LARGE_THAN_2", "FAIL_STRUCTURE_NUM_BULGE_MORE_THAN_2", "EXCEPTION"):
        assert seen[name] >= 20, (name, seen[name])


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


def _random_structure(rng):
    """A balanced dot-bracket string of 56..300 chars: random nesting, multiloops, several outermost stems."""
    def helix(depth):
        arm = int(rng.integers(2, 12))
        if depth <= 0 or rng.random() < 0.45:
            inner = "." * int(rng.integers(3, 12))
        else:
            inner = "".join("." * int(rng.integers(0, 6)) + helix(depth - 1) for _ in range(int(rng.integers(1, 4))))
            inner += "." * int(rng.integers(0, 6))
        lb, rb = int(rng.integers(0, 4)), int(rng.integers(0, 4))
        return "(" * arm + "." * lb + inner + "." * rb + ")" * arm
    s = "." * int(rng.integers(0, 8))
    for _ in range(int(rng.integers(1, 4))):
        s += helix(int(rng.integers(0, 4))) + "." * int(rng.integers(0, 8))
    return s


def test_native_and_python_classifier_on_random_structures():
    """3 000 random balanced structures: mirfold_classify() == the Python rules, and -- in the build container --
    == the reference's own is_stem_loop / filter_ss / has_one_good_bifurcation applied the way its parser does."""
    import numpy as np
    from mir_prefer_b200 import _lib
    lib = _lib.load()
    ns = None
    if os.path.exists("/root/reference/miR_PREFeR.py"):
        import sys
        sys.path.insert(0, GOLDEN)
        import ref_extract
        ns = ref_extract.load()
    rng = np.random.default_rng(77)
    n = kinds = 0
    while n < 3000:
        ss = _random_structure(rng)
        if not 56 <= len(ss) <= 300:
            continue
        n += 1
        e, start = -int(rng.integers(0, 9000)), int(rng.integers(1, 300))
        want = S.classify(ss, e, start)
        assert _native_classify_one(lib, ss, e, start) == want
        kinds += len(want)
        if ns is not None:
            ne = float("%.2f" % (e / 100.)) / len(ss)
            ref = []
            if ns["is_stem_loop"](ss, 3):
                ref.append((ne, start, ss, 0))
            else:
                for off, sub in ns["filter_ss"](ss)[0]:
                    if ns["is_stem_loop"](sub, 3):
                        ref.append((ne, start + off, sub, 0))
                    elif ns["has_one_good_bifurcation"](sub):
                        ref.append((ne, start + off, sub, 1))
            assert ref == want
    assert kinds > 300


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
        native = list(S.structures_from_result_native(headers, res, 55, 3))
    assert direct == list(S.get_structures_next_extendregion(str(out), 55, 3))
    assert native == direct


@pytest.mark.gpu
def test_native_classifier_at_scale(mf):
    """mirfold_classify() vs the Python rules on 600 folded loci (threaded path, > 256 records)."""
    from mir_prefer_b200.corpus import synth_loci
    seqs = synth_loci(77, 600, "parity")
    headers = [">c:%d-%d + 1-22 0 1,22,+" % (k, k + len(s)) for k, s in enumerate(seqs)]
    with mf.fold(seqs, 300) as res:
        native = list(S.structures_from_result_native(headers, res, 55, 3))
        direct = list(S.structures_from_result(headers, res, 55, 3))
    assert native == direct and sum(len(x[2]) for x in native) > 5000


@pytest.mark.gpu
def test_duplex_kernel_matches_reference_fixture(mf, stage3):
    st = stage3["structures"]
    qs = [(st[q["ss"]], q["mature"], q["fold_start"], q["region"][0], q["region"][1], q["strand"]) for q in stage3["queries"]]
    got = mf.duplex(qs)
    for g, q in zip(got, stage3["queries"]):
        want = q["result"] if isinstance(q["result"], str) else tuple(q["result"])
        assert g == want


@pytest.mark.gpu
def test_duplex_kernel_matches_rare_branch_fixture(mf, stage3b):
    """k_duplex on the rare verdicts and on unbalanced input: code 100 exactly where the reference raises."""
    from mir_prefer_b200.fold import DUPLEX_EXCEPTION
    st = stage3b["structures"]
    qs = [(st[q["ss"]], q["mature"], q["fold_start"], q["region"][0], q["region"][1], q["strand"]) for q in stage3b["queries"]]
    got = mf.duplex(qs)
    for g, q in zip(got, stage3b["queries"]):
        want = q["result"] if isinstance(q["result"], str) else tuple(q["result"])
        assert g == (DUPLEX_EXCEPTION if want == "EXCEPTION" else want)


@pytest.mark.gpu
def test_duplex_kernel_matches_oracle_on_fresh_folds(mf):
    """All (structure x mature) pairs of freshly folded loci: device verdicts == oracle restatement."""
    import numpy as np
    import duplex_oracle as D
    from mir_prefer_b200.corpus import synth_loci
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
