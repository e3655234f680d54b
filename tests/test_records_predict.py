"""locus records straight into the fold (records.py) and the predict-stage
consumer fed by batched device duplex verdicts (predict.py), against fixtures produced by the
reference's own functions (tests/golden/make_golden_predict.py, make_golden_stage13.py)."""
import json
import os
import sys

import pytest

from conftest import GOLDEN

sys.path.insert(0, GOLDEN)
import predict_stub as PS  # noqa: E402
from make_golden_predict import make_loci  # noqa: E402  (seeded generator only; no reference access)

from mir_prefer_b200 import predict as P  # noqa: E402
from mir_prefer_b200 import records as R  # noqa: E402
from mir_prefer_b200 import structures as S  # noqa: E402


@pytest.fixture(scope="module")
def predict_cases():
    return json.load(open(os.path.join(GOLDEN, "predict.json")))["cases"]


def run_ours(aln, ssrecs, allow_no_star, details, maturestar):
    return [PS.canon(x) for x in P.filter_next_loci(
        aln, ssrecs, PS.mapinfo_stub, PS.SAMPLES, True, allow_no_star, details, 18, 24, 20, maturestar, PS.check_expression_stub)]


# ------------------------------------------------------------------------------------ f1 (CPU)
def test_header_line_matches_reference_format():
    """SURVEY Appendix C (dump_piece, miR_PREFeR.py:1124-1143): header text and parser fields."""
    rec = R.LocusRecord("Chr1", (100, 400), "+", (150, 171), "0", [(150, 171, "+")], [(150, 171, "+", 55)], "ACGT")
    assert R.header_line(rec) == ">Chr1:100-400 + 150-171 0 150,171,+ M:150-171/+/55"
    rec2 = R.LocusRecord("scaffold_7:x", (5, 320), "-", (40, 62), "L", [(40, 62, "-"), (70, 91, "-")], [], "ACGT")
    h = R.header_line(rec2)
    assert h == ">scaffold_7:x:5-320 - 40-62 L 40,62,-;70,91,-"
    sp = h.split()
    assert (sp[2], sp[3]) == ("40-62", "L")                      # what get_structures_next_extendregion reads
    assert R.parse_header(h) == ("scaffold_7:x", (5, 320), "-", (40, 62), "L", [(40, 62, "-"), (70, 91, "-")], [])
    assert R.parse_header(R.header_line(rec))[6] == [(150, 171, "+", 55)]
    no_peaks = R.LocusRecord("c", (1, 9), "+", (2, 3), "R", [], [], "")
    assert R.header_line(no_peaks) == ">c:1-9 + 2-3 R "          # trailing blank like the reference's string concat


def test_reverse_complement_quirk():
    """miR_PREFeR.py:232-239: maketrans('ATGCU','UACGA') -- upper case only, output in RNA letters."""
    assert R.get_reverse_complement("AACGT") == "ACGUU"
    assert R.get_reverse_complement("ACGU") == "ACGU"
    assert R.get_reverse_complement("aacgtN") == "Ntgcaa"       # lower case / IUPAC: reversed, not complemented
    assert R.get_reverse_complement("AaCcGgTt") == "tAgCcGaU"
    assert R.get_reverse_complement("") == ""
    ref = os.path.join("/root/reference", "miR_PREFeR.py")
    if os.path.exists(ref):                                      # build container: the reference's own table
        import ast
        ns = {"string": type("S", (), {"maketrans": staticmethod(str.maketrans)})}
        for node in ast.parse(open(ref).read()).body:
            if isinstance(node, ast.FunctionDef) and node.name in ("get_complement", "get_reverse_complement"):
                exec(compile(ast.Module(body=[node], type_ignores=[]), ref, "exec"), ns)
        import numpy as np
        rng = np.random.default_rng(5)
        for _ in range(200):
            s = "".join(rng.choice(list("ACGTUacgtuNRYKMnX-"), size=int(rng.integers(0, 80))))
            assert R.get_reverse_complement(s) == ns["get_reverse_complement"](s)


def test_fasta_round_trip(tmp_path):
    recs = [R.LocusRecord("Chr2", (10, 40), "-", (12, 33), "0", [(12, 33, "-")], [(12, 33, "-", 7)], "AACCGGTTAAGGnnACGT"),
            R.LocusRecord("Chr2", (50, 60), "+", (52, 58), "L", [(52, 58, "+")], [], ""),
            R.LocusRecord("Chr2", (70, 90), "+", (72, 78), "R", [(72, 78, "+")], [], "GGGGAAAACCCC")]
    p = tmp_path / "shard.fa"
    R.write_fasta(recs, str(p))
    text = p.read_text()
    assert text.split("\n")[1] == R.get_reverse_complement("AACCGGTTAAGGnnACGT")
    assert text.split("\n")[3] == ""                              # empty sequence still owns a line
    back = R.records_from_fasta(text)
    assert [(b.seqid, b.region, b.strand, b.locus, b.tag, b.peaks, b.matures) for b in back] == \
           [(r.seqid, r.region, r.strand, r.locus, r.tag, r.peaks, r.matures) for r in recs]
    q = tmp_path / "again.fa"
    R.write_fasta(back, str(q))
    assert q.read_text() == text


# ------------------------------------------------------------------------------------ f2 (CPU)
def test_predict_consumer_matches_reference_with_duplex_oracle(predict_cases):
    """check_loci / filter_next_loci restatement vs the reference's own functions (fixture), with the
    CPU duplex oracle answering get_maturestar_info."""
    import duplex_oracle as DO

    def maturestar(ss, mature, foldstart, foldend, rs, re_, strand):
        assert foldend == foldstart + len(ss)
        return DO.maturestar(ss, mature, foldstart, rs, re_, strand)

    loci = {}
    nlist = 0
    for case in predict_cases:
        key = (case["seed"], case["nloci"])
        if key not in loci:
            loci[key] = make_loci(*key)
        aln, ssrecs = loci[key]
        got = run_ours(aln, ssrecs, case["allow_no_star"], case["output_details"], maturestar)
        assert got == case["expected"]
        nlist += sum(isinstance(x, list) for x in got)
    assert nlist >= 40                                            # the fixture does contain predicted miRNAs


def test_check_loci_early_exits():
    region = ["Chr1", (100, 300), "+"]
    boom = lambda *a: (_ for _ in ()).throw(AssertionError("must not be called"))   # noqa: E731
    r = P.check_loci(("0", []), [(110, 131, "+", 9)], region, {}, "0", PS.SAMPLES, True, True, 18, 24, 20, boom, boom)
    assert r == {tuple(region): {"PEAK_PASS_DEPTH": "PASSED", "HAS_STEMLOOP_STRUCTURE": "FAILED"}}
    st = [(-0.5, 1, "(((...)))", 0)]
    for matures in ([], [(110, 127, "+", 9)], [(110, 135, "+", 9)]):
        r = P.check_loci(("0", st), matures, region, {}, "0", PS.SAMPLES, True, True, 18, 24, 20, boom, boom)
        assert r[tuple(region)]["HAS_MATURE_SIZE_IN_RANGE"] == "FAILED"


# ------------------------------------------------------------------------------------ composed end to end
@pytest.fixture(scope="module")
def e2e_cases():
    d = json.load(open(os.path.join(GOLDEN, "e2e_mirna.json")))
    for c in d["cases"]:
        c["records"] = [R.LocusRecord(x[0], tuple(x[1]), x[2], tuple(x[3]), x[4], [tuple(p) for p in x[5]], [tuple(m) for m in x[6]], x[7])
                        for x in c["records"]]
    return d


def _aln_of(records):
    return [[[r.seqid, r.region, r.strand], r.tag, {}, list(r.matures)] for r in records]


def test_end_to_end_mirna_list_oracle_path(e2e_cases, oracle, tmp_path):
    """CPU: records -> FASTA -> oracle RNALfold text -> parser restatement -> consumer restatement with the duplex oracle
    == the list the reference's own binary + parser + filter_next_loci produced (tests/golden/make_golden_e2e.py)."""
    import duplex_oracle as DO

    def maturestar(ss, mature, foldstart, foldend, rs, re_, strand):
        return DO.maturestar(ss, mature, foldstart, rs, re_, strand)

    case = e2e_cases["cases"][0]
    fa = tmp_path / "shard.fa"
    R.write_fasta(case["records"], str(fa))
    out = tmp_path / "shard_rnalfoldoutput_0"
    out.write_text(oracle.fold_text(fa.read_text(), e2e_cases["span"]))
    ssrecs = list(S.get_structures_next_extendregion(str(out), 55, 3))
    for key in ("11", "10", "01", "00"):
        got = run_ours(_aln_of(case["records"]), ssrecs, key[0] == "1", key[1] == "1", maturestar)
        assert got == case["expected"][key], key
    assert sum(isinstance(x, list) for x in case["expected"]["11"]) >= 5


@pytest.mark.gpu
def test_end_to_end_mirna_list_fused_device_path(mf, e2e_cases):
    """GPU: the same records through ONE fused device pass (fold + stage 1 + stage 3, mirfold_fold_candidates) and the host
    consumer: identical final miRNA list and identical reasons, for every option combination."""
    for case in e2e_cases["cases"]:
        ssrecs, table, cand = R.candidates_of_records(mf, case["records"], e2e_cases["span"])
        assert cand.nverdicts > 100 and cand.nstructs > 100
        for key in ("11", "10", "01", "00"):
            got = run_ours(_aln_of(case["records"]), ssrecs, key[0] == "1", key[1] == "1", table)
            assert got == case["expected"][key], (case["seed"], key)


@pytest.mark.gpu
def test_fused_candidates_equal_separate_stages(mf):
    """mirfold_fold_candidates == mirfold_fold + mirfold_classify + mirfold_duplex on the same records (several chunks,
    two pipelines on one GPU), and the download is a small fraction of the hit text."""
    import numpy as np
    import mir_prefer_b200 as mp
    from mir_prefer_b200.corpus import synth_loci
    seqs = synth_loci(61, 400, "arabidopsis") + ["", "ACGT"]
    rng = np.random.default_rng(9)
    regions, matures, moff = [], [], [0]
    for s in seqs:
        rs = int(rng.integers(1, 50000))
        regions.append([rs, rs + len(s)])
        for _ in range(int(rng.integers(0, 4))):
            mlen = int(rng.choice([16, 20, 21, 22, 25]))
            strand = "+-"[int(rng.integers(2))]
            m0 = rs + int(rng.integers(0, max(1, len(s) - mlen)))
            matures.append((m0, m0 + mlen, strand, int(rng.integers(1, 99))))
        moff.append(len(matures))
    buf, off = mf.pack(seqs)
    with mf.fold_packed(buf, off, 300) as res:
        per_rec = res.classify(55)
        d2h_full = res.stats["d2h_bytes"]
    queries, keys = [], []
    for r, structs in enumerate(per_rec):
        for k, (_e, fs, ss, _t) in enumerate(structs):
            for m in matures[moff[r]:moff[r + 1]]:
                if 18 <= m[1] - m[0] <= 24:
                    queries.append((ss, (m[0], m[1]), fs, regions[r][0], regions[r][1], m[2]))
                    keys.append((r, k, m))
    want = dict(zip(keys, mf.duplex(queries)))
    os.environ["MIRFOLD_MEM_BUDGET_MB"] = "64"
    try:
        with mp.MirFold(devices=[0, 0]) as m2:
            cand = m2.fold_candidates(buf, off, 300, regions, matures, np.array(moff, np.uint64))
    finally:
        del os.environ["MIRFOLD_MEM_BUDGET_MB"]
    assert cand.stats["n_chunks"] >= 4
    got = {}
    for r in range(len(seqs)):
        assert cand.structures(r) == per_rec[r], r
        for k, m, v in cand.verdicts_of(r):
            got[(r, k, m)] = v
    assert got == want and len(want) > 300
    assert cand.stats["d2h_bytes"] < 0.25 * d2h_full


# ------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_predict_consumer_with_device_duplex_table(mf, predict_cases):
    """Same fixture, get_maturestar_info answered by ONE mirfold_duplex launch per batch."""
    loci = {}
    for case in predict_cases:
        key = (case["seed"], case["nloci"])
        if key not in loci:
            aln, ssrecs = make_loci(*key)
            table = P.DuplexTable(mf, P.duplex_items(aln, ssrecs), 18, 24)
            assert table.n_queries > 100
            loci[key] = (aln, ssrecs, table)
        aln, ssrecs, table = loci[key]
        assert run_ours(aln, ssrecs, case["allow_no_star"], case["output_details"], table) == case["expected"]


@pytest.mark.gpu
def test_fold_records_equals_text_pipeline(mf, oracle, tmp_path):
    """Records -> device -> structure tuples == records -> FASTA -> RNALfold text -> reference-format parser;
    the RNALfold text itself is byte-identical to the oracle's for the shard the reference would write."""
    from mir_prefer_b200.corpus import synth_loci
    seqs = synth_loci(31, 10, (120, 330)) + ["", "ACGTNNacgtACGTTGCA"]
    recs = []
    for k, s in enumerate(seqs):
        strand = "-" if k % 3 == 1 else "+"
        tag = ["0", "L", "R"][k % 3]
        recs.append(R.LocusRecord("Chr%d" % (k % 2 + 1), (1000 * k + 1, 1000 * k + 1 + len(s)), strand, (1000 * k + 40, 1000 * k + 61),
                                  tag, [(1000 * k + 40, 1000 * k + 61, strand)], [(1000 * k + 40, 1000 * k + 61, strand, 30 + k)], s))
    recs.append(recs[3])                                          # duplicated record (both-strand L/R quirk): folded once, kept twice
    fa = tmp_path / "shard.fa"
    R.write_fasta(recs, str(fa))
    want_text = oracle.fold_text(fa.read_text(), 300)          # the oracle CLI restates RNALfold's main() as well
    with R.fold_records(mf, recs, 300) as rf:
        assert rf.rnalfold_text() == want_text
        assert mf.fold_text(fa.read_text(), 300) == want_text
        out = tmp_path / "shard.rnalfold"
        rf.write_rnalfold_text(str(out))
        direct = list(rf.structures(55))
        assert rf.result.stats["nt"] == sum(len(s) for s in set(R.fold_sequence(r) for r in recs))
    parsed = list(S.get_structures_next_extendregion(str(out), 55))
    assert len(direct) == len(recs) and direct == parsed
