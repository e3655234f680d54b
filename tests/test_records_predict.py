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
