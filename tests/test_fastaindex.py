"""in-memory `samtools faidx` replacement (host-only), pinned to the stdout of the reference's own
bundled samtools 0.1.18 (tests/golden/faidx.json, made by tests/golden/make_golden_faidx.py) and to synthetic FASTA files."""
import json
import os

import numpy as np

from conftest import GOLDEN

from mir_prefer_b200.fastaindex import FastaIndex, write_fai


def make_fasta(tmp_path, width=50, crlf=False):
    rng = np.random.default_rng(9)
    seqs = {"Chr1": "".join(rng.choice(list("ACGTacgtNRY"), size=777)),
            "Chr2 description text": "".join(rng.choice(list("ACGT"), size=50)),     # exactly one full line
            "scaffold:7": "".join(rng.choice(list("ACGT"), size=121)),
            "empty": ""}
    nl = "\r\n" if crlf else "\n"
    p = tmp_path / ("g_crlf.fa" if crlf else "g.fa")
    with open(p, "w", newline="") as f:
        for name, s in seqs.items():
            f.write(">" + name + nl)
            for k in range(0, len(s), width):
                f.write(s[k:k + width] + nl)
    return str(p), {k.split()[0]: v for k, v in seqs.items()}


def test_fetch_matches_slicing_and_clips(tmp_path):
    for crlf in (False, True):
        path, seqs = make_fasta(tmp_path, crlf=crlf)
        fx = FastaIndex(path)
        assert fx.names() == ["Chr1", "Chr2", "scaffold:7", "empty"]
        rng = np.random.default_rng(1)
        for _ in range(300):
            name = ["Chr1", "Chr2", "scaffold:7"][int(rng.integers(3))]
            a, b = sorted(int(x) for x in rng.integers(1, len(seqs[name]) + 40, size=2))
            assert fx.fetch(name, a, b) == seqs[name][a - 1:b]                 # 1-based inclusive, end clipped
            assert fx.fetch_region("%s:%d-%d" % (name, a, b)) == seqs[name][a - 1:b]
        assert fx.fetch("Chr1", 0, 3) == seqs["Chr1"][:3]                      # start < 1 is clamped
        assert fx.fetch("Chr1", 778, 800) == "" and fx.fetch("nope", 1, 5) == "" and fx.fetch("empty", 1, 5) == ""
        assert fx.fetch_region("Chr2") == seqs["Chr2"] and fx.fetch_region("Chr1:700") == seqs["Chr1"][699:]
        assert fx.fetch_region("scaffold:7") == seqs["scaffold:7"]             # contig name containing ':'
        assert fx.fetch_region("scaffold:7:10-20") == seqs["scaffold:7"][9:20]


def test_reference_call_pattern(tmp_path):
    """dump_piece: region text 'seqid:start-(end-1)' for the half-open extend region, stdout minus the
    header line joined (miR_PREFeR.py:1098-1105)."""
    path, seqs = make_fasta(tmp_path)
    fx = FastaIndex(path)
    ext = (100, 400)
    region = "Chr1:%d-%d" % (ext[0], ext[1] - 1)
    out = fx.faidx_stdout(region)
    assert out.split("\n")[0] == ">" + region and max(len(x) for x in out.split("\n")[1:]) == 60
    joined = "".join(out.split("\n")[1:])
    assert joined == seqs["Chr1"][99:399] == fx.extend_region_sequence("Chr1", ext)
    assert len(joined) == ext[1] - ext[0]


def test_write_fai(tmp_path):
    path, seqs = make_fasta(tmp_path)
    rows = [line.split("\t") for line in open(write_fai(path)).read().splitlines()]
    assert [(r[0], int(r[1])) for r in rows] == [(k, len(v)) for k, v in seqs.items()]
    raw = open(path, "rb").read()
    for name, length, off, lb, lw in ((r[0], int(r[1]), int(r[2]), int(r[3]), int(r[4])) for r in rows):
        if length:
            assert (lb, lw) == (50, 51)
            assert raw[off:off + min(lb, length)].decode() == seqs[name][:min(lb, length)]


def test_matches_reference_samtools_stdout(tmp_path):
    """Every query of the golden fixture: the text `samtools faidx fasta region` printed (header + 60-column lines),
    what dump_piece keeps of it ("".join(stdout.split("\\n")[1:]), miR_PREFeR.py:1105), and the .fai index."""
    d = json.load(open(os.path.join(GOLDEN, "faidx.json")))
    nq = 0
    for k, case in enumerate(d["cases"]):
        fa = tmp_path / ("g%d.fa" % k)
        fa.write_text(case["fasta"])
        fx = FastaIndex(str(fa))
        assert open(write_fai(str(fa))).read() == case["fai"]
        for q in case["queries"]:
            kept = "".join(q["stdout"].split("\n")[1:])
            assert fx.fetch_region(q["region"]) == kept, q["region"]
            if q["rc"] == 0:                       # an unknown bare contig name makes samtools 0.1.18 crash without output
                assert fx.faidx_stdout(q["region"]) == q["stdout"], q["region"]
            nq += 1
    assert nq >= 90
