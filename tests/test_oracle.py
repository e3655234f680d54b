"""CPU: the oracle restatement against the golden vectors produced by the reference's RNALfold."""
import hashlib
import json
import os

import pytest

from conftest import GOLDEN, ROOT, golden_cases
from mir_prefer_b200.corpus import lcg_records, records_to_fasta


@pytest.mark.parametrize("name,L", golden_cases())
def test_oracle_matches_golden_text(oracle, name, L):
    text = open(os.path.join(GOLDEN, name + ".in")).read()
    want = open(os.path.join(GOLDEN, "%s.L%d.out" % (name, L))).read()
    assert oracle.fold_text(text, L) == want


@pytest.mark.parametrize("seed", [3, 1])
def test_oracle_matches_sha256_pins(oracle, seed):
    """SURVEY.md App. C pins (sha256 of RNALfold's stdout on LCG-generated records)."""
    pins = json.load(open(os.path.join(GOLDEN, "pins.json")))["pins"]
    p = next(x for x in pins if x["seed"] == seed)
    text = records_to_fasta(lcg_records(p["seed"], p["nrec"], p["lo"], p["span"]))
    assert hashlib.sha256(text.encode()).hexdigest() == p["sha256_stdin"]
    out = oracle.fold_text(text, p["L"])
    assert hashlib.sha256(out.encode()).hexdigest() == p["sha256_stdout"]


def test_oracle_matches_reference_binary_when_present(oracle):
    """Where oracle/_ref/RNALfold (the reference's own binary) is staged, compare live."""
    if not oracle.have_rlf():
        pytest.skip("oracle/_ref/RNALfold not staged")
    from mir_prefer_b200.corpus import synth_loci
    text = records_to_fasta([("s%d" % k, s) for k, s in enumerate(synth_loci(77, 6, (60, 340)))])
    assert oracle.fold_text(text, 300) == oracle.fold_text(text, 300, binary=oracle.RLF)


def test_oracle_struct_api_matches_text(oracle):
    seq = "GGGGGTTTTCCCCCAAAAGGGGGTTTTCCCCC"
    r = oracle.fold(seq, 30)
    assert r["hits"] == [(".(((((....)))))", -930, 18), (".(((((....))))).", -980, 9),
                         ("(((((....)))))....(((((....)))))", -2010, 1)]
    assert r["total"] == -2010


def test_gpu_mini_check_digests_are_the_oracles():
    """tools/gpu_mini_check.py compares GPU results with digests computed here on the CPU: they must be the oracle's."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("gpu_mini_check", os.path.join(ROOT, "tools", "gpu_mini_check.py"))
    mc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mc)
    import oracle as O
    from mir_prefer_b200.corpus import synth_loci
    O.build()
    for n, L in mc.CASES[:3]:
        o = O.fold(synth_loci(700 + n, 1, (n, n))[0], L)
        assert mc.digest(o["hits"], o["total"]) == mc.WANT[(n, L)], (n, L)
