"""CPU: host-side logic (RNALfold I/O contract, packing, sharding)."""
import numpy as np

from mir_prefer_b200.fold import MirFold, convert_sequence, format_record, parse_rnalfold_input
from mir_prefer_b200.shard import dp_cells, lpt_shards
from corpus import cells


def test_parse_rnalfold_input_contract():
    text = ">h1 a b\nACGT extra\n\n*c\n  GGG\n@\nIGNORED\n"
    assert parse_rnalfold_input(text) == [("echo", ">h1 a b"), ("seq", "ACGT"), ("echo", ""), ("echo", "*c"), ("seq", "GGG")]


def test_format_record_matches_printf():
    s = format_record("ggggaaaacccc", [("((((....))))", -540, 1)], -540)
    assert s == "((((....)))) ( -5.40)    1\nGGGGAAAACCCC\n ( -5.40)\n"
    s = format_record("ACGT", [("(((...)))", -2010, 1234)], -12345)
    assert s == "(((...))) (-20.10) 1234\nACGU\n (-123.45)\n"
    assert format_record("AAAAA", [(".", 0, 1)], 0) == ". (  0.00)    1\nAAAAA\n (  0.00)\n"


def test_convert_sequence():
    assert convert_sequence("acgtNnkx") == "ACGUNNKX"


def test_pack():
    buf, off = MirFold.pack(["ACG", "", b"TT"])
    assert buf.tobytes() == b"ACGTT" and off.tolist() == [0, 3, 3, 5]


def test_dp_cells_matches_definition():
    for n in (0, 3, 4, 5, 9, 50, 299, 300, 301, 325, 600, 2000):
        for L in (20, 150, 300, 500):
            assert int(dp_cells(n, L)) == cells(n, L), (n, L)


def test_lpt_shards_cover_and_balance():
    rng = np.random.default_rng(0)
    lens = rng.integers(5, 3000, size=500)
    shards = lpt_shards(lens, 300, 8)
    allidx = np.concatenate(shards)
    assert sorted(allidx.tolist()) == list(range(500))
    loads = [int(dp_cells(lens[s], 300).sum()) for s in shards]
    assert max(loads) <= 1.02 * (sum(loads) / 8) + int(dp_cells(lens.max(), 300))
