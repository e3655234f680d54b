"""CPU: host-side logic (RNALfold I/O contract, packing, sharding)."""
import numpy as np

from mir_prefer_b200.fold import MirFold, convert_sequence, format_record, parse_rnalfold_input, plan_shards
from mir_prefer_b200.corpus import cells


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
    """The library's closed-form DP-cell count (shard weights, stats.cells) == SURVEY 8(d)'s row sum."""
    for n in (0, 3, 4, 5, 9, 50, 299, 300, 301, 325, 600, 2000):
        for L in (20, 150, 300, 500):
            _, c = plan_shards([n], L, 1)
            assert int(c[0]) == cells(n, L), (n, L)


def test_lpt_shards_cover_and_balance():
    """mirfold_plan_shards (the plan mirfold_fold itself uses): every record in exactly one shard, loads within
    one largest locus of the mean, deterministic, and one shard == everything."""
    rng = np.random.default_rng(0)
    lens = rng.integers(0, 3000, size=5000)
    shard_of, load = plan_shards(lens, 300, 8)
    assert shard_of.max() == 7 and len(shard_of) == 5000
    want = np.zeros(8, np.int64)
    for n, g in zip(lens.tolist(), shard_of.tolist()):
        want[g] += cells(n, 300)
    assert want.tolist() == load.tolist()
    assert load.max() <= load.mean() + cells(int(lens.max()), 300)
    assert load.max() <= 1.002 * load.mean()
    again, _ = plan_shards(lens, 300, 8)
    assert (again == shard_of).all()
    one, tot = plan_shards(lens, 300, 1)
    assert (one == 0).all() and int(tot[0]) == int(load.sum())
    # heavy tail (BASELINE configs[3]: 2-10 kb loci next to short ones)
    lens = np.concatenate([rng.integers(2000, 10000, size=300), rng.integers(5, 400, size=3000)])
    _, load = plan_shards(lens, 300, 8)
    assert load.max() <= 1.01 * load.mean()


def test_plan_shards_rejects_bad_arguments():
    import ctypes as C
    from mir_prefer_b200 import _lib
    lib = _lib.load()
    off = np.array([0, 10, 5], np.uint64)   # decreasing
    assert lib.mirfold_plan_shards(off.ctypes.data_as(C.POINTER(C.c_uint64)), 2, 300, 2, None, None) == -3
    off = np.array([0, 10, 15], np.uint64)
    assert lib.mirfold_plan_shards(off.ctypes.data_as(C.POINTER(C.c_uint64)), 2, 300, 0, None, None) == -3
    assert lib.mirfold_plan_shards(off.ctypes.data_as(C.POINTER(C.c_uint64)), 2, 300, 2, None, None) == 0
