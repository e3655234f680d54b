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


def test_fill_units_tile_every_row_of_every_locus():
    """mirfold_plan_fill_units (the shape the band fill itself uses): buckets by length, 608-nt tiles for long loci at
    narrow spans, 864-nt tiles (and single 864 units for 609..864 nt) at wide spans, the generic kernel only where no
    tile would own 64 rows.  For every tiled locus: every row i has an owner tile that contains all of (i, i+4..i+dmax),
    owner tiles change exactly at multiples of tile_step, and the last tile ends at base n."""
    from mir_prefer_b200.fold import plan_fill_units
    BIG = 300   # spans from here on use the 864 bucket (host.cu: MF_BIG_TILE_MIN_SPAN)
    for L in (5, 30, 150, 299, 300, 400, 500, 544, 545, 700, 800, 801, 1000):
        for n in (5, 6, 160, 161, 352, 353, 608, 609, 700, 864, 865, 866, 927, 928, 929, 1228, 1229, 1500, 2500, 5000, 10000):
            p = plan_fill_units(n, L)
            dmax = min(L, n - 1)
            assert p["dmax"] == dmax
            if n <= 608:
                assert p["kernel"] == p["stride"] == (160 if n <= 160 else 352 if n <= 352 else 608)
                assert p["n_units"] == (1 if dmax >= 4 else 0) and p["tile_len"] == n
                continue
            big = dmax >= BIG and (n <= 864 or 864 - dmax >= 64)
            if big and n <= 864:
                assert (p["kernel"], p["stride"], p["n_units"], p["tile_len"]) == (864, 864, 1, n), (n, L, p)
                continue
            TL = 864 if big else 608
            if TL - dmax < 64:
                assert p["kernel"] == 0 and p["n_units"] == 1 and p["stride"] >= n and p["stride"] % 32 == 0, (n, L, p)
                assert p["stride"] not in (160, 352, 608, 864)
                continue
            S = TL - dmax
            assert (p["kernel"], p["stride"], p["tile_len"], p["tile_step"]) == (TL, TL, TL, S), (n, L, p)
            nt = p["n_units"]
            assert nt == (n - TL + S - 1) // S + 1 and nt >= 2
            assert p["band_cells"] == nt * TL * (dmax - 3)
            starts = [min(t * S, n - TL) for t in range(nt)]
            assert starts[-1] == n - TL
            for i in list(range(1, min(n - 4, 3 * S) + 1)) + list(range(max(1, n - 2 * TL), n - 3)):
                t = min((i - 1) // S, nt - 1)
                a = starts[t]
                assert a + 1 <= i, (n, L, i)
                assert min(i + dmax, n) <= a + TL, (n, L, i)
    assert plan_fill_units(4, 300)["n_units"] == 0
