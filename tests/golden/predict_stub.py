"""Deterministic stand-ins for the samtools-backed expression statistics of the predict stage
(check_expression_new / gen_mapinfo_each_sample are outside the hot path) and a canonical JSON form
of check_loci / filter_next_loci outputs.  Shared by make_golden_predict.py (which drives the
reference's own check_loci / filter_next_loci with them) and by the tests (which drive ours)."""
import zlib

SAMPLES = ["sampleA", "sampleB"]


class _Lcg:
    def __init__(self, seed):
        self.x = seed & 0x7FFFFFFF

    def pick(self, seq):
        self.x = (1103515245 * self.x + 12345) % (1 << 31)
        return seq[(self.x >> 16) % len(seq)]


def check_expression_stub(dict_mapinfo_region, samplenames, fold0, fold1, mature, mdepth, star, strand, allow_3nt_overhang):
    r = _Lcg(zlib.crc32(repr((fold0, fold1, tuple(mature), mdepth, tuple(star), strand, bool(allow_3nt_overhang))).encode()))
    info = {
        "mature_star_distance": r.pick([2, 4, 5, 12, 20, 33]),
        "total_depth_star": r.pick([0, 0, 0, 3, 25, 400]),
        "mature_star_ratio_total": r.pick([0.05, 0.19, 0.2, 0.6, 0.95]),
        "mature_iso_star_ratio_total": r.pick([0.3, 0.79, 0.8, 0.9, 1.0]),
        "total_depth_mature": r.pick([20, 100, 101, 999, 1000, 5000]),
        "mature_depth_each_sample": [r.pick([0, 2, 9, 50]) for _ in samplenames],
    }
    for s in samplenames:
        info[s] = {"ratio_bases_with_reads_start": r.pick([0.1, 0.4, 0.5, 0.51, 0.9])}
    if r.pick([0, 0, 1]):
        info["max_imperfect_star"] = r.pick([5, 50])
        info["imperfect_star_start"] = star[0] + r.pick([-1, 1, 2])
        info["imperfect_star_end"] = star[1] + r.pick([-1, 1, 2])
    return info


def mapinfo_stub(*args):
    return {"region": list(args[-3:])}


def canon(obj):
    """JSON-able canonical form: tuples -> lists, dicts -> sorted [repr(key), value] pairs."""
    if isinstance(obj, dict):
        return {"__dict__": sorted([[repr(k), canon(v)] for k, v in obj.items()], key=lambda kv: kv[0])}
    if isinstance(obj, (list, tuple)):
        return [canon(v) for v in obj]
    if isinstance(obj, float):
        return float.hex(obj)
    return obj
