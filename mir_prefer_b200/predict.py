"""Predict-stage consumer of the fold results.

Mirrors the reference's per-locus decision procedure and its positional pairing of fold records
with alignment records:
  check_loci(...)          miR_PREFeR.py:2206-2341
  filter_next_loci(...)    miR_PREFeR.py:2344-2432
with the one call that sits on the hot path -- get_maturestar_info() for every
(mature, structure) pair, miR_PREFeR.py:2258-2262 -- answered from ONE batched device launch
(`DuplexTable`, libmirfold's mirfold_duplex) instead of a Python call per pair.  The sequential
`lowest_energy` rule (a structure is only looked at if its normalised energy is not above the best
accepted so far, :2246-2252) depends on the expression checks and therefore stays on the host; the
device answers the superset of pairs and the host loop picks what the reference would have asked for.

Expression statistics (check_expression_new, gen_mapinfo_each_sample: samtools-backed, SURVEY 2
rows 6-9) are outside the hot path: they are injected as callables with the reference's signatures,
so the unmodified reference functions plug in.
"""

from .fold import DUPLEX_EXCEPTION

PASSED, FAILED = "PASSED", "FAILED"


def _size_ok(m0, m1, lo, hi):
    return lo <= m1 - m0 <= hi


class DuplexTable:
    """get_maturestar_info() answers for every (size-admissible mature, structure) pair of a batch
    of records, computed by one mirfold_duplex() call.  Call it like the reference function."""

    def __init__(self, mf, items, min_mature_len, max_mature_len):
        """items: iterable of (structures, matures, region) with structures =
        [(norm_energy, fold_start, ss, sstype)], matures = [(m0, m1, strand, depth)] and region =
        [seqid, (start, end), strand] as in the alignment dump (:1133, :1151)."""
        keys, queries = [], []
        seen = set()
        for structures, matures, region in items:
            rs, re_ = region[1][0], region[1][1]
            for m0, m1, strand, _depth in matures:
                if not _size_ok(m0, m1, min_mature_len, max_mature_len):
                    continue
                for _e, foldstart, ss, _t in structures:
                    k = (ss, m0, m1, foldstart, rs, re_, strand)
                    if k not in seen:
                        seen.add(k)
                        keys.append(k)
                        queries.append((ss, (m0, m1), foldstart, rs, re_, strand))
        self.n_queries = len(queries)
        self._table = dict(zip(keys, mf.duplex(queries))) if queries else {}

    @classmethod
    def from_candidates(cls, cand, records):
        """The same table from a fused device pass (MirFold.fold_candidates): `cand` already holds the verdict of every
        (candidate structure, size-admissible mature) pair of every record; `records` are the LocusRecords (or anything
        with .region) the candidates were computed for, in order."""
        self = cls.__new__(cls)
        self._table = {}
        for r, rec in enumerate(records):
            rs, re_ = rec.region[0], rec.region[1]
            structs = cand.structures(r)
            for k, (m0, m1, strand, _depth), verdict in cand.verdicts_of(r):
                _e, foldstart, ss, _t = structs[k]
                self._table[(ss, m0, m1, foldstart, rs, re_, strand)] = verdict
        self.n_queries = cand.nverdicts
        return self

    def __call__(self, ss, mature, foldstart, foldend, regionstart, regionend, strand):
        v = self._table[(ss, mature[0], mature[1], foldstart, regionstart, regionend, strand)]
        if v == DUPLEX_EXCEPTION:
            # the table answers the superset of pairs; only a pair the reference really asks for may raise like it
            raise KeyError("get_maturestar_info: unbalanced structure around the mature (the reference raises here)")
        return v


def check_loci(structures, matures, region, dict_mapinfo_region, which, samplenames, allow_3nt_overhang, allow_no_star,
               min_mature_len, max_mature_len, peak_depth, maturestar, check_expression):
    """miR_PREFeR.py:2206-2341.  Returns the list of miRNA records of the region, or -- if there is
    none -- the dict of reasons, exactly as the reference builds it (only the first failing check
    of a (mature, structure) pair is recorded).  `structures` is (which, [structure tuples]);
    `maturestar` / `check_expression` stand for get_maturestar_info / check_expression_new."""
    key = tuple(region)
    why = {key: {"PEAK_PASS_DEPTH": PASSED}}
    if not structures[1]:
        why[key]["HAS_STEMLOOP_STRUCTURE"] = FAILED
        return why
    why[key]["HAS_STEMLOOP_STRUCTURE"] = PASSED
    if not any(_size_ok(m[0], m[1], min_mature_len, max_mature_len) for m in matures):
        why[key]["HAS_MATURE_SIZE_IN_RANGE"] = FAILED     # also when there is no mature at all (0 == 0)
        return why
    why[key]["HAS_MATURE_SIZE_IN_RANGE"] = PASSED

    rs, re_ = region[1][0], region[1][1]
    found = []
    for m0, m1, strand, mdepth in sorted(matures, key=lambda m: m[3], reverse=True):   # deepest first, stable
        if not _size_ok(m0, m1, min_mature_len, max_mature_len):
            continue
        best_energy, best = 0, []
        for energy, foldstart, ss, _sstype in structures[1]:
            if energy > best_energy:
                continue
            note = why[key][(m0, m1, strand, ss)] = {}
            info = maturestar(ss, (m0, m1), foldstart, foldstart + len(ss), rs, re_, strand)
            if isinstance(info, str):
                note[info] = FAILED
                continue
            star0, star1, fold0, fold1 = info[0], info[1], info[2], info[3]
            expr = check_expression(dict_mapinfo_region, samplenames, fold0, fold1, (m0, m1), mdepth, (star0, star1),
                                    strand, allow_3nt_overhang)

            def record(has_star):
                return [region[0], fold0, fold1, m0, m1, star0, star1, ss, strand, has_star, expr]

            if expr["mature_star_distance"] <= 4:
                note["FAIL_EXPRESS_PATTERN_MATURE_STAR_TOO_CLOSE"] = FAILED
                note["ss_info"] = record(True)
            elif expr["total_depth_star"] > 0:
                if expr["mature_star_ratio_total"] < 0.2:
                    note["FAIL_EXPRESS_PATTERN_HAS_STAR_BUT_TOO_FEW_READS_MAPPED_TO_DUPLEX"] = FAILED
                    note["ss_info"] = record(True)
                else:
                    best = record(True)
                    if "max_imperfect_star" in expr:
                        best[5], best[6] = expr["imperfect_star_start"], expr["imperfect_star_end"]
                    best_energy = energy
            elif not allow_no_star:
                note["FAIL_EXPRESS_PATTERN_NO_STAR_EXPRESSION_DISALLOW_NO_STAR"] = FAILED
                note["ss_info"] = record(False)
                note["expression_info"] = expr
            elif any(expr[sample]["ratio_bases_with_reads_start"] > 0.5 for sample in samplenames):
                note["FAIL_EXPRESS_PATTERN_NO_STAR_EXPRESSION_TOO_MANY_START"] = FAILED
                note["ss_info"] = record(False)
                note["expression_info"] = expr
            else:
                everywhere = all(x > 0 for x in expr["mature_depth_each_sample"])
                if expr["mature_iso_star_ratio_total"] >= 0.8 and (everywhere or expr["total_depth_mature"] >= 1000):
                    best = record(False)
                    best_energy = energy
                else:
                    if expr["mature_iso_star_ratio_total"] < 0.8:
                        note["FAIL_EXPRESS_PATTERN_NO_STAR_MATURE_STAR_RATIO_TOO_SMALL"] = FAILED
                    if expr["total_depth_mature"] <= 100:
                        note["FAIL_EXPRESS_PATTERN_NO_STAR_MATURE_DEPTH_TOO_SMALL"] = FAILED
                    if not everywhere:
                        note["FAIL_EXPRESS_PATTERN_NO_STAR_MATURE_NOT_IN_ALL_SAMPLE"] = FAILED
                    note["ss_info"] = record(True)      # the reference flags has_star here (:2331)
                    note["expression_info"] = expr
        if best:
            found.append(best)
    return found if found else why


def filter_next_loci(aln_records, ss_records, mapinfo, samplenames, allow_3nt_overhang, allow_no_star, output_details,
                     min_mature_len, max_mature_len, peak_depth, maturestar, check_expression):
    """miR_PREFeR.py:2344-2432 as a generator over in-memory records.
    aln_records: iterable of (region, which, dict_aln, matures) -- the alignment dump, in order;
    ss_records : iterable of (which, peak, structures) -- one per fold record, same order
                 (RecordFold.structures(minlen) or get_structures_next_extendregion());
    mapinfo(seqid, start, end) stands for gen_mapinfo_each_sample on the combined BAM."""
    aln = iter(aln_records)
    ss_iter = iter(ss_records)

    def judge(ss_rec, matures, region, which):
        return check_loci((ss_rec[0], ss_rec[2]), matures, region, mapinfo(region[0], region[1][0], region[1][1]), which,
                          samplenames, allow_3nt_overhang, allow_no_star, min_mature_len, max_mature_len, peak_depth,
                          maturestar, check_expression)

    for region, which, _dict_aln, matures in aln:
        if which == "0":                       # a locus with a single extend region
            rec = next(ss_iter)
            res = judge(rec, matures, region, which)
            if isinstance(res, dict):
                if not output_details:
                    continue
                res["which"], res["peak"] = "0", rec[1]
            yield res
            continue
        rec_l, rec_r = next(ss_iter), next(ss_iter)      # left and right extend region of one locus
        region_r, which_r, _aln_r, matures_r = next(aln)
        peak = rec_l[1]
        res = judge(rec_l, matures, region, which)
        if isinstance(res, dict):
            res["which"], res["peak"] = "L", peak
        yield res                                        # the L verdict is yielded even without output_details (:2406-2410)
        if isinstance(res, dict):                        # the right side is only tried when the left one failed
            res = judge(rec_r, matures_r, region_r, which_r)
            if isinstance(res, dict):
                if not output_details:
                    continue
                res["which"], res["peak"] = "R", peak
            yield res


def duplex_items(aln_records, ss_records):
    """(structures, matures, region) triples for DuplexTable from the two record streams."""
    return [(ss[2], a[3], a[0]) for a, ss in zip(aln_records, ss_records)]
