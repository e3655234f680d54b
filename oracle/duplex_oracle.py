"""CPU restatement of the reference's miRNA/miRNA* duplex check (stage 3 of the hot path).

TEST INFRASTRUCTURE ONLY (tests/, smoke, bench cpu legs).  Restates get_maturestar_info() and its
helpers -- /root/reference/miR_PREFeR.py:1727-1999 (pos_genome_2_local :1727, pos_local_2_genome
:1771, stat_duplex :1815, pass_stat_duplex :1848, get_maturestar_info :1876) -- as integer logic on
a pair table, the same formulation the CUDA kernel (mir_prefer_b200/csrc/duplex.cu) uses.
Parity pin: oracle/check_duplex_oracle.py compares it with the reference's own functions
(AST-extracted, build container only) on randomized queries; tests/golden/stage3.json holds
reference outputs for the GPU box.

Returns either the reference's 9-tuple or its FAIL_* string.
"""

FAIL_NAMES = [
    "PASS",
    "FAIL_STRUCTURE_MATCHED_BASES",
    "FAIL_STRUCTURE_MATURE_NOT_IN_FOLD_REGION",
    "FAIL_STRUCTURE_MATURE_NOT_IN_ONE_ARM",
    "FAIL_STRUCTURE_MATURE_MATCH_SMALL_THAN_14",
    "FAIL_STRUCTURE_MATURE_STAR_OVERLAP",
    "FAIL_STRUCTURE_STAR_OUT_OF_FOLD_REGION",
    "FAIL_STRUCTURE_STAR_NOT_IN_ONE_ARM",
    "FAIL_STRUCTURE_TOO_MANY_BULGE_OR_LOOP",
    "FAIL_STRUCTURE_MAX_BULGE_LARGE_THAN_2",
    "FAIL_STRUCTURE_TOTAL_LOOP_SIZE_LARGER_THAN_5",
    "FAIL_STRUCTURE_NUM_BULGE_MORE_THAN_2",
]


def maturestar(ss, mature, foldstart, regionstart, regionend, strand):
    n = len(ss)
    m0, m1 = mature
    partner = [-1] * n
    st = []
    for k, ch in enumerate(ss):
        if ch == "(":
            st.append(k)
        elif ch == ")":
            if not st:
                return FAIL_NAMES[1]
            o = st.pop()
            partner[o] = k
            partner[k] = o
    if strand == "+":
        g0 = regionstart + foldstart - 1
        g1 = g0 + n
        l0, l1 = m0 - g0, m1 - g0
    else:
        g1 = regionend - foldstart + 1
        g0 = g1 - n
        l0, l1 = g1 - m1, g1 - m0
    if not (m0 >= g0 and m1 <= g1):
        return FAIL_NAMES[2]
    lo, hi = l0, max(l0, l1)
    n_open = sum(1 for k in range(lo, hi) if ss[k] == "(")
    n_close = sum(1 for k in range(lo, hi) if ss[k] == ")")
    if n_open and n_close:
        return FAIL_NAMES[3]
    if n_open + n_close < 14:
        return FAIL_NAMES[4]
    sym = "(" if n_open else ")"
    prime5 = bool(n_open)
    arm = [k for k in range(lo, hi) if ss[k] == sym]
    firstbp, lastbp = arm[0], arm[-1]
    if partner[lastbp] < 0 or partner[firstbp] < 0:
        raise KeyError("unmatched bracket in the mature (dict_bp[...] of MP:1932-1933)")
    star_start = partner[lastbp] - (l1 - 1 - lastbp) + 2
    star_end = partner[firstbp] + (firstbp - l0) + 3
    if l0 <= star_start:
        if star_start - l1 < 3:
            return FAIL_NAMES[5]
        if star_end > n:
            return FAIL_NAMES[6]
    if star_start <= l0:
        if l0 - star_end < 3:
            return FAIL_NAMES[5]
        if star_start < 0:
            return FAIL_NAMES[6]
    inner = [k for k in arm if k < l1 - 2]
    if not inner or partner[inner[-1]] < 0:
        raise KeyError("dict_bp[mend] of MP:1949")
    mend = inner[-1]
    sstart, send = partner[mend], partner[firstbp]
    md = ss[l0:mend + 1]
    sd = ss[sstart:send + 1]
    total_dots = md.count(".") + sd.count(".")
    total_bps = len(md) - md.count(".")
    if total_bps < 14:
        return FAIL_NAMES[4]
    star_ss = ss[star_start:star_end]
    if "(" in star_ss and ")" in star_ss:
        return FAIL_NAMES[7]
    # stat_duplex on the concatenation; "open" is whichever bracket kind appears first
    cat = md + sd
    po, pc = cat.find("("), cat.find(")")
    oc, cc = ("(", ")") if not (po > pc) else (")", "(")
    st, pairs = [], {}
    for k, ch in enumerate(cat):
        if ch == oc:
            st.append(k)
        elif ch == cc:
            pairs[st.pop()] = k      # IndexError where the reference's stat_duplex pops an empty list
    keys = sorted(pairs)
    n_loops = n_bulges = tot_loop = max_bulge = 0
    for a, b in zip(keys, keys[1:]):
        ga, gb = b - a - 1, pairs[a] - pairs[b] - 1
        if ga == 0 and gb == 0:
            continue
        if ga == gb:
            n_loops += 1
            tot_loop += ga
        else:
            n_bulges += 1
            max_bulge = max(max_bulge, ga, gb)
    if n_loops + n_bulges > 5:
        return FAIL_NAMES[8]
    if max_bulge > 2:
        return FAIL_NAMES[9]
    if tot_loop > 5:
        return FAIL_NAMES[10]
    if n_bulges > 2:
        return FAIL_NAMES[11]
    if strand == "+":
        gs0, gs1 = g0 + star_start, g0 + star_end
    else:
        gs0, gs1 = g1 - star_end, g1 - star_start
    return (gs0, gs1, g0, g1, star_ss, prime5, ss[l0:l1], total_dots, total_bps)
