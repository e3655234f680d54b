"""Stage (1) of the hot path on the host: from fold results to the candidate-structure tuples the
predict stage consumes -- without the RNALfold text round trip.

Mirrors (same names, arguments, return values and error behaviour):
  get_structures_next_extendregion(rnalfoldoutname, minlen, minloop=3)   miR_PREFeR.py:1541-1599
  is_stem_loop(ss, minloopsize)                                          miR_PREFeR.py:1602-1608
  has_one_good_bifurcation(ss)                                           miR_PREFeR.py:1611-1659
  filter_ss(ss)                                                          miR_PREFeR.py:1685-1724
plus structures_from_result(), which yields the identical tuples straight from a FoldResult.
Written from the behaviour of those functions (tests/golden/stage1.json is produced by the
reference's own code); no arithmetic here beyond one float division per structure.
"""
import re

_ENERGY = re.compile(r"\(\s*(-?[0-9]+[\.]?[0-9]*)\s*\)")


def _pair_table(ss):
    """partner index of every bracket; raises IndexError on an unmatched ')' like list.pop()."""
    partner, open_stack = {}, []
    for k, ch in enumerate(ss):
        if ch == "(":
            open_stack.append(k)
        elif ch == ")":
            o = open_stack.pop()
            partner[o] = k
            partner[k] = o
    return partner


def is_stem_loop(ss, minloopsize):
    """One hairpin loop of at least minloopsize: nothing opens after the first ')'."""
    return ss.find(")") - ss.rfind("(") - 1 >= minloopsize


def filter_ss(ss):
    """Cut a multi-stem structure at its outermost stems.  Every outermost stem gives the piece
    [end of the previous outermost stem + 1, start of the next one) -- i.e. with both flanking gaps;
    pieces longer than 55 are returned as (offset, substructure).  Also returns the number of
    outermost stems.  (KeyError when ss has no pair, like the reference.)"""
    partner = _pair_table(ss)
    stems = []
    pos = ss.find("(")
    first = True
    while pos != -1:
        close = partner[pos]            # KeyError for pos == -1 on the first round, as in the reference
        stems.append((pos, close))
        pos = ss.find("(", close)
        first = False
    if first:
        partner[-1]                     # reference: dict_pair[ss.find('(')] with no '(' -> KeyError
    pieces = []
    for k, (o, c) in enumerate(stems):
        begin = 0 if k == 0 else stems[k - 1][1] + 1
        end = len(ss) if k + 1 == len(stems) else stems[k + 1][0]
        if end - begin > 55:
            pieces.append((begin, ss[begin:end]))
    return (pieces, len(stems))


def has_one_good_bifurcation(ss):
    """True for a structure of the form ( () () ): exactly one place where a stem opens right after
    another one closed *inside* an enclosing stem, with the two inner stems reasonably centred
    (partner positions: span < 0.5, first > 0.25, second < 0.75 of the length)."""
    depth_stack = []
    partner = {}
    prev_paren = "("      # kind of the last bracket seen
    prev_pos = 0
    n_bif = 0
    left_close = right_open = 0
    for k, ch in enumerate(ss):
        if ch == "(":
            if prev_paren == ")":
                if not depth_stack:
                    if k != 0:
                        return False          # ()() at the top level
                else:
                    if n_bif >= 1:
                        return False
                    n_bif = 1
                    left_close, right_open = prev_pos, k
            depth_stack.append(k)
            prev_paren, prev_pos = "(", k
        elif ch == ")":
            o = depth_stack.pop()
            partner[k] = o
            partner[o] = k
            prev_paren, prev_pos = ")", k
    n = len(ss)
    if float(partner[right_open] - partner[left_close]) / n < 0.5:
        if float(partner[left_close]) / n > 0.25:
            if float(partner[right_open]) / n < 0.75:
                return True
    return False


def classify(ss, energy_dcal, start, minloop=3):
    """The (norm_energy, fold_start, ss, sstype) tuples one printed hairpin contributes
    (miR_PREFeR.py:1570-1589).  energy is the printed %6.2f value, so energy_dcal/100."""
    norm_energy = float("%.2f" % (energy_dcal / 100.)) / len(ss)
    if is_stem_loop(ss, minloop):
        return [(norm_energy, start, ss, 0)]
    out = []
    pieces, _ = filter_ss(ss)
    for off, sub in pieces:
        if is_stem_loop(sub, minloop):
            out.append((norm_energy, start + off, sub, 0))
        elif has_one_good_bifurcation(sub):
            out.append((norm_energy, start + off, sub, 1))
    return out


def structures_from_result(headers, result, minlen, minloop=3):
    """Generator equivalent to get_structures_next_extendregion() fed from a FoldResult:
    one (which, peak, [(norm_energy, fold_start, ss, sstype)]) per record, input order.
    `headers` are the FASTA header lines of the records ('>SEQID:S-E STRAND PEAK TAG ...')."""
    for r, header in enumerate(headers):
        sp = header.strip().split()
        which, peak = sp[3], sp[2]
        structures = []
        for ss, e, start in result.hits(r):
            if len(ss) < minlen:
                continue
            structures.extend(classify(ss, e, start, minloop))
        yield (which, peak, structures)


def structures_from_result_native(headers, result, minlen, minloop=3):
    """structures_from_result() with the classification done by libmirfold's mirfold_classify() on all host
    cores instead of per-hit Python string handling (8 s -> well under 1 s for 10 k loci); same tuples."""
    per_rec = result.classify(minlen, minloop)
    for header, structures in zip(headers, per_rec):
        sp = header.strip().split()
        yield (sp[3], sp[2], structures)


def get_structures_next_extendregion(rnalfoldoutname, minlen, minloop=3):
    """Drop-in for the reference parser over an RNALfold-format text file (as written by
    MirFold.fold_fasta_files)."""
    structures = []
    which = 0
    peak = ""
    first = True
    with open(rnalfoldoutname) as f:
        for line in f:
            sp = line.strip().split()
            if line.startswith(">"):
                if not first:
                    yield (which, peak, structures)
                structures = []
                which, peak = sp[3], sp[2]
                first = False
            elif len(sp) >= 3 and len(sp[0]) >= minlen:
                e = float(_ENERGY.search(line).group(1))
                ss, start = sp[0], int(sp[-1])
                norm_energy = e / len(ss)
                if is_stem_loop(ss, minloop):
                    structures.append((norm_energy, start, ss, 0))
                else:
                    pieces, _ = filter_ss(ss)
                    for off, sub in pieces:
                        if is_stem_loop(sub, minloop):
                            structures.append((norm_energy, start + off, sub, 0))
                        elif has_one_good_bifurcation(sub):
                            structures.append((norm_energy, start + off, sub, 1))
        yield (which, peak, structures)
