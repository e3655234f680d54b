// duplex_dev.cuh -- get_maturestar_info() for one (structure, mature) pair as a device function, shared by
// k_duplex (queries from the host, mirfold_duplex) and the fused candidate kernels (candidates.cu).
//
// Follows /root/reference/miR_PREFeR.py:1876-1999 (coordinate maps :1727/:1771, stat_duplex :1815,
// pass_stat_duplex :1848).  The pair table and the duplex re-pairing are stack walks in `scratch`
// (4*maxlen shorts of shared memory per caller).  Verdict codes index mirfold_duplex_fail_name(); DPX_EXC
// marks input on which the reference raises.
#pragma once
#include "../../include/mirfold.h"

#define DPX_EXC 100   // the reference would raise a Python exception (unbalanced input)

__device__ __forceinline__ void dev_duplex_eval(const char *__restrict__ ss, int n, int fold_start, int m0_, int m1_,
                                                int region_start, int region_end, int strand, short *scratch, int maxlen,
                                                mirfold_duplex_verdict &V)
{
    short *partner = scratch;            // [maxlen]
    short *stk = partner + maxlen;       // [maxlen]
    short *cpair = stk + maxlen;         // [2*maxlen] partner (close index) of opens in the concatenation
    V.code = 0; V.star_start = V.star_end = V.fold_start = V.fold_end = 0;
    V.star_ss_begin = V.star_ss_end = V.mature_ss_begin = V.mature_ss_end = 0;
    V.prime5 = 0; V.total_dots = V.total_bps = 0;
#define FAIL(c) do { V.code = (c); return; } while (0)
    if (n > maxlen) FAIL(DPX_EXC);
    int sp = 0;
    for (int k = 0; k < n; k++) {
        partner[k] = -1;
        const char ch = ss[k];
        if (ch == '(') stk[sp++] = (short)k;
        else if (ch == ')') {
            if (sp == 0) FAIL(1);
            const int o = stk[--sp];
            partner[o] = (short)k;
            partner[k] = (short)o;
        }
    }
    const int m0 = m0_, m1 = m1_;
    int g0, g1, l0, l1;
    if (strand == '+') { g0 = region_start + fold_start - 1; g1 = g0 + n; l0 = m0 - g0; l1 = m1 - g0; }
    else { g1 = region_end - fold_start + 1; g0 = g1 - n; l0 = g1 - m1; l1 = g1 - m0; }
    V.fold_start = g0; V.fold_end = g1; V.mature_ss_begin = l0; V.mature_ss_end = l1;
    if (!(m0 >= g0 && m1 <= g1)) FAIL(2);
    int n_open = 0, n_close = 0, firstbp = -1, lastbp = -1;
    for (int k = l0; k < l1; k++) { n_open += ss[k] == '('; n_close += ss[k] == ')'; }
    if (n_open && n_close) FAIL(3);
    if (n_open + n_close < 14) FAIL(4);
    const char sym = n_open ? '(' : ')';
    V.prime5 = n_open ? 1 : 0;
    int mend = -1;
    for (int k = l0; k < l1; k++)
        if (ss[k] == sym) {
            if (firstbp < 0) firstbp = k;
            lastbp = k;
            if (k < l1 - 2) mend = k;
        }
    if (partner[lastbp] < 0 || partner[firstbp] < 0) FAIL(DPX_EXC);   // unmatched '(': dict_bp[...] raises KeyError (MP:1932-1933)
    const int star_start = partner[lastbp] - (l1 - 1 - lastbp) + 2;
    const int star_end = partner[firstbp] + (firstbp - l0) + 3;
    V.star_ss_begin = star_start; V.star_ss_end = star_end;
    if (l0 <= star_start) {
        if (star_start - l1 < 3) FAIL(5);
        if (star_end > n) FAIL(6);
    }
    if (star_start <= l0) {
        if (l0 - star_end < 3) FAIL(5);
        if (star_start < 0) FAIL(6);
    }
    if (mend < 0 || partner[mend] < 0) FAIL(DPX_EXC);                   // dict_bp[mend] raises (MP:1949)
    const int sstart = partner[mend], send = partner[firstbp];
    const int Lm = mend + 1 - l0, Lsd = max(0, send + 1 - sstart);
    int dots_m = 0, dots_s = 0;
    for (int k = l0; k <= mend; k++) dots_m += ss[k] == '.';
    for (int k = sstart; k <= send; k++) dots_s += ss[k] == '.';
    V.total_dots = dots_m + dots_s;
    V.total_bps = Lm - dots_m;
    if (V.total_bps < 14) FAIL(4);
    {
        int so = 0, sc = 0;
        for (int k = max(star_start, 0); k < min(star_end, n); k++) { so += ss[k] == '('; sc += ss[k] == ')'; }
        if (so && sc) FAIL(7);
    }
    // stat_duplex on cat = mature_duplex + star_duplex
    const int Lc = Lm + Lsd;
    int po = -1, pc = -1;
    for (int k = 0; k < Lc; k++) {
        const char ch = k < Lm ? ss[l0 + k] : ss[sstart + k - Lm];
        if (ch == '(' && po < 0) po = k;
        if (ch == ')' && pc < 0) pc = k;
    }
    const char oc = (po > pc) ? ')' : '(', cc = (po > pc) ? '(' : ')';
    sp = 0;
    for (int k = 0; k < Lc; k++) {
        cpair[k] = -1;
        const char ch = k < Lm ? ss[l0 + k] : ss[sstart + k - Lm];
        if (ch == oc) stk[sp++] = (short)k;
        else if (ch == cc) {
            if (sp == 0) FAIL(DPX_EXC);
            cpair[stk[--sp]] = (short)k;
        }
    }
    int n_loops = 0, n_bulges = 0, tot_loop = 0, max_bulge = 0, prev = -1;
    for (int k = 0; k < Lc; k++) {
        if (cpair[k] < 0) continue;
        if (prev >= 0) {
            const int ga = k - prev - 1, gb = cpair[prev] - cpair[k] - 1;
            if (!(ga == 0 && gb == 0)) {
                if (ga == gb) { n_loops++; tot_loop += ga; }
                else { n_bulges++; max_bulge = max(max_bulge, max(ga, gb)); }
            }
        }
        prev = k;
    }
    if (n_loops + n_bulges > 5) FAIL(8);
    if (max_bulge > 2) FAIL(9);
    if (tot_loop > 5) FAIL(10);
    if (n_bulges > 2) FAIL(11);
    if (strand == '+') { V.star_start = g0 + star_start; V.star_end = g0 + star_end; }
    else { V.star_start = g1 - star_end; V.star_end = g1 - star_start; }
#undef FAIL
}
