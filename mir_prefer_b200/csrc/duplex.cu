// duplex.cu -- stage 3: miRNA/miRNA* duplex checks on the device (K5).
//
// Replaces get_maturestar_info() and helpers (/root/reference/miR_PREFeR.py:1727-1999: coordinate
// maps :1727/:1771, stat_duplex :1815, pass_stat_duplex :1848, get_maturestar_info :1876).  One warp
// per (structure, mature) query; the pair table and the duplex re-pairing are stack walks, done by
// lane 0 in shared-memory scratch (queries are tiny, throughput comes from thousands of resident
// warps).  Every (structure x mature) pair is evaluated because the host-side energy filter of
// check_loci (MP:2246-2302) depends on expression data.  Verdict codes index mirfold_duplex_fail_name().
#include "../../include/mirfold.h"
#include "mirfold_internal.cuh"

#define DPX_EXC 100   // the reference would raise a Python exception (unbalanced input)

__global__ void __launch_bounds__(128) k_duplex(const char *__restrict__ arena, const mirfold_duplex_query *__restrict__ qs,
                                                unsigned long long nq, mirfold_duplex_verdict *__restrict__ out, int maxlen)
{
    extern __shared__ short scratch[];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned long long q = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + wib;
    if (q >= nq || lane != 0) return;
    short *partner = scratch + (size_t)wib * 4 * maxlen;   // [maxlen]
    short *stk = partner + maxlen;                         // [maxlen]
    short *cpair = stk + maxlen;                           // [2*maxlen] partner (close index) of opens in the concatenation
    const mirfold_duplex_query Q = qs[q];
    const char *ss = arena + Q.ss_off;
    const int n = Q.ss_len;
    mirfold_duplex_verdict V;
    V.code = 0; V.star_start = V.star_end = V.fold_start = V.fold_end = 0;
    V.star_ss_begin = V.star_ss_end = V.mature_ss_begin = V.mature_ss_end = 0;
    V.prime5 = 0; V.total_dots = V.total_bps = 0;
#define FAIL(c) do { V.code = (c); out[q] = V; return; } while (0)
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
    const int m0 = Q.mature_start, m1 = Q.mature_end;
    int g0, g1, l0, l1;
    if (Q.strand == '+') { g0 = Q.region_start + Q.fold_start - 1; g1 = g0 + n; l0 = m0 - g0; l1 = m1 - g0; }
    else { g1 = Q.region_end - Q.fold_start + 1; g0 = g1 - n; l0 = g1 - m1; l1 = g1 - m0; }
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
    if (Q.strand == '+') { V.star_start = g0 + star_start; V.star_end = g0 + star_end; }
    else { V.star_start = g1 - star_end; V.star_end = g1 - star_start; }
    out[q] = V;
#undef FAIL
}

cudaError_t launch_duplex(const char *arena, const mirfold_duplex_query *qs, unsigned long long nq,
                          mirfold_duplex_verdict *out, int maxlen, cudaStream_t st)
{
    if (nq == 0) return cudaSuccess;
    const size_t per_warp = (size_t)4 * maxlen * sizeof(short);
    int wpb = (int)(40960 / (per_warp ? per_warp : 1));
    wpb = wpb < 1 ? 1 : (wpb > 4 ? 4 : wpb);
    const size_t smem = per_warp * wpb;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_duplex, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const unsigned long long blocks = (nq + wpb - 1) / wpb;
    k_duplex<<<(unsigned)blocks, wpb * 32, smem, st>>>(arena, qs, nq, out, maxlen);
    return cudaGetLastError();
}
