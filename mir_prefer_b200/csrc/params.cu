// params.cu -- energy parameter set and the per-lane term schedules of the fill kernels.
//
// Replaces scale_parameters()/paramT of RNALfold 1.8.5 (RLF @0x415440; tables of SURVEY.md Appendix D,
// extracted by oracle/extract_tables.py into turner99_v185_tables.inc): T = 37 C copies of the *37 tables,
// dangles clamped <= 0, MLintern[t] = 40 (+50 for AU/GU closures), MLbase = 0, MLclosing = 340.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "mirfold_internal.cuh"
#include "turner99_v185_tables.inc"

namespace {

// ------------------------------------------------------------------ narrow-kernel schedule
// Word-terms of the 16-bit pair ring (see k_fill_s16).  For a cell on diagonal d, pair slot m
// (0..15) holds the inner diagonals of loop sizes (s_lo, s_hi) = (2m, 2m-1) for even d and
// (2m+1, 2m) for odd d.  A word-term is (ring, m, x-offset): generic terms read Cm at row offset u
// for both sizes; bulges read c+AU at offset 0 (5' side unpaired = 0) as pairs and at offset s
// (3' side) as single halves.  The bank class of a word-term is (xoff - 17 m) mod 32 and lane = class,
// so every unrolled iteration is one conflict-free LDS.
void build_s16_schedule(DevParams &P)
{
    const int ninio = T99_F_ninio37[2], maxninio = T99_MAX_NINIO;
    struct Term { int m, xo, ring, clo, chi; bool vlo, vhi; };
    auto generic_ok = [](int s, int u) { const int v = s - u; return s >= 0 && s <= 30 && u >= 1 && v >= 1 && !(u <= 2 && v <= 2); };
    auto gconst = [&](int s, int u) { return T99_internal_loop37[s] + std::min(maxninio, std::abs(2 * u - s) * ninio); };
    for (int par = 0; par < 2; par++) {
        std::vector<Term> G[32], B[32];
        for (int m = 0; m < 16; m++) {
            const int slo = par ? 2 * m + 1 : 2 * m, shi = par ? 2 * m : 2 * m - 1;
            for (int u = 1; u <= 30; u++) {
                const bool a = generic_ok(slo, u), b = generic_ok(shi, u);
                if (a || b) G[((u - MF16_SKEW * m) % 32 + 32) % 32].push_back({m, u, 0, a ? gconst(slo, u) : 0, b ? gconst(shi, u) : 0, a, b});
            }
            const bool a = slo >= 2 && slo <= 30, b = shi >= 2 && shi <= 30;
            if (a || b)
                B[((0 - MF16_SKEW * m) % 32 + 32) % 32].push_back({m, 0, 1, a ? T99_bulge37[slo] - MF16_DBIAS : 0, b ? T99_bulge37[shi] - MF16_DBIAS : 0, a, b});
            if (a) B[((slo - MF16_SKEW * m) % 32 + 32) % 32].push_back({m, slo, 1, T99_bulge37[slo] - MF16_DBIAS, 0, true, false});
            if (b) B[((shi - MF16_SKEW * m) % 32 + 32) % 32].push_back({m, shi, 1, 0, T99_bulge37[shi] - MF16_DBIAS, false, true});
        }
        for (int lane = 0; lane < 32; lane++) {
            std::stable_sort(G[lane].begin(), G[lane].end(), [](const Term &x, const Term &y) { return (x.vlo && x.vhi) < (y.vlo && y.vhi); });
            int nmask = 0;
            for (const Term &t : G[lane]) nmask += !(t.vlo && t.vhi);
            if ((int)G[lane].size() > MF16_NQG || (int)B[lane].size() > MF16_NQB || nmask > MF16_NMG) {
                fprintf(stderr, "mirfold: 16-bit schedule overflow (par %d lane %d: %zu generic, %zu bulge, %d masked)\n", par, lane,
                        G[lane].size(), B[lane].size(), nmask);
                abort();
            }
            auto put = [&](int q, const Term *t) {
                if (t) {
                    P.s16_td[par][q][lane] = (unsigned)t->m | ((unsigned)t->xo << 4) | ((unsigned)t->ring << 10);
                    P.s16_cst[par][q][lane] = ((unsigned)t->clo & 0xffffu) | ((unsigned)t->chi << 16);
                } else {   // no term: read the all-INF row at a lane-private bank
                    P.s16_td[par][q][lane] = (unsigned)lane << 4 | 1u << 11;
                    P.s16_cst[par][q][lane] = 0;
                }
                const unsigned mk = t ? ((t->vlo ? 0xffffu : 0u) | (t->vhi ? 0xffff0000u : 0u)) : 0xffffffffu;
                if (q < MF16_NMG) P.s16_mk[par][q][lane] = mk;
                else if (q >= MF16_NQG) P.s16_mk[par][MF16_NMG + q - MF16_NQG][lane] = mk;
            };
            for (int q = 0; q < MF16_NQG; q++) put(q, q < (int)G[lane].size() ? &G[lane][q] : nullptr);
            for (int q = 0; q < MF16_NQB; q++) put(MF16_NQG + q, q < (int)B[lane].size() ? &B[lane][q] : nullptr);
        }
    }
}

// ------------------------------------------------------------------ parameter set (a10)
}  // namespace

void build_params(DevParams &P)
{
    memset(&P, 0, sizeof P);
    for (int s = 0; s <= MF_MAX_SPAN + 1; s++) {
        if (s <= 30) P.hairpinE[s] = T99_hairpin37[s];
        else P.hairpinE[s] = T99_hairpin37[30] + (int)(T99_lxc37 * log(s / 30.));
    }
    for (int k = 0; k < 31; k++) { P.bulge[k] = T99_bulge37[k]; P.internal_loop[k] = T99_internal_loop37[k]; }
    for (int k = 0; k < 64; k++) { P.stack[k] = T99_stack37[k]; P.pair[k] = (unsigned char)T99_BP_pair[k]; }
    for (int k = 0; k < 200; k++) { P.mismatchI[k] = T99_mismatchI37[k]; P.mismatchH[k] = T99_mismatchH37[k]; }
    for (int k = 0; k < 40; k++) {  // dangles are clamped to <= 0 by scale_parameters
        P.dangle5[k] = std::min(0, T99_dangle5_37[k]);
        P.dangle3[k] = std::min(0, T99_dangle3_37[k]);
    }
    for (int t = 0; t < 8; t++) {
        P.MLintern[t] = T99_ML_intern37 + (t > 2 ? T99_TerminalAU : 0);
        P.rtype[t] = (unsigned char)T99_rtype[t];
    }
    memcpy(P.int11, T99_int11_37, sizeof P.int11);
    memcpy(P.int21, T99_int21_37, sizeof P.int21);
    memcpy(P.int22, T99_int22_37, sizeof P.int22);
    P.MLclosing = T99_ML_closing37;
    P.TerminalAU = T99_TerminalAU;
    // tetraloop bonus by packed 6-mer (first listed entry wins, like strstr)
    for (int k = T99_N_TETRALOOPS - 1; k >= 0; k--) {
        int code = 0;
        bool ok = true;
        for (int c = 0; c < 6; c++) {
            const char ch = T99_Tetraloops[7 * k + c];
            int b = ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : ch == 'U' ? 3 : -1;
            if (b < 0) ok = false;
            code |= (b & 3) << (2 * c);
        }
        if (ok) P.tetra[code] = (short)T99_TETRA_ENERGY37[k];
    }
    // generic interior-loop constants: iteration it pairs loop sizes s=it and s=30-it in one warp
    const int ninio = T99_F_ninio37[2], maxninio = T99_MAX_NINIO;
    for (int it = 0; it < 16; it++)
        for (int lane = 0; lane < 32; lane++) {
            int s, u;
            if (it < 15) { if (lane <= it) { s = it; u = lane; } else { s = 30 - it; u = lane - it - 1; } }
            else { s = 15; u = lane; }
            const int v = s - u;
            int val = MF_INF;
            if (u >= 0 && v >= 0 && u <= s) {
                const bool special = (u == 0 || v == 0 || (u <= 2 && v <= 2));
                if (!special) val = T99_internal_loop37[s] + std::min(maxninio, std::abs(u - v) * ninio);
            }
            P.ilc[it][lane] = val;
        }
    // skewed-ring schedule: lane = bank class (u - A*(u+v)) mod 32, <= MF_GEN_ITERS terms per lane
    {
        int fill[32] = {0};
        for (int k = 0; k < MF_GEN_ITERS; k++)
            for (int l = 0; l < 32; l++) { P.gen_c[k][l] = MF_INF; P.gen_us[k][l] = 0; }
        for (int u = 0; u <= 30; u++)
            for (int v = 0; u + v <= 30; v++) {
                if (u == 0 || v == 0 || (u <= 2 && v <= 2)) continue;
                const int cls = (((u - MF_SKEW_A * (u + v)) % 32) + 32) % 32;
                const int k = fill[cls]++;
                if (k >= MF_GEN_ITERS) { fprintf(stderr, "mirfold: skew schedule overflow\n"); abort(); }
                P.gen_c[k][cls] = T99_internal_loop37[u + v] + std::min(maxninio, std::abs(u - v) * ninio);
                P.gen_us[k][cls] = u | ((u + v) << 8) | (1 << 16);
            }
    }
    int m = 0;
    for (int u = 0; u <= 30; u++)
        for (int v = 0; v <= 30 - u; v++) { P.uv[m][0] = (unsigned char)u; P.uv[m][1] = (unsigned char)v; m++; }
    build_s16_schedule(P);
}
