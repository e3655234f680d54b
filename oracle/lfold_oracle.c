/*
 * lfold_oracle.c -- CPU restatement of the reference fold stage's arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (mir_prefer_b200/) may import, link
 * or execute this file; it exists so that tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg have a debuggable checker (it can expose the c / fML / f3 matrices, which
 * the reference binary cannot).
 *
 * What it restates: `RNALfold -L <span>` of ViennaRNA 1.8.5 with default flags (-d1, 37 C,
 * tetraloop bonus on), which is what miR-PREFeR's fold stage shells out to
 * (/root/reference/miR_PREFeR.py:3053, :3064).  ViennaRNA's source is NOT part of the reference
 * checkout -- only the binary dependency/Linux/x64/RNALfold (ViennaRNA 1.8.5, unstripped,
 * DWARF) -- so this file follows the behavioural spec in SURVEY.md Appendix A, whose items cite
 * the binary's DWARF line numbers ("RLF Lfold.c:NNN").  Sections below name the spec item.
 *
 * Parity pin: oracle/check_oracle.py runs this against the reference binary itself (sha256
 * golden pins of SURVEY.md App. C + randomized corpora); tests/golden holds RLF outputs.
 *
 * Design differs from the reference on purpose (this is a restatement, not a copy): full
 * band matrices indexed [i][d] instead of L+5 rolling rows, explicit accessor functions that
 * return INF outside the computed band, a sector stack on the heap.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../mir_prefer_b200/csrc/turner99_v185_tables.inc"

#define INF 1000000
#define TURN 3
#define MAXLOOP 30

/* ------------------------------------------------------------------ parameters (App. A.3, a10) */
typedef struct {
    int stack[8][8], hairpin[31], bulge[31], internal_loop[31];
    int mismatchI[8][5][5], mismatchH[8][5][5], dangle5[8][5], dangle3[8][5];
    int int11[8][8][5][5], int21[8][8][5][5][5], int22[8][8][5][5][5][5];
    int MLintern[8], MLbase, MLclosing, TerminalAU, ninio2, max_ninio;
    int tetra_energy[64];
    double lxc;
    int pair[8][8], rtype[8];
} params_t;

static params_t P;
static int P_ready = 0;

/* scale_parameters at T=37: tables are taken as-is; dangles clamped to <= 0; MLintern[t] gets
 * TerminalAU for t>2 (RLF params.c:58-61, :88-95 per SURVEY A.3). */
static void params_init(void)
{
    if (P_ready) return;
    memcpy(P.stack, T99_stack37, sizeof P.stack);
    memcpy(P.hairpin, T99_hairpin37, sizeof P.hairpin);
    memcpy(P.bulge, T99_bulge37, sizeof P.bulge);
    memcpy(P.internal_loop, T99_internal_loop37, sizeof P.internal_loop);
    memcpy(P.mismatchI, T99_mismatchI37, sizeof P.mismatchI);
    memcpy(P.mismatchH, T99_mismatchH37, sizeof P.mismatchH);
    memcpy(P.int11, T99_int11_37, sizeof P.int11);
    memcpy(P.int21, T99_int21_37, sizeof P.int21);
    memcpy(P.int22, T99_int22_37, sizeof P.int22);
    memcpy(P.pair, T99_BP_pair, sizeof P.pair);
    memcpy(P.rtype, T99_rtype, sizeof P.rtype);
    for (int t = 0; t < 8; t++)
        for (int b = 0; b < 5; b++) {
            int d5 = T99_dangle5_37[t * 5 + b], d3 = T99_dangle3_37[t * 5 + b];
            P.dangle5[t][b] = d5 > 0 ? 0 : d5;
            P.dangle3[t][b] = d3 > 0 ? 0 : d3;
        }
    P.TerminalAU = T99_TerminalAU;
    for (int t = 0; t < 8; t++) P.MLintern[t] = T99_ML_intern37 + (t > 2 ? P.TerminalAU : 0);
    P.MLbase = 0;
    P.MLclosing = T99_ML_closing37;
    P.ninio2 = T99_F_ninio37[2];
    P.max_ninio = T99_MAX_NINIO;
    P.lxc = T99_lxc37;
    for (int k = 0; k < T99_N_TETRALOOPS; k++) P.tetra_energy[k] = T99_TETRA_ENERGY37[k];
    P_ready = 1;
}

/* ------------------------------------------------------------------ fold state */
typedef struct {
    int n, Ls;      /* length, L* = min(L, n)                               (A: Lfold.c:149) */
    int W;          /* row width of band matrices = Ls + 6                                   */
    const char *seq; /* converted sequence, 0-based                                          */
    short *S, *S1;  /* 1-based codes                                                         */
    int *c, *m, *dml; /* [ (n+2) rows ][ W ]                                                 */
    int *f3;        /* n+3                                                                   */
} fold_t;

static inline int enc(char ch)
{   /* A.1: index in "_ACGUTXKI", indices >4 decremented; anything else 0 */
    switch (ch) {
    case 'A': return 1; case 'C': return 2; case 'G': return 3; case 'U': return 4;
    case 'T': return 4; case 'X': return 5; case 'K': return 6; case 'I': return 7;
    default: return 0;
    }
}
static const int ALIAS[8] = {0, 1, 2, 3, 4, 3, 2, 0};

static inline int ptype(const fold_t *f, int i, int j)
{   /* A.1: typed only for 4 <= j-i <= L*-1, j <= n */
    int d = j - i;
    if (i < 1 || j > f->n || d < TURN + 1 || d >= f->Ls) return 0;
    return P.pair[f->S[i]][f->S[j]];
}
static inline int band_get(const fold_t *f, const int *a, int i, int j)
{
    int d = j - i;
    if (i < 1 || i > f->n || j > f->n || d < 0 || d >= f->W) return INF;
    return a[(size_t)i * f->W + d];
}
#define C_(i, j) band_get(f, f->c, (i), (j))
#define M_(i, j) band_get(f, f->m, (i), (j))
#define D_(i, j) band_get(f, f->dml, (i), (j))
static inline int F_(const fold_t *f, int i) { return (i >= 1 && i <= f->n + 2) ? f->f3[i] : 0; }
static inline int AU(int t) { return t > 2 ? P.TerminalAU : 0; }

/* A.2 hairpin */
static int hairpin_E(const fold_t *f, int i, int j)
{
    int s = j - i - 1, t = ptype(f, i, j);
    int e = s <= 30 ? P.hairpin[s] : P.hairpin[30] + (int)(P.lxc * log(s / 30.));
    if (s == 4) {
        char six[7];
        memcpy(six, f->seq + (i - 1), 6);
        six[6] = 0;
        const char *hit = strstr(T99_Tetraloops, six);
        if (hit) e += P.tetra_energy[(hit - T99_Tetraloops) / 7];
    }
    if (s == 3) e += AU(t);            /* no mismatch, Triloops empty */
    else e += P.mismatchH[t][f->S1[i + 1]][f->S1[j - 1]];
    return e;
}

/* A.2 two-loop closed by (i,j) [type t] and (p,q) [t2 = rtype of its type] */
static int loop_E(const fold_t *f, int i, int j, int p, int q, int t, int t2)
{
    int n1 = p - i - 1, n2 = j - q - 1;
    int nl = n1 > n2 ? n1 : n2, ns = n1 > n2 ? n2 : n1;
    const short *S1 = f->S1;
    if (nl == 0) return P.stack[t][t2];
    if (ns == 0) {
        int e = P.bulge[nl];
        return nl == 1 ? e + P.stack[t][t2] : e + AU(t) + AU(t2);
    }
    if (ns == 1 && nl == 1) return P.int11[t][t2][S1[i + 1]][S1[j - 1]];
    if (ns == 1 && nl == 2)
        return n1 == 1 ? P.int21[t][t2][S1[i + 1]][S1[q + 1]][S1[j - 1]]
                       : P.int21[t2][t][S1[q + 1]][S1[i + 1]][S1[p - 1]];
    if (n1 == 2 && n2 == 2) return P.int22[t][t2][S1[i + 1]][S1[p - 1]][S1[q + 1]][S1[j - 1]];
    int nin = (nl - ns) * P.ninio2;
    if (nin > P.max_ninio) nin = P.max_ninio;
    return P.internal_loop[n1 + n2] + nin + P.mismatchI[t][S1[i + 1]][S1[j - 1]] +
           P.mismatchI[t2][S1[q + 1]][S1[p - 1]];
}

static inline int imin(int a, int b) { return a < b ? a : b; }

/* ------------------------------------------------------------------ A.3 fill */
static void fill_row(fold_t *f, int i)
{
    const int n = f->n, Ls = f->Ls, W = f->W;
    const short *S1 = f->S1;
    int jmax = imin(n, i + Ls);
    for (int j = i + TURN + 1; j <= jmax; j++) {
        int t = ptype(f, i, j), cij = INF;
        if (t) {
            cij = hairpin_E(f, i, j);
            int pmax = imin(j - 2 - TURN, i + MAXLOOP + 1);
            for (int p = i + 1; p <= pmax; p++) {
                int minq = j - i + p - MAXLOOP - 2;
                if (minq < p + 1 + TURN) minq = p + 1 + TURN;
                for (int q = minq; q < j; q++) {
                    int t2 = ptype(f, p, q);
                    if (!t2) continue;
                    cij = imin(cij, loop_E(f, i, j, p, q, t, P.rtype[t2]) + C_(p, q));
                }
            }
            int tt = P.rtype[t];
            int d3 = P.dangle3[tt][S1[i + 1]], d5 = P.dangle5[tt][S1[j - 1]];
            int dec = D_(i + 1, j - 1);
            dec = imin(dec, D_(i + 2, j - 1) + d3 + P.MLbase);
            dec = imin(dec, D_(i + 1, j - 2) + d5 + P.MLbase);
            dec = imin(dec, D_(i + 2, j - 2) + d3 + d5 + 2 * P.MLbase);
            cij = imin(cij, P.MLclosing + P.MLintern[t] + dec);
        }
        f->c[(size_t)i * W + (j - i)] = cij;

        int mij = M_(i + 1, j) + P.MLbase;
        mij = imin(mij, M_(i, j - 1) + P.MLbase);
        mij = imin(mij, cij + P.MLintern[t]);
        int ta = ptype(f, i + 1, j);
        mij = imin(mij, C_(i + 1, j) + P.dangle5[ta][S1[i]] + P.MLintern[ta] + P.MLbase);
        int tb = ptype(f, i, j - 1);
        mij = imin(mij, C_(i, j - 1) + P.dangle3[tb][S1[j]] + P.MLintern[tb] + P.MLbase);
        int tc = ptype(f, i + 1, j - 1);
        mij = imin(mij, C_(i + 1, j - 1) + P.dangle5[tc][S1[i]] + P.dangle3[tc][S1[j]] +
                            P.MLintern[tc] + 2 * P.MLbase);
        int dec = INF;
        for (int k = i + 1 + TURN; k <= j - 2 - TURN; k++) dec = imin(dec, M_(i, k) + M_(k + 1, j));
        f->dml[(size_t)i * W + (j - i)] = dec;
        mij = imin(mij, dec);
        f->m[(size_t)i * W + (j - i)] = mij;
    }
}

static void f3_row(fold_t *f, int i)
{
    const int n = f->n, Ls = f->Ls;
    const short *S1 = f->S1;
    int fi = F_(f, i + 1);
    for (int j = i + TURN + 1; j < n && j <= i + Ls; j++) {
        int t = ptype(f, i, j);
        if (t) {
            int e = C_(i, j) + AU(t);
            fi = imin(fi, e + F_(f, j + 1));
            fi = imin(fi, e + P.dangle3[t][S1[j + 1]] + F_(f, j + 2));
        }
        t = ptype(f, i + 1, j);
        if (t) {
            int e = C_(i + 1, j) + P.dangle5[t][S1[i]] + AU(t);
            fi = imin(fi, e + F_(f, j + 1));
            fi = imin(fi, e + P.dangle3[t][S1[j + 1]] + F_(f, j + 2));
        }
    }
    if (n <= i + Ls) {
        int t = ptype(f, i, n);
        if (t) fi = imin(fi, C_(i, n) + AU(t));
        t = ptype(f, i + 1, n);
        if (t) fi = imin(fi, C_(i + 1, n) + P.dangle5[t][S1[i]] + AU(t));
    }
    f->f3[i] = fi;
}

/* ------------------------------------------------------------------ A.4 traceback */
typedef struct { int i, j, ml; } sect_t;

/* returns malloc'ed finalised structure string, or NULL on "backtrack failed" */
static char *traceback(fold_t *f, int start, int md)
{
    const int n = f->n;
    const short *S1 = f->S1;
    int ndash = imin(n - start, md) + 1;
    char *st = (char *)calloc((size_t)ndash + 4, 1);
    memset(st, '-', (size_t)ndash);
    int cap = 4 * (ndash + 8), sp = 0;
    sect_t *stk = (sect_t *)malloc(sizeof(sect_t) * (size_t)cap);
#define PUSH(a, b, c_) do { stk[sp].i = (a); stk[sp].j = (b); stk[sp].ml = (c_); sp++; } while (0)
    PUSH(start, imin(n, start + md + 1), 0);
    while (sp > 0) {
        sp--;
        int i = stk[sp].i, j = stk[sp].j, ml = stk[sp].ml;
        if (j < i + TURN + 1) continue;
        if (ml == 0) {
            int fij = F_(f, i);
            if (fij == F_(f, i + 1)) { PUSH(i + 1, j, 0); continue; }
            int traced = 0, jj = 0, k;
            for (k = i + TURN + 1; k <= j; k++) {
                jj = k + 1;
                int t = ptype(f, i + 1, k);
                if (t) {
                    int cc = C_(i + 1, k) + P.dangle5[t][S1[i]] + AU(t);
                    if (fij == cc + F_(f, k + 1)) traced = i + 1;
                    if (k < n && fij == F_(f, k + 2) + cc + P.dangle3[t][S1[k + 1]]) { traced = i + 1; jj = k + 2; }
                }
                t = ptype(f, i, k);
                if (t) {
                    int cc = C_(i, k) + AU(t);
                    if (fij == cc + F_(f, k + 1)) traced = i;
                    if (k < n && fij == F_(f, k + 2) + cc + P.dangle3[t][S1[k + 1]]) { traced = i; jj = k + 2; }
                }
                if (traced) break;
            }
            if (!traced) goto fail;
            if (j == n) PUSH(jj, j, 0);
            i = traced; j = k;
            st[i - start] = '('; st[j - start] = ')';
            if (jj == j + 2 && j < n) st[j + 1 - start] = '.';
        } else {
            int fij = M_(i, j);
            if (M_(i, j - 1) + P.MLbase == fij) { PUSH(i, j - 1, 1); continue; }
            if (M_(i + 1, j) + P.MLbase == fij) { PUSH(i + 1, j, 1); continue; }
            int t = ptype(f, i, j);
            int cij = C_(i, j) + P.MLintern[t];
            t = ptype(f, i + 1, j);
            int ci1j = C_(i + 1, j) + P.dangle5[t][S1[i]] + P.MLintern[t] + P.MLbase;
            t = ptype(f, i, j - 1);
            int cij1 = C_(i, j - 1) + P.dangle3[t][S1[j]] + P.MLintern[t] + P.MLbase;
            t = ptype(f, i + 1, j - 1);
            int ci1j1 = C_(i + 1, j - 1) + P.dangle5[t][S1[i]] + P.dangle3[t][S1[j]] + P.MLintern[t] + 2 * P.MLbase;
            if (fij == cij || fij == ci1j || fij == cij1 || fij == ci1j1) {
                if (fij == ci1j) i++;
                else if (fij == cij1) j--;
                else if (fij == ci1j1) { i++; j--; }
                st[i - start] = '('; st[j - start] = ')';
            } else {
                int k;
                for (k = i + 1 + TURN; k <= j - 2 - TURN; k++)
                    if (fij == M_(i, k) + M_(k + 1, j)) break;
                if (k > j - 2 - TURN) goto fail;
                PUSH(i, k, 1);
                PUSH(k + 1, j, 1);
                continue;
            }
        }
        /* "repeat": (i,j) is a known pair; walk the stem/interior loops down */
        for (;;) {
            int cij = C_(i, j), t = ptype(f, i, j);
            if (cij == hairpin_E(f, i, j)) break;
            int found = 0;
            int pmax = imin(j - 2 - TURN, i + MAXLOOP + 1);
            for (int p = i + 1; p <= pmax && !found; p++) {
                int minq = j - i + p - MAXLOOP - 2;
                if (minq < p + 1 + TURN) minq = p + 1 + TURN;
                for (int q = j - 1; q >= minq; q--) {
                    int t2 = ptype(f, p, q);
                    if (!t2) continue;
                    if (cij == loop_E(f, i, j, p, q, t, P.rtype[t2]) + C_(p, q)) {
                        st[p - start] = '('; st[q - start] = ')';
                        i = p; j = q; found = 1;
                        break;
                    }
                }
            }
            if (found) continue;
            /* multiloop */
            int tt = P.rtype[t];
            int mm = P.MLclosing + P.MLintern[tt];
            int d5 = P.dangle5[tt][S1[j - 1]], d3 = P.dangle3[tt][S1[i + 1]];
            int i1 = i + 1, j1 = j - 1, k;
            for (k = i + 2 + TURN; k < j - 2 - TURN; k++) {
                if (cij == M_(i + 1, k) + M_(k + 1, j - 1) + mm) break;
                if (cij == M_(i + 2, k) + M_(k + 1, j - 1) + mm + d3 + P.MLbase) { i1 = i + 2; break; }
                if (cij == M_(i + 1, k) + M_(k + 1, j - 2) + mm + d5 + P.MLbase) { j1 = j - 2; break; }
                if (cij == M_(i + 2, k) + M_(k + 1, j - 2) + mm + d3 + d5 + 2 * P.MLbase) { i1 = i + 2; j1 = j - 2; break; }
            }
            if (k > j - 3 - TURN) goto fail;
            PUSH(i1, k, 1);
            PUSH(k + 1, j1, 1);
            break;
        }
        if (sp + 4 > cap) { cap *= 2; stk = (sect_t *)realloc(stk, sizeof(sect_t) * (size_t)cap); }
    }
    free(stk);
    {
        int k = (int)strlen(st) - 1;
        for (; k > 0 && st[k] == '-'; k--) st[k] = 0;
        for (; k >= 0; k--) if (st[k] == '-') st[k] = '.';
    }
    return st;
fail:
    free(stk);
    free(st);
    return NULL;
#undef PUSH
}

/* ------------------------------------------------------------------ public result type */
typedef struct {
    int n_hits, cap_hits;
    int *start;       /* 1-based printed start                                   */
    int *energy;      /* dcal: F(start) - F(start+len)                            */
    char **ss;        /* finalised structures, print order                        */
    int total;        /* F(1), dcal                                               */
    int failed;       /* nonzero if a traceback failed (reference would abort)    */
    char *conv;       /* converted sequence (uppercase, T->U)                     */
    int n;
    /* optional matrix dumps (kept when keep_matrices) */
    int W, Ls;
    int *c, *m, *f3;
} lfold_result;

static void add_hit(lfold_result *r, const fold_t *f, char *ss, int start)
{
    if (r->n_hits == r->cap_hits) {
        r->cap_hits = r->cap_hits ? 2 * r->cap_hits : 32;
        r->start = (int *)realloc(r->start, sizeof(int) * (size_t)r->cap_hits);
        r->energy = (int *)realloc(r->energy, sizeof(int) * (size_t)r->cap_hits);
        r->ss = (char **)realloc(r->ss, sizeof(char *) * (size_t)r->cap_hits);
    }
    int len = (int)strlen(ss);
    r->start[r->n_hits] = start;
    r->energy[r->n_hits] = F_(f, start) - F_(f, start + len);
    r->ss[r->n_hits] = strdup(ss);
    r->n_hits++;
}

/* Fold one raw sequence token (A.6 conversion, A.3 fill, A.5 emission). */
lfold_result *lfold_oracle_fold(const char *raw, int n, int L, int keep_matrices)
{
    params_init();
    lfold_result *r = (lfold_result *)calloc(1, sizeof *r);
    r->n = n;
    r->conv = (char *)malloc((size_t)n + 1);
    for (int k = 0; k < n; k++) {
        char ch = raw[k];
        if (ch >= 'a' && ch <= 'z') ch = (char)(ch - 32);
        if (ch == 'T') ch = 'U';
        r->conv[k] = ch;
    }
    r->conv[n] = 0;
    fold_t F, *f = &F;
    f->n = n;
    f->Ls = L < n ? L : n;
    f->W = f->Ls + 6;
    f->seq = r->conv;
    f->S = (short *)calloc((size_t)n + 3, sizeof(short));
    f->S1 = (short *)calloc((size_t)n + 3, sizeof(short));
    for (int k = 1; k <= n; k++) { f->S[k] = (short)enc(r->conv[k - 1]); f->S1[k] = (short)ALIAS[f->S[k]]; }
    size_t cells = (size_t)(n + 2) * (size_t)f->W;
    f->c = (int *)malloc(sizeof(int) * cells);
    f->m = (int *)malloc(sizeof(int) * cells);
    f->dml = (int *)malloc(sizeof(int) * cells);
    for (size_t k = 0; k < cells; k++) f->c[k] = f->m[k] = f->dml[k] = INF;
    f->f3 = (int *)calloc((size_t)n + 4, sizeof(int));

    /* A.5 emission state machine, static state is per-call here (one Lfold() per process
     * call sequence in the reference keeps statics across records, but they are always
     * left at do_bt=0, prev=NULL when i==1 completes). */
    int do_bt = 0, prev_i = 0;
    char *prev = NULL;
    for (int i = n - TURN - 1; i >= 1 && !r->failed; i--) {
        fill_row(f, i);
        f3_row(f, i);
        if (F_(f, i) != F_(f, i + 1)) do_bt = 1;
        else if (do_bt) {
            char *ss = traceback(f, i + 1, f->Ls + 1);
            if (!ss) { r->failed = 1; break; }
            if (prev) {
                size_t ls = strlen(ss), lp = strlen(prev);
                int off = prev_i - i;
                if ((size_t)i + ls < (size_t)prev_i + lp || strncmp(ss + off, prev, lp) != 0)
                    add_hit(r, f, prev, prev_i);
                free(prev);
            }
            prev = ss; prev_i = i + 1; do_bt = 0;
        }
        if (i == 1) {
            if (prev) { add_hit(r, f, prev, prev_i); free(prev); prev = NULL; }
            else do_bt = 1;
            if (do_bt) {
                char *ss = traceback(f, 1, f->Ls);
                if (!ss) { r->failed = 1; break; }
                add_hit(r, f, ss, 1);
                free(ss);
            }
            do_bt = 0;
        }
    }
    free(prev);
    r->total = n >= 1 ? F_(f, 1) : 0;
    if (keep_matrices) { r->W = f->W; r->Ls = f->Ls; r->c = f->c; r->m = f->m; r->f3 = f->f3; }
    else { free(f->c); free(f->m); free(f->f3); }
    free(f->dml); free(f->S); free(f->S1);
    return r;
}

void lfold_oracle_free(lfold_result *r)
{
    if (!r) return;
    for (int k = 0; k < r->n_hits; k++) free(r->ss[k]);
    free(r->ss); free(r->start); free(r->energy); free(r->conv);
    free(r->c); free(r->m); free(r->f3);
    free(r);
}

/* flat accessors for ctypes */
int lfold_oracle_nhits(const lfold_result *r) { return r->n_hits; }
int lfold_oracle_total(const lfold_result *r) { return r->total; }
int lfold_oracle_failed(const lfold_result *r) { return r->failed; }
int lfold_oracle_hit_start(const lfold_result *r, int k) { return r->start[k]; }
int lfold_oracle_hit_energy(const lfold_result *r, int k) { return r->energy[k]; }
const char *lfold_oracle_hit_ss(const lfold_result *r, int k) { return r->ss[k]; }
const char *lfold_oracle_conv(const lfold_result *r) { return r->conv; }
int lfold_oracle_W(const lfold_result *r) { return r->W; }
/* band matrices, [i][d] with row stride W (1-based i), INF where not computed */
const int *lfold_oracle_c(const lfold_result *r) { return r->c; }
const int *lfold_oracle_m(const lfold_result *r) { return r->m; }
const int *lfold_oracle_f3(const lfold_result *r) { return r->f3; }

/* ------------------------------------------------------------------ A.6 program I/O */
/* Reads multi-FASTA-ish text from `in`, writes RNALfold-identical text to `out`. */
int lfold_oracle_stream(FILE *in, FILE *out, int L)
{
    char *line = NULL;
    size_t cap = 0;
    ssize_t len;
    for (;;) {
        if ((len = getline(&line, &cap, in)) < 0) break;
        if (len && line[len - 1] == '\n') line[--len] = 0;
        while (line[0] == '*' || line[0] == 0 || line[0] == '>') {
            fprintf(out, "%s\n", line);
            if ((len = getline(&line, &cap, in)) < 0) goto done;
            if (len && line[len - 1] == '\n') line[--len] = 0;
        }
        if (strcmp(line, "@") == 0) break;
        /* first whitespace-delimited token */
        char *tok = line;
        while (*tok == ' ' || *tok == '\t' || *tok == '\r' || *tok == '\v' || *tok == '\f') tok++;
        int n = 0;
        while (tok[n] && tok[n] != ' ' && tok[n] != '\t' && tok[n] != '\r' && tok[n] != '\v' && tok[n] != '\f' && tok[n] != '\n') n++;
        lfold_result *r = lfold_oracle_fold(tok, n, L, 0);
        if (r->failed) { fprintf(stderr, "backtrack failed\n"); lfold_oracle_free(r); free(line); return 1; }
        for (int k = 0; k < r->n_hits; k++)
            fprintf(out, "%s (%6.2f) %4d\n", r->ss[k], r->energy[k] / 100., r->start[k]);
        fprintf(out, "%s\n (%6.2f)\n", r->conv, r->total / 100.);
        lfold_oracle_free(r);
    }
done:
    free(line);
    return 0;
}

#ifdef LFOLD_ORACLE_MAIN
int main(int argc, char **argv)
{
    int L = 150;  /* RNALfold default */
    for (int a = 1; a < argc; a++)
        if (!strcmp(argv[a], "-L") && a + 1 < argc) L = atoi(argv[++a]);
    return lfold_oracle_stream(stdin, stdout, L);
}
#endif
