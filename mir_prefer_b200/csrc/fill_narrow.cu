// fill_narrow.cu -- K2, the narrow (16-bit) band-fill kernel as a per-CTA task queue, for sm_100a.
//
// Same arithmetic as fill.cu's kernels (RNALfold fill_arrays, SURVEY.md 8a rows a5-a6, Appendix
// A.2/A.3, "RLF Lfold.c:189-346"); one CTA owns one locus and walks the band two anti-diagonals at
// a time.  The dependencies of the recursion leave more slack than a diagonal-by-diagonal sweep uses:
//   * the interior-loop search of a cell on diagonal d (bulges >= 2 nt and the generic loops) only
//     reads c of diagonals <= d-4; the seven table-driven small loops, the hairpin and the multiloop
//     closing read c of diagonals <= d-2 and DML of diagonals d-2..d-4;
//   * fML(d) reads c(d), c(d-1), c(d-2), fML(d-1) and DML(d); DML(d) reads fML of diagonals <= d-5.
// So in step X (even) five kinds of work are independent of each other and run concurrently:
//   P  interior-loop minima of the typed cells of diagonals X, X+1      (warp per cell, LDS + VIADDMNMX.S16x2)
//   T  small loops / hairpin / multiloop closing + stores of X-2, X-1   (thread per typed cell)
//   R  fML of diagonals X-4, X-3; INF for the untyped cells of X-2, X-1; typed lists of X+2, X+3
//   A  DML of diagonals X-1, X from 16-bit fML row pairs                (thread per row pair)
//   misc  ring offsets of the next step, counters
// Every kind is cut into warp-sized items; the 16 (12, 8) warps of the CTA pull items from two
// shared-memory queues (P | everything else; half of the warps start on each and switch when
// theirs runs dry) and meet at ONE __syncthreads per two diagonals.
//
// The interior-loop window is the 16-bit pair ring of fill.cu's k_fill_s16, with 16 pair slots:
// sG = c + mismatchI (generic loops), sB = c + AU (bulges, and the exact c of the small loops and of
// fML).  Pair pp = d'>>1 lives in slot pp & 15, rotated by (17 pp) & 31 words; lane l owns the
// word-terms of bank class l (DevParams::s16_*), so a typed cell costs 10 conflict-free LDS + 10
// VIADDMNMX.S16x2 per lane.  Exact while c > MF16_GUARD and fML > MF16M_GUARD; otherwise the locus
// is flagged and redone by the 32-bit kernel (k_fill_smem).
#include "mirfold_internal.cuh"
#include "fill_common.cuh"

#define MFQ_NPS 16     /* pair slots of the ring: rows X-32..X-3 are read while X-2, X-1 are written */
#define MFQ_PC 8       /* typed cells per P item                                                       */
#define MFQ_EPART 32   /* split positions per bulk DML item                                            */
#define MFQ_SW 8       /* DML strip width (diagonals per bulk pass)                                    */
#define FULLMASK 0xffffffffu

template <int NS>
struct FillQSmem {
    static constexpr int RS = NS + 32;
    static constexpr int RW = MFQ_NPS * RS;             // words per ring
    // word offsets into the dynamic shared-memory window
    static constexpr int oG = 0;
    static constexpr int oInf = RW;                     // one all-INF row (terms a lane does not have)
    static constexpr int oB = RW + RS;
    static constexpr int oList = oB + RW;               // [3][2][NS] words: i*4 | (AU - mismatchI + bias) << 16
    static constexpr int oMrow = oList + 6 * NS;        // [2][NS] ints: fML of the newest odd diagonal
    static constexpr int oMM = oMrow + 2 * NS;          // mismatchI[200]
    static constexpr int oMy = oMM + 200;               // [2][2][NS] shorts: interior-loop minimum per listed cell
    static constexpr int oOff = oMy + 2 * NS;           // [2][2][NQ][32] ushorts: word offsets of the word-terms
    static constexpr int oS = oOff + 2 * MF16_NQ * 32;  // bytes from here: sS[NS+8] | sS1[NS+8] | pair[64]
    static constexpr size_t bytes = (size_t)oS * 4 + 2 * (NS + 8) + 64 + 16;
};

// word offset of the word-term `td` of lane `lane` for the step whose pair base is ppb = (d-2)>>1
template <int NS>
__device__ __forceinline__ unsigned q_term_off(unsigned td, int lane, int ppb)
{
    using SM = FillQSmem<NS>;
    if (td >> 11) return (unsigned)(SM::oInf + ((lane + MF16_SKEW * ppb) & 31));
    const int m = td & 15, xo = (td >> 4) & 63, pp = ppb - m;
    if (pp < 0) return (unsigned)(SM::oInf + xo);   // d < 34 only: the inner diagonal does not exist
    return (unsigned)((((td >> 10) & 1) ? SM::oB : SM::oG) + (pp & (MFQ_NPS - 1)) * SM::RS + ((MF16_SKEW * pp) & 31) + xo);
}
__device__ __forceinline__ void q_ring_put(unsigned int *ring, int RS, int d, int x, int v)
{
    const int pp = d >> 1;
    unsigned short *w = (unsigned short *)(ring + (pp & (MFQ_NPS - 1)) * RS + ((MF16_SKEW * pp) & 31) + x);
    w[d & 1] = (unsigned short)v;
}
__device__ __forceinline__ int q_ring_get(const unsigned int *ring, int RS, int dd, int x)
{
    const int pp = dd >> 1;
    const short *w = (const short *)(ring + (pp & (MFQ_NPS - 1)) * RS + ((MF16_SKEW * pp) & 31) + x);
    return w[dd & 1];
}

// T: the seven table-driven two-loops + hairpin + d1 multiloop closing of a typed cell (A.2, A.3);
// branch-free (all lookups are independent and overlap), c(p,q) from the bulge ring (c + AU).
template <class StrideT>
__device__ __forceinline__ int q_cell_tail(const DevParams *__restrict__ P, const unsigned char *sS,
                                           const unsigned char *sS1, const unsigned char *sPair,
                                           const unsigned int *sB, int RS, const int *rD, StrideT NS, int i,
                                           int d, int t, int si1, int sj1, int K)
{
    const int j = i + d;
    const int AUp = P->TerminalAU;
    int best = MF_INF;
#pragma unroll
    for (int m = 0; m < 7; m++) {
        const int u = (m == 1 || m == 3 || m == 4) ? 1 : (m >= 5 ? 2 : 0);
        const int v = (m == 2 || m == 3 || m == 5) ? 1 : ((m == 4 || m == 6) ? 2 : 0);
        // m: 0 (0,0)  1 (1,0)  2 (0,1)  3 (1,1)  4 (1,2)  5 (2,1)  6 (2,2)
        const bool ok = (u + v <= K);
        const int p = ok ? i + 1 + u : i + 1, q = ok ? j - 1 - v : j - 1;   // clamped: always a legal address
        const int t2 = sPair[sS[p] * 8 + sS[q]];
        const int r2 = dev_rtype(t2);
        const int c2 = q_ring_get(sB, RS, ok ? d - 2 - u - v : d - 2, p - 1) - (t2 > 2 ? AUp : 0);
        const int sp1 = sS1[p - 1], sq1 = sS1[q + 1];
        int e;
        if (m == 0) e = P->stack[t * 8 + r2];
        else if (m == 1 || m == 2) e = P->bulge[1] + P->stack[t * 8 + r2];
        else if (m == 3) e = P->int11[((t * 8 + r2) * 5 + si1) * 5 + sj1];
        else if (m == 4) e = P->int21[(((t * 8 + r2) * 5 + si1) * 5 + sq1) * 5 + sj1];           // n1 = 1, n2 = 2
        else if (m == 5) e = P->int21[(((r2 * 8 + t) * 5 + sq1) * 5 + si1) * 5 + sp1];           // n1 = 2, n2 = 1
        else e = P->int22[((((t * 8 + r2) * 5 + si1) * 5 + sp1) * 5 + sq1) * 5 + sj1];
        if (ok && t2) best = min(best, e + c2);
    }
    best = min(best, dev_hairpin(P, sS, sS1, i, j, t));
    const int tt = dev_rtype(t);
    const int d3 = P->dangle3[tt * 5 + si1], d5 = P->dangle5[tt * 5 + sj1];
    int dec = MF_INF;
    if (d - 2 >= 4) dec = rD[((d - 2) & (MF_RING_DML - 1)) * NS + i];                          // DML(i+1,j-1)
    if (d - 3 >= 4) {
        dec = min(dec, rD[((d - 3) & (MF_RING_DML - 1)) * NS + i + 1] + d3);                   // DML(i+2,j-1)
        dec = min(dec, rD[((d - 3) & (MF_RING_DML - 1)) * NS + i] + d5);                       // DML(i+1,j-2)
    }
    if (d - 4 >= 4) dec = min(dec, rD[((d - 4) & (MF_RING_DML - 1)) * NS + i + 1] + d3 + d5);  // DML(i+2,j-2)
    return min(best, P->MLclosing + P->MLintern[t] + dec);
}

// R: fML(i, i+d) from its six boundary terms + DML (A.3).  mA = fML(i+1,j), mB = fML(i,j-1) (only read
// when d >= 5); c of the three newest diagonals comes from the bulge ring (exact inside the guarded range).
template <class StrideT>
__device__ __forceinline__ int q_fml(const DevParams *__restrict__ P, const unsigned char *sS, const unsigned char *sS1,
                                     const unsigned char *sPair, const unsigned int *sB, int RS, int mA, int mB,
                                     const int *rD, StrideT NS, int i, int d, int Ls)
{
    const int j = i + d;
    const int AUp = P->TerminalAU;
    const int t = (d < Ls) ? sPair[sS[i] * 8 + sS[j]] : 0;
    int m = MF_INF;
    if (d - 1 >= 4) {
        m = min(mA, mB);
        const int ta = sPair[sS[i + 1] * 8 + sS[j]];                // (i+1, j)
        if (ta) m = min(m, q_ring_get(sB, RS, d - 1, i) - (ta > 2 ? AUp : 0) + P->dangle5[ta * 5 + sS1[i]] + P->MLintern[ta]);
        const int tb = sPair[sS[i] * 8 + sS[j - 1]];                // (i, j-1)
        if (tb) m = min(m, q_ring_get(sB, RS, d - 1, i - 1) - (tb > 2 ? AUp : 0) + P->dangle3[tb * 5 + sS1[j]] + P->MLintern[tb]);
    }
    if (t) m = min(m, q_ring_get(sB, RS, d, i - 1) - (t > 2 ? AUp : 0) + P->MLintern[t]);
    if (d - 2 >= 4) {
        const int tc = sPair[sS[i + 1] * 8 + sS[j - 1]];            // (i+1, j-1)
        if (tc) m = min(m, q_ring_get(sB, RS, d - 2, i) - (tc > 2 ? AUp : 0) + P->dangle5[tc * 5 + sS1[i]] +
                               P->dangle3[tc * 5 + sS1[j]] + P->MLintern[tc]);
    }
    return min(m, rD[(d & (MF_RING_DML - 1)) * NS + i - 1]);
}

// store fML(i, i+d): 32-bit band + the 16-bit row-pair copy the DML items read (lo half of word
// i-1, hi half of word i-2)
__device__ __forceinline__ void q_store_fml(int *Mb, unsigned int *Mp, int NS, int d, int i, int m, int *sFlag)
{
    Mb[(d - 4) * NS + i - 1] = m;
    const int m16 = (m >= MF_INF / 2) ? MF16M_INF : max(m, -32768);
    if (m < MF16M_GUARD) *sFlag = 1;
    unsigned short *w = (unsigned short *)(Mp + (d - 4) * NS + i - 1);
    w[0] = (unsigned short)m16;
    if (i >= 2) w[-1] = (unsigned short)m16;
}

template <int NS, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_fill_q16(FillLaunch a)
{
    using SM = FillQSmem<NS>;
    constexpr int RS = SM::RS;
    constexpr int NW = NT / 32;
    constexpr int NWP = NW / 2;                         // warps that start on the P queue
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned int *sW = (unsigned int *)smem_raw;
    unsigned int *sG = sW + SM::oG, *sB = sW + SM::oB;
    unsigned int *sList = sW + SM::oList;
    int *sMrow = (int *)(sW + SM::oMrow);
    int *sMM = (int *)(sW + SM::oMM);
    short *sMy = (short *)(sW + SM::oMy);
    unsigned short *sOff = (unsigned short *)(sW + SM::oOff);
    unsigned char *sS = (unsigned char *)(sW + SM::oS);
    unsigned char *sS1 = sS + NS + 8;
    unsigned char *sPair = sS1 + NS + 8;
    __shared__ int sCnt[4][2];    // typed cells per listed diagonal, slot (X/2) & 3
    __shared__ int sNext[2][4];   // queue heads, [step parity][queue]
    __shared__ int sFlag;

    const LocusDesc L = a.loci[blockIdx.x];
    const DevParams *__restrict__ P = a.P;
    const int tid = threadIdx.x, lane = tid & 31;
    // Everything that steers a warp through the queues goes through redux.sync: its result lives in a
    // uniform register, so the compiler knows the branches are warp-uniform (no divergence handling, and
    // the P loop's ring loads become LDS [R + UR]).
    const int wid = __reduce_min_sync(FULLMASK, tid >> 5);
    const int n = __reduce_min_sync(FULLMASK, L.n), Ls = __reduce_min_sync(FULLMASK, L.Ls), dmax = __reduce_min_sync(FULLMASK, L.dmax);
    const int AUp = P->TerminalAU;

    for (int k = tid; k < NS + 8; k += NT) {
        const unsigned char b = (k < n + 3) ? a.codes[L.seq_off + k] : 0;
        sS[k] = b & 7;
        sS1[k] = b >> 4;
    }
    for (int k = tid; k < SM::oList; k += NT) sW[k] = MF16_INF2;   // both rings + the INF row
    for (int k = tid; k < 200; k += NT) sMM[k] = P->mismatchI[k];
    if (tid < 64) sPair[tid] = P->pair[tid];
    if (tid < 8) (&sCnt[0][0])[tid] = 0;
    if (tid < 8) (&sNext[0][0])[tid] = 0;
    if (tid == 0) sFlag = 0;

    int *Cb = a.C + L.band_off;
    int *Mb = a.M + L.band_off;
    unsigned int *Mp = a.Mp + L.band_off;
    int *rD = a.ring + L.ring_off;   // [MF_RING_DML][NS]
    for (int k = tid; k < MF_RING_DML * NS; k += NT) rD[k] = MF_INF;
    __syncthreads();

    // typed lists of diagonals 4, 5 (list buffer 2, count slot 2) and the term offsets of step 4
    for (int r = 4; r <= min(5, dmax); r++) {
        unsigned int *list = sList + (2 * 2 + (r & 1)) * NS;
        for (int i = tid + 1; i <= n - r; i += NT) {
            const int t = (r < Ls) ? sPair[sS[i] * 8 + sS[i + r]] : 0;
            if (t) {
                const int dl = (t > 2 ? AUp : 0) - sMM[(t * 5 + sS1[i + 1]) * 5 + sS1[i + r - 1]] + MF16_DBIAS;
                list[atomicAdd(&sCnt[2][r & 1], 1)] = (unsigned)(i * 4) | ((unsigned)dl << 16);
            }
        }
    }
    for (int k = tid; k < 2 * MF16_NQ * 32; k += NT) sOff[k] = (unsigned short)q_term_off<NS>((&P->s16_td[0][0][0])[k], k & 31, 1);
    __syncthreads();

    const unsigned smem_base = (unsigned)__cvta_generic_to_shared(smem_raw);

    for (int X = 4; X - 4 <= dmax; X += 2) {
        const int h = X >> 1;
        // ---- item counts of this step (uniform over the CTA)
        const int cP0 = __reduce_min_sync(FULLMASK, (X <= dmax) ? sCnt[h & 3][0] : 0);
        const int cP1 = __reduce_min_sync(FULLMASK, (X + 1 <= dmax) ? sCnt[h & 3][1] : 0);
        const int cT0 = __reduce_min_sync(FULLMASK, (X - 2 >= 4 && X - 2 <= dmax) ? sCnt[(h - 1) & 3][0] : 0);
        const int cT1 = __reduce_min_sync(FULLMASK, (X - 2 >= 4 && X - 1 <= dmax) ? sCnt[(h - 1) & 3][1] : 0);
        const int nP0 = (cP0 + MFQ_PC - 1) / MFQ_PC, nP1 = (cP1 + MFQ_PC - 1) / MFQ_PC;
        // DML: step X with (X-10) % SW == 0 computes the bulk of the strip X-1 .. X+SW-2 (all split terms whose
        // two fML operands lie on diagonals <= X-5, the newest complete one); every other step finalises its
        // own two diagonals X-1, X with the few terms that involve the diagonals completed since.
        const int sA = (X >= 10 && X - 1 <= dmax) ? ((X - 10) & (MFQ_SW - 1)) : -1;
        const int nparts = (sA == 0) ? (X - 8 + MFQ_EPART - 1) / MFQ_EPART : 1;
        const int nchunkA = (sA >= 0) ? ((n - (X - 1) + 1) / 2 + 31) / 32 : 0;
        const int nA = nparts * nchunkA;
        const int nR = (n - max(4, X - 4) + 30) / 31;
        const int nT0 = (cT0 + 31) / 32, nT1 = (cT1 + 31) / 32;
        const int bA = nA, bR = bA + nR, bT0 = bR + nT0, bT1 = bT0 + nT1, total2 = bT1 + 1;
        const int lbP = h % 3, lbT = (h + 2) % 3, lbL = (h + 1) % 3;

        // queues: 0 = P items of diagonal X, 1 = P items of diagonal X+1, 2 = everything else.  The first
        // half of the warps takes them in the order 0,1,2, the other half 2,0,1.
        for (int pass = 0; pass < 3; pass++) {
            const int q = (wid < NWP) ? pass : (pass == 0 ? 2 : pass - 1);
            if (q < 2) {
                const int par = q;
                const int cnt = par ? cP1 : cP0;
                const int total = par ? nP1 : nP0;
                if (__reduce_min_sync(FULLMASK, *(volatile int *)&sNext[h & 1][q]) >= total) continue;
                // per-lane word-term addresses (bytes, shared window), packed constants and masks of this parity
                unsigned off[MF16_NQ], cst[MF16_NQ], mk[MF16_NMK];
#pragma unroll
                for (int qq = 0; qq < MF16_NQ; qq++) {
                    off[qq] = smem_base + 4u * sOff[(((h & 1) * 2 + par) * MF16_NQ + qq) * 32 + lane];
                    asm("" : "+r"(off[qq]));   // keep the byte address as one register (no re-association in the cell loop)
                    cst[qq] = P->s16_cst[par][qq][lane];
                }
#pragma unroll
                for (int qq = 0; qq < MF16_NMK; qq++) mk[qq] = P->s16_mk[par][qq][lane];
                unsigned listb = smem_base + 4u * (unsigned)(SM::oList + (lbP * 2 + par) * NS);       // shared byte addresses
                unsigned myb = smem_base + 4u * (unsigned)SM::oMy + 2u * (unsigned)(((h & 1) * 2 + par) * NS);
                asm volatile("" : "+r"(listb), "+r"(myb));   // opaque: kept in registers, not rematerialised in the cell loop
                for (;;) {
                    int k = 0;
                    if (lane == 0) k = atomicAdd(&sNext[h & 1][q], 1);
                    k = __reduce_max_sync(FULLMASK, k);
                    if (k >= total) break;
                    // ------------------------------------------------------------ P item
                    const int c0 = k * MFQ_PC, c1 = __reduce_min_sync(FULLMASK, min(cnt, c0 + MFQ_PC));
                    for (int c = c0; c < c1; c += 2) {
                        // an odd tail evaluates its last cell twice (same value stored to my[c] and my[c+1], which is unused)
                        const unsigned ea = dev_lds(listb + 4u * c), eb = dev_lds(listb + 4u * min(c + 1, c1 - 1));
                        __syncwarp();   // tells the compiler the warp is converged: redux without a divergence check
                        // the cell's byte offset is warp-uniform: redux puts it into a uniform register and the
                        // ten loads become LDS [R + UR] with no per-load address arithmetic
                        const unsigned ia = (unsigned)__reduce_min_sync(FULLMASK, (int)(ea & 0xffffu));
                        const unsigned ib = (unsigned)__reduce_min_sync(FULLMASK, (int)(eb & 0xffffu));
                        unsigned aG = MF16_INF2, aB = MF16_INF2, bG = MF16_INF2, bB = MF16_INF2;
#pragma unroll
                        for (int qq = 0; qq < MF16_NQG; qq++) {
                            unsigned wa = dev_lds(off[qq] + ia), wb = dev_lds(off[qq] + ib);
                            if (qq < MF16_NMG) {
                                wa = (wa & mk[qq]) | (~mk[qq] & MF16_INF2);
                                wb = (wb & mk[qq]) | (~mk[qq] & MF16_INF2);
                            }
                            aG = __viaddmin_s16x2(wa, cst[qq], aG);
                            bG = __viaddmin_s16x2(wb, cst[qq], bG);
                        }
#pragma unroll
                        for (int qq = 0; qq < MF16_NQB; qq++) {
                            unsigned wa = dev_lds(off[MF16_NQG + qq] + ia), wb = dev_lds(off[MF16_NQG + qq] + ib);
                            wa = (wa & mk[MF16_NMG + qq]) | (~mk[MF16_NMG + qq] & MF16_INF2);
                            wb = (wb & mk[MF16_NMG + qq]) | (~mk[MF16_NMG + qq] & MF16_INF2);
                            aB = __viaddmin_s16x2(wa, cst[MF16_NQG + qq], aB);
                            bB = __viaddmin_s16x2(wb, cst[MF16_NQG + qq], bB);
                        }
                        const unsigned acca = __viaddmin_s16x2(aB, __byte_perm(ea, 0, 0x3232), aG);   // relative to the outer mismatch
                        const unsigned accb = __viaddmin_s16x2(bB, __byte_perm(eb, 0, 0x3232), bG);
                        int va = min((int)(short)(acca & 0xffffu), (int)acca >> 16);
                        int vb = min((int)(short)(accb & 0xffffu), (int)accb >> 16);
                        va = warp_min(va);
                        vb = warp_min(vb);
                        if (lane == 0) dev_sts16x2(myb + 2u * c, va, vb);
                    }
                }
                continue;
            }
            for (;;) {
                int k = 0;
                if (lane == 0) k = atomicAdd(&sNext[h & 1][2], 1);
                k = __reduce_max_sync(FULLMASK, k);
                if (k >= total2) break;
                if (k < bA) {
                    // ------------------------------------------------------------ A item
                    const int chunk = k / nparts, part = k - chunk * nparts;
                    const int i = 2 * (chunk * 32 + lane) + 1;       // rows i (lo half) and i+1 (hi half)
                    if (i <= n - (X - 1)) {
                        if (sA == 0) {
                            // bulk of the strip: row s is diagonal X-1+s and takes the split positions e in
                            // [max(4, s+3), X-5] (row 0: X-6), i.e. fML(i,i+e) + fML(i+e+1, i+X-1+s) with both spans <= X-5
                            const int e0 = 4 + part * MFQ_EPART, e1 = min(X - 5, e0 + MFQ_EPART - 1);
                            const unsigned int *pa = Mp + (e0 - 4) * NS + (i - 1);       // fML(i, i+e) | fML(i+1, i+1+e)
                            const unsigned int *pb = Mp + (X - 6 - e0) * NS + (i + e0);  // row s: + s*NS
                            unsigned acc[MFQ_SW];
#pragma unroll
                            for (int sr = 0; sr < MFQ_SW; sr++) acc[sr] = MF16M_INF2;
                            int e = e0;
                            for (; e <= min(e1, MFQ_SW + 1); e++) {                      // head: rows start at e = s+3
                                const unsigned av = pa[0];
#pragma unroll
                                for (int sr = 0; sr < MFQ_SW; sr++)
                                    if (e >= sr + 3 && (sr > 0 || e <= X - 6)) acc[sr] = __viaddmin_s16x2(av, pb[sr * NS], acc[sr]);
                                pa += NS;
                                pb -= (NS - 1);
                            }
                            const int emain = min(e1, X - 6);
#pragma unroll 2
                            for (; e <= emain; e++) {
                                const unsigned av = pa[0];
#pragma unroll
                                for (int sr = 0; sr < MFQ_SW; sr++) acc[sr] = __viaddmin_s16x2(av, pb[sr * NS], acc[sr]);
                                pa += NS;
                                pb -= (NS - 1);
                            }
                            if (e <= e1) {                                               // e = X-5: not row 0
                                const unsigned av = pa[0];
#pragma unroll
                                for (int sr = 1; sr < MFQ_SW; sr++) acc[sr] = __viaddmin_s16x2(av, pb[sr * NS], acc[sr]);
                            }
#pragma unroll
                            for (int sr = 0; sr < MFQ_SW; sr++) {
                                const int dd = X - 1 + sr;
                                if (dd <= dmax) {
                                    int *dst = &rD[(dd & (MF_RING_DML - 1)) * NS + (i - 1)];
                                    const int lo = (int)(short)(acc[sr] & 0xffffu), hi = (int)acc[sr] >> 16;
                                    if (i <= n - dd && lo < MF16M_VALID) atomicMin(dst, lo);
                                    if (i + 1 <= n - dd && hi < MF16M_VALID) atomicMin(dst + 1, hi);
                                }
                            }
                        } else {
                            // finalise diagonals X-1, X: the strip's bulk (step X - sA) stopped at fML diagonal D
                            const int D = X - sA - 5;
#pragma unroll
                            for (int r = 0; r < 2; r++) {
                                const int dd = X - 1 + r;
                                if (dd <= dmax) {
                                    const int eh = dd - 2 - D;
                                    unsigned acc = MF16M_INF2;
                                    for (int e = 4; e <= eh; e++) {
                                        // fML(i,i+e) + fML(i+e+1,i+dd)   and the mirrored split e' = dd-e-1
                                        acc = __viaddmin_s16x2(Mp[(e - 4) * NS + i - 1], Mp[(dd - e - 5) * NS + i + e], acc);
                                        acc = __viaddmin_s16x2(Mp[(dd - e - 5) * NS + i - 1], Mp[(e - 4) * NS + i + dd - e - 1], acc);
                                    }
                                    int *dst = &rD[(dd & (MF_RING_DML - 1)) * NS + (i - 1)];
                                    const int lo = (int)(short)(acc & 0xffffu), hi = (int)acc >> 16;
                                    if (i <= n - dd && lo < MF16M_VALID) atomicMin(dst, lo);
                                    if (i + 1 <= n - dd && hi < MF16M_VALID) atomicMin(dst + 1, hi);
                                }
                            }
                        }
                    }
                } else if (k < bR) {
                    // ------------------------------------------------------------ R item: rows 31*chunk+1 .. +32
                    const int i = 31 * (k - bA) + lane + 1;
                    const int d0 = X - 4, d1 = X - 3;
                    int m0 = MF_INF;
                    if (d0 >= 4 && i <= n - d0) {
                        const int *Mprev = sMrow + ((h + 1) & 1) * NS;       // fML of diagonal X-5
                        int mA = MF_INF, mB = MF_INF;
                        if (d0 >= 5) { mA = Mprev[i]; mB = Mprev[i - 1]; }
                        m0 = q_fml(P, sS, sS1, sPair, sB, RS, mA, mB, rD, NS, i, d0, Ls);
                        q_store_fml(Mb, Mp, NS, d0, i, m0, &sFlag);
                    }
                    const int up = __shfl_down_sync(FULLMASK, m0, 1);        // fML(i+1, i+1+d0)
                    if (lane < 31) {
                        if (d1 >= 5 && d1 <= dmax && i <= n - d1) {
                            const int m1 = q_fml(P, sS, sS1, sPair, sB, RS, up, m0, rD, NS, i, d1, Ls);
                            q_store_fml(Mb, Mp, NS, d1, i, m1, &sFlag);
                            sMrow[(h & 1) * NS + i - 1] = m1;
                        }
                        // INF for the untyped cells of the diagonals the T items of this step fill
#pragma unroll
                        for (int s = 0; s < 2; s++) {
                            const int r = X - 2 + s;
                            if (r >= 4 && r <= dmax && i <= n - r) {
                                const int t = (r < Ls) ? sPair[sS[i] * 8 + sS[i + r]] : 0;
                                if (!t) {
                                    q_ring_put(sG, RS, r, i - 1, MF16_INF);
                                    q_ring_put(sB, RS, r, i - 1, MF16_INF);
                                    Cb[(r - 4) * NS + i - 1] = MF_INF;
                                }
                            }
                        }
                        // DML rows of the strip whose bulk the next step computes
                        if (X >= 8 && ((X - 8) & (MFQ_SW - 1)) == 0 && i <= n - (X + 1)) {
#pragma unroll
                            for (int sr = 0; sr < MFQ_SW; sr++) rD[((X + 1 + sr) & (MF_RING_DML - 1)) * NS + i - 1] = MF_INF;
                        }
                    }
                    // typed lists of the diagonals the P items of the next step search
#pragma unroll
                    for (int s = 0; s < 2; s++) {
                        const int r = X + 2 + s;
                        if (r <= dmax) {                                       // uniform
                            int t = 0;
                            if (lane < 31 && i <= n - r && r < Ls) t = sPair[sS[i] * 8 + sS[i + r]];
                            const unsigned bal = __ballot_sync(FULLMASK, t != 0);
                            if (bal) {
                                int base = 0;
                                if (lane == 0) base = atomicAdd(&sCnt[(h + 1) & 3][s], __popc(bal));
                                base = __shfl_sync(FULLMASK, base, 0);
                                if (t) {
                                    const int dl = (t > 2 ? AUp : 0) - sMM[(t * 5 + sS1[i + 1]) * 5 + sS1[i + r - 1]] + MF16_DBIAS;
                                    sList[(lbL * 2 + s) * NS + base + __popc(bal & ((1u << lane) - 1))] =
                                        (unsigned)(i * 4) | ((unsigned)dl << 16);
                                }
                            }
                        }
                    }
                } else if (k < bT1) {
                    // ------------------------------------------------------------ T item: 32 typed cells of X-2 / X-1
                    const int s = (k >= bT0) ? 1 : 0;
                    const int d = X - 2 + s;
                    const int c = ((s ? k - bT0 : k - bR) << 5) + lane;
                    if (c < (s ? cT1 : cT0)) {
                        const int K = min(30, d - 6);
                        const int i = (int)(sList[(lbT * 2 + s) * NS + c] & 0xffffu) >> 2, j = i + d;
                        const int my = sMy[(((h + 1) & 1) * 2 + s) * NS + c];
                        const int t = sPair[sS[i] * 8 + sS[j]];
                        const int si1 = sS1[i + 1], sj1 = sS1[j - 1];
                        int best = (my < MF16_VALID) ? my + sMM[(t * 5 + si1) * 5 + sj1] : MF_INF;
                        best = min(best, q_cell_tail(P, sS, sS1, sPair, sB, RS, rD, NS, i, d, t, si1, sj1, K));
                        const int tt = dev_rtype(t);
                        const int mm = sMM[(tt * 5 + sS1[j + 1]) * 5 + sS1[i - 1]];
                        Cb[(d - 4) * NS + i - 1] = best;
                        if (best < MF16_GUARD) sFlag = 1;
                        q_ring_put(sG, RS, d, i - 1, max(best + mm, -32768));
                        q_ring_put(sB, RS, d, i - 1, max(best + (tt > 2 ? AUp : 0), -32768));
                    }
                } else {
                    // ------------------------------------------------------------ misc: next step's term offsets, counters
                    for (int kk = lane; kk < 2 * MF16_NQ * 32; kk += 32)
                        sOff[((h + 1) & 1) * 2 * MF16_NQ * 32 + kk] = (unsigned short)q_term_off<NS>((&P->s16_td[0][0][0])[kk], kk & 31, h);
                    if (lane < 4) sNext[(h + 1) & 1][lane] = 0;
                    if (lane < 2) sCnt[(h + 2) & 3][lane] = 0;
                }
            }
        }
        __syncthreads();
    }
    if (tid == 0 && sFlag) a.flags[blockIdx.x] = 1;
}

template <int NS, int NT, int MINB>
static cudaError_t configure_q_bucket()
{
    return cudaFuncSetAttribute(k_fill_q16<NS, NT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FillQSmem<NS>::bytes);
}
cudaError_t fill_narrow_configure_device()
{
    cudaError_t e;
    if ((e = configure_q_bucket<608, 512, 2>()) != cudaSuccess) return e;
    if ((e = configure_q_bucket<352, 384, 3>()) != cudaSuccess) return e;
    return configure_q_bucket<160, 256, 4>();
}

// bucket: 0 = stride 608, 1 = stride 352, 2 = stride 160.  `a.loci` / `a.flags` already point at the bucket.
cudaError_t launch_fill_narrow(const FillLaunch &a, int bucket, cudaStream_t st)
{
    if (a.nloci <= 0) return cudaSuccess;
    switch (bucket) {
    case 0: k_fill_q16<608, 512, 2><<<a.nloci, 512, FillQSmem<608>::bytes, st>>>(a); break;
    case 1: k_fill_q16<352, 384, 3><<<a.nloci, 384, FillQSmem<352>::bytes, st>>>(a); break;
    default: k_fill_q16<160, 256, 4><<<a.nloci, 256, FillQSmem<160>::bytes, st>>>(a); break;
    }
    return cudaGetLastError();
}
