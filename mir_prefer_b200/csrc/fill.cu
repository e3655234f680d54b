// fill.cu -- sequence encoding (K1), c/fML band fill (K2) and the f3 scan (K3) for sm_100a.
//
// Replaces the arithmetic of RNALfold's fill_arrays (behavioural spec
// Appendix A.1-A.3, "RLF Lfold.c:170-398").  One CTA owns one locus and walks the band by
// anti-diagonal d = j-i (all cells of a diagonal are independent).  Restructuring relative to the
// reference's row-by-row loop (results are bit-identical, the order of evaluation is not):
//   * DML(i,j) = min_k fML(i,k)+fML(k+1,j) only reads diagonals <= d-5, so it is produced in
//     strips of five diagonals ahead of the wavefront (phase A), one thread per row with the
//     five running minima in registers: one fML(i,k) load feeds five min-plus terms.
//   * the <=30-nt interior-loop search is a min-plus stencil over the previous 31 diagonals:
//     generic loops use Cm(p,q) = c(p,q)+mismatchI[rtype][..] (+INF if (p,q) cannot pair) and a
//     per-(u,v) constant, so no pair-type test is needed; a warp evaluates the 496 (u,v) terms
//     of one typed cell in 16 fully-populated iterations (diagonals s and 30-s share a warp)
//     and finishes with a warp min-reduction (redux.sync).
#include "mirfold_internal.cuh"

// ------------------------------------------------------------------------------------ K1
__global__ void k_prepare(const char *__restrict__ raw, const LocusDesc *__restrict__ loci, int nloci,
                          unsigned char *__restrict__ codes, int *__restrict__ F)
{
    // one CTA per locus; RNALfold main(): toupper, T->U; encode_char + alias (A.1)
    const LocusDesc L = loci[blockIdx.x];
    const char *src = raw + L.raw_off;
    unsigned char *dst = codes + L.seq_off;
    int *f = F + L.seq_off;
    for (int k = threadIdx.x; k < L.n + 3; k += blockDim.x) {
        int s = 0;
        if (k >= 1 && k <= L.n) {
            char ch = src[k - 1];
            if (ch >= 'a' && ch <= 'z') ch -= 32;
            switch (ch) {
            case 'A': s = 1; break;
            case 'C': s = 2; break;
            case 'G': s = 3; break;
            case 'U': case 'T': s = 4; break;
            case 'X': s = 5; break;
            case 'K': s = 6; break;
            case 'I': s = 7; break;
            default: s = 0;
            }
        }
        const int alias = (0x02343210 >> (4 * s)) & 7;  // {0,1,2,3,4,3,2,0}
        dst[k] = (unsigned char)(s | (alias << 4));
        f[k] = 0;
    }
}

cudaError_t launch_prepare(const char *raw, const LocusDesc *loci, int nloci, unsigned long long, unsigned char *codes,
                           int *F, cudaStream_t st)
{
    if (nloci == 0) return cudaSuccess;
    k_prepare<<<nloci, 128, 0, st>>>(raw, loci, nloci, codes, F);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------ helpers
__device__ __forceinline__ int warp_min(int v) { return __reduce_min_sync(0xffffffffu, v); }

__device__ __forceinline__ int dev_loop_energy(const DevParams *__restrict__ P, int t, int t2, int n1, int n2, int si1,
                                               int sj1, int sp1, int sq1)
{   // A.2 two-loop energy; t2 already rtype'd
    const int nl = max(n1, n2), ns = min(n1, n2);
    if (nl == 0) return P->stack[t * 8 + t2];
    if (ns == 0) {
        int e = P->bulge[nl];
        if (nl == 1) return e + P->stack[t * 8 + t2];
        return e + (t > 2 ? P->TerminalAU : 0) + (t2 > 2 ? P->TerminalAU : 0);
    }
    if (ns == 1 && nl == 1) return P->int11[((t * 8 + t2) * 5 + si1) * 5 + sj1];
    if (ns == 1 && nl == 2) {
        if (n1 == 1) return P->int21[(((t * 8 + t2) * 5 + si1) * 5 + sq1) * 5 + sj1];
        return P->int21[(((t2 * 8 + t) * 5 + sq1) * 5 + si1) * 5 + sp1];
    }
    if (n1 == 2 && n2 == 2) return P->int22[((((t * 8 + t2) * 5 + si1) * 5 + sp1) * 5 + sq1) * 5 + sj1];
    return P->internal_loop[n1 + n2] + min(300, (nl - ns) * 50) + P->mismatchI[(t * 5 + si1) * 5 + sj1] +
           P->mismatchI[(t2 * 5 + sq1) * 5 + sp1];
}

// hairpin energy of the pair (i,j) of type t; sS/sS1 are 1-based code arrays
__device__ __forceinline__ int dev_hairpin(const DevParams *__restrict__ P, const unsigned char *sS,
                                           const unsigned char *sS1, int i, int j, int t)
{
    const int s = j - i - 1;
    int e = P->hairpinE[s];
    if (s == 4) {
        int code = 0, ok = 1;
#pragma unroll
        for (int k = 0; k < 6; k++) {
            const int b = sS[i + k];
            ok &= (b >= 1 && b <= 4);
            code |= ((b - 1) & 3) << (2 * k);
        }
        if (ok) e += P->tetra[code];
    }
    if (s == 3) e += (t > 2 ? P->TerminalAU : 0);
    else e += P->mismatchH[(t * 5 + sS1[i + 1]) * 5 + sS1[j - 1]];
    return e;
}

// ------------------------------------------------------------------------------------ K2 (v1)
// Dynamic smem: sS[n+3] | sS1[n+3] | pairtab[64] | list[n] (int)
template <int NT>
__global__ void __launch_bounds__(NT) k_fill(FillLaunch a)
{
    extern __shared__ unsigned char smem_raw[];
    const LocusDesc L = a.loci[blockIdx.x];
    const int n = L.n, Ls = L.Ls, dmax = L.dmax;
    const DevParams *__restrict__ P = a.P;
    const int npad = (n + 3 + 15) & ~15;
    unsigned char *sS = smem_raw;
    unsigned char *sS1 = sS + npad;
    unsigned char *sPair = sS1 + npad;
    int *sList = (int *)(sPair + 64);
    __shared__ int sCount;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NW = NT / 32;

    for (int k = tid; k < n + 3; k += NT) {
        const unsigned char b = a.codes[L.seq_off + k];
        sS[k] = b & 7;
        sS1[k] = b >> 4;
    }
    if (tid < 64) sPair[tid] = P->pair[tid];
    __syncthreads();

    int *__restrict__ C = a.C + L.band_off;
    int *__restrict__ M = a.M + L.band_off;
    int *__restrict__ rCm = a.ring + L.ring_off;                       // [MF_RING_CM][n]
    int *__restrict__ rD = rCm + (unsigned long long)MF_RING_CM * n;   // [MF_RING_DML][n]

    for (int d = 4; d <= dmax; d++) {
        // ---------------- phase A: DML for the strip d..d+4 (needs fML on diagonals <= d-1 only)
        if ((d - 4) % 5 == 0) {
            const int d1 = min(d + 4, dmax);
            const int R = n - d;                       // rows of the first strip diagonal
            const int Rpad = (R + 31) & ~31;
            int nparts = NT / Rpad;
            nparts = max(1, min(nparts, 8));
            // reset ring slots of the strip
            for (int s = 0; s <= d1 - d; s++)
                for (int i = tid; i < n; i += NT) rD[((d + s) & (MF_RING_DML - 1)) * n + i] = MF_INF;
            __syncthreads();
            const int emax = d1 - 5;                   // e ranges 4..emax (for the widest diagonal)
            if (emax >= 4) {
                const int span = emax - 4 + 1;
                const int per = (span + nparts - 1) / nparts;
                for (int base = 0; base < Rpad * nparts; base += NT) {
                    const int idx = base + tid;
                    const int part = idx / Rpad, i = idx - part * Rpad + 1;
                    if (part < nparts && i <= R) {
                        const int e0 = 4 + part * per, e1 = min(emax, e0 + per - 1);
                        int acc[5] = {MF_INF, MF_INF, MF_INF, MF_INF, MF_INF};
                        for (int e = e0; e <= e1; e++) {
                            const int av = M[band_doff(n, e) + (i - 1)];   // fML(i, i+e)
#pragma unroll
                            for (int s = 0; s < 5; s++) {
                                const int dd = d + s;              // target diagonal
                                if (dd <= d1 && e <= dd - 5 && i <= n - dd) {
                                    const int bv = M[band_doff(n, dd - 1 - e) + (i + e)];  // fML(i+e+1, i+dd)
                                    acc[s] = min(acc[s], av + bv);
                                }
                            }
                        }
#pragma unroll
                        for (int s = 0; s < 5; s++) {
                            const int dd = d + s;
                            if (dd <= d1 && i <= n - dd && acc[s] < MF_INF) {
                                int *dst = &rD[(dd & (MF_RING_DML - 1)) * n + (i - 1)];
                                if (nparts == 1) *dst = acc[s];
                                else atomicMin(dst, acc[s]);
                            }
                        }
                    }
                }
            }
            __syncthreads();
        }

        // ---------------- phase C: c on diagonal d
        const int ncell = n - d;
        int *__restrict__ Cd = C + band_doff(n, d);
        int *__restrict__ CmD = rCm + (d & (MF_RING_CM - 1)) * n;
        if (tid == 0) sCount = 0;
        __syncthreads();
        for (int i = tid + 1; i <= ncell; i += NT) {
            const int t = (d < Ls) ? sPair[sS[i] * 8 + sS[i + d]] : 0;
            if (t) sList[atomicAdd(&sCount, 1)] = i;
            else { Cd[i - 1] = MF_INF; CmD[i - 1] = MF_INF; }
        }
        __syncthreads();
        const int ntyped = sCount;
        const int K = min(30, d - 6);   // largest u+v
        for (int c = wid; c < ntyped; c += NW) {
            const int i = sList[c], j = i + d;
            const int t = sPair[sS[i] * 8 + sS[j]];
            const int si1 = sS1[i + 1], sj1 = sS1[j - 1];
            int best = MF_INF;
            // generic interior loops: 496 (u,v) terms, diagonals s and 30-s share an iteration
#pragma unroll 4
            for (int it = 0; it < 16; it++) {
                int s, u;
                if (it < 15) { if (lane <= it) { s = it; u = lane; } else { s = 30 - it; u = lane - it - 1; } }
                else { s = 15; u = lane; }
                const int cst = P->ilc[it][lane];
                if (s <= K && cst < MF_INF) {
                    const int v = rCm[((d - 2 - s) & (MF_RING_CM - 1)) * n + (i + u)];  // Cm(i+1+u, .)
                    best = min(best, v + cst);
                }
            }
            best += P->mismatchI[(t * 5 + si1) * 5 + sj1];
            // special two-loops: stack, bulges, 1x1, 1x2, 2x1, 2x2 (65 terms)
            for (int r = 0; r < 3; r++) {
                const int m = lane + 32 * r;
                int u, v;
                if (m == 0) { u = 0; v = 0; }
                else if (m <= 30) { u = m; v = 0; }
                else if (m <= 60) { u = 0; v = m - 30; }
                else if (m <= 64) { u = 1 + ((m - 61) >> 1); v = 1 + ((m - 61) & 1); }
                else { u = 99; v = 99; }
                if (u + v <= K) {
                    const int p = i + 1 + u, q = j - 1 - v;
                    const int t2 = sPair[sS[p] * 8 + sS[q]];
                    if (t2) {
                        const int e = dev_loop_energy(P, t, P->rtype[t2], u, v, si1, sj1, sS1[p - 1], sS1[q + 1]);
                        best = min(best, e + C[band_doff(n, q - p) + (p - 1)]);
                    }
                }
            }
            best = warp_min(best);
            if (lane == 0) {
                best = min(best, dev_hairpin(P, sS, sS1, i, j, t));
                // multiloop closing with d1 dangles (A.3)
                const int tt = P->rtype[t];
                const int d3 = P->dangle3[tt * 5 + si1], d5 = P->dangle5[tt * 5 + sj1];
                int dec = MF_INF;
                if (d - 2 >= 4) dec = rD[((d - 2) & (MF_RING_DML - 1)) * n + i];                       // DML(i+1,j-1)
                if (d - 3 >= 4) {
                    dec = min(dec, rD[((d - 3) & (MF_RING_DML - 1)) * n + i + 1] + d3);                // DML(i+2,j-1)
                    dec = min(dec, rD[((d - 3) & (MF_RING_DML - 1)) * n + i] + d5);                    // DML(i+1,j-2)
                }
                if (d - 4 >= 4) dec = min(dec, rD[((d - 4) & (MF_RING_DML - 1)) * n + i + 1] + d3 + d5);  // DML(i+2,j-2)
                best = min(best, P->MLclosing + P->MLintern[t] + dec);
                Cd[i - 1] = best;
                CmD[i - 1] = best + P->mismatchI[(tt * 5 + sS1[j + 1]) * 5 + sS1[i - 1]];
            }
        }
        __syncthreads();

        // ---------------- phase M: fML on diagonal d
        int *__restrict__ Md = M + band_doff(n, d);
        const int *__restrict__ Mp = (d - 1 >= 4) ? M + band_doff(n, d - 1) : nullptr;
        const int *__restrict__ Cp = (d - 1 >= 4) ? C + band_doff(n, d - 1) : nullptr;
        const int *__restrict__ Cpp = (d - 2 >= 4) ? C + band_doff(n, d - 2) : nullptr;
        const int *__restrict__ Dd = rD + (d & (MF_RING_DML - 1)) * n;
        for (int i = tid + 1; i <= ncell; i += NT) {
            const int j = i + d;
            const int t = (d < Ls) ? sPair[sS[i] * 8 + sS[j]] : 0;
            int m = MF_INF;
            if (Mp) m = min(Mp[i], Mp[i - 1]);                       // fML(i+1,j), fML(i,j-1)
            m = min(m, Cd[i - 1] + P->MLintern[t]);
            if (Cp) {
                const int ta = sPair[sS[i + 1] * 8 + sS[j]];         // (i+1, j)
                m = min(m, Cp[i] + P->dangle5[ta * 5 + sS1[i]] + P->MLintern[ta]);
                const int tb = sPair[sS[i] * 8 + sS[j - 1]];         // (i, j-1)
                m = min(m, Cp[i - 1] + P->dangle3[tb * 5 + sS1[j]] + P->MLintern[tb]);
            }
            if (Cpp) {
                const int tc = sPair[sS[i + 1] * 8 + sS[j - 1]];     // (i+1, j-1)
                m = min(m, Cpp[i] + P->dangle5[tc * 5 + sS1[i]] + P->dangle3[tc * 5 + sS1[j]] + P->MLintern[tc]);
            }
            m = min(m, Dd[i - 1]);
            Md[i - 1] = m;
        }
        __syncthreads();
    }
}

cudaError_t launch_fill(const FillLaunch &a, cudaStream_t st)
{
    if (a.nloci == 0) return cudaSuccess;
    constexpr int NT = 512;
    const int npad = (a.max_n + 3 + 15) & ~15;
    const size_t smem = (size_t)2 * npad + 64 + (size_t)a.max_n * sizeof(int) + 16;
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(k_fill<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    k_fill<NT><<<a.nloci, NT, smem, st>>>(a);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------ K3
// f3 scan (A.3, "RLF Lfold.c:355-398"): one warp per locus, lanes over j, rows sequential.
__global__ void __launch_bounds__(128) k_f3(const LocusDesc *__restrict__ loci, int nloci,
                                            const unsigned char *__restrict__ codes, const int *__restrict__ Call,
                                            int *__restrict__ Fall, const DevParams *__restrict__ P)
{
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= nloci) return;
    const LocusDesc L = loci[w];
    const int n = L.n, Ls = L.Ls;
    const unsigned char *__restrict__ cd = codes + L.seq_off;
    const int *__restrict__ C = Call + L.band_off;
    int *F = Fall + L.seq_off;
    const int AUp = P->TerminalAU;
    for (int i = n - 4; i >= 1; i--) {
        int best = MF_INF;
        const int si = cd[i] & 7, s1i = cd[i] >> 4, si1 = cd[i + 1] & 7;
        const int jend = min(n - 1, i + Ls);
        for (int j = i + 4 + lane; j <= jend; j += 32) {
            const int d = j - i;
            const int cj = cd[j], cj1 = cd[j + 1];
            const int sj = cj & 7, s1j1 = cj1 >> 4;
            const int f1 = F[j + 1], f2 = F[j + 2];
            int t = (d < Ls) ? P->pair[si * 8 + sj] : 0;
            if (t) {
                const int e = C[band_doff(n, d) + (i - 1)] + (t > 2 ? AUp : 0);
                best = min(best, min(e + f1, e + P->dangle3[t * 5 + s1j1] + f2));
            }
            t = (d - 1 >= 4) ? P->pair[si1 * 8 + sj] : 0;
            if (t) {
                const int e = C[band_doff(n, d - 1) + i] + P->dangle5[t * 5 + s1i] + (t > 2 ? AUp : 0);
                best = min(best, min(e + f1, e + P->dangle3[t * 5 + s1j1] + f2));
            }
        }
        if (lane == 0 && n <= i + Ls) {
            const int d = n - i, sj = cd[n] & 7;
            int t = (d < Ls) ? P->pair[si * 8 + sj] : 0;
            if (t) best = min(best, C[band_doff(n, d) + (i - 1)] + (t > 2 ? AUp : 0));
            t = (d - 1 >= 4) ? P->pair[si1 * 8 + sj] : 0;
            if (t) best = min(best, C[band_doff(n, d - 1) + i] + P->dangle5[t * 5 + s1i] + (t > 2 ? AUp : 0));
        }
        best = warp_min(best);
        if (lane == 0) F[i] = min(F[i + 1], best);
        __syncwarp();
    }
}

cudaError_t launch_f3(const LocusDesc *loci, int nloci, const unsigned char *codes, const int *C, int *F,
                      const DevParams *P, cudaStream_t st)
{
    if (nloci == 0) return cudaSuccess;
    const int warps_per_block = 4;
    k_f3<<<(nloci + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, st>>>(loci, nloci, codes, C, F, P);
    return cudaGetLastError();
}
