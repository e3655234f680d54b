// fill_common.cuh -- device helpers shared by the band-fill kernels (fill.cu, fill_narrow.cu).
// Loop energies follow SURVEY.md Appendix A.2 (RLF fold.c HairpinE @0x40d7a0, LoopEnergy @0x40c730).
#pragma once
#include "mirfold_internal.cuh"

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

__device__ __forceinline__ void dev_prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// shared-memory load from an absolute 32-bit shared address (volatile: never hoisted or merged)
__device__ __forceinline__ unsigned dev_lds(unsigned addr)
{
    unsigned w;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(addr));
    return w;
}

// two adjacent 16-bit shared-memory stores at an absolute shared address
__device__ __forceinline__ void dev_sts16x2(unsigned addr, int a, int b)
{
    asm volatile("st.shared.u16 [%0], %1;\n\tst.shared.u16 [%0+2], %2;" ::"r"(addr), "h"((unsigned short)a), "h"((unsigned short)b) : "memory");
}
__device__ __forceinline__ int dev_rtype(int t) { return t ? (((t - 1) ^ 1) + 1) : 0; }   // {0,2,1,4,3,6,5}

// ---- bulk asynchronous copies (TMA engine, 1-D): a finished band row leaves the SM as ONE cp.async.bulk from the
// shared-memory row instead of one STG per cell (SASS: UBLKCP).  `bytes` is a multiple of 16, both addresses are
// 16-byte aligned.  The issuing thread fences the generic-proxy writes of the row (ordered before by a barrier)
// into the async proxy first; dev_bulk_wait_read() returns once the engine no longer reads shared memory.
__device__ __forceinline__ void dev_bulk_store_row(void *gdst, const void *ssrc, unsigned bytes)
{
    const unsigned saddr = (unsigned)__cvta_generic_to_shared(ssrc);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(saddr), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void dev_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void dev_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
