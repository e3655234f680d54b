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
#include "duplex_dev.cuh"

__global__ void __launch_bounds__(128) k_duplex(const char *__restrict__ arena, const mirfold_duplex_query *__restrict__ qs,
                                                unsigned long long nq, mirfold_duplex_verdict *__restrict__ out, int maxlen)
{
    extern __shared__ short scratch[];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned long long q = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + wib;
    if (q >= nq || lane != 0) return;
    const mirfold_duplex_query Q = qs[q];
    mirfold_duplex_verdict V;
    dev_duplex_eval(arena + Q.ss_off, Q.ss_len, Q.fold_start, Q.mature_start, Q.mature_end, Q.region_start, Q.region_end, Q.strand,
                    scratch + (size_t)wib * 4 * maxlen, maxlen, V);
    out[q] = V;
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
