// peak.cu -- integer-issue roofline microbenchmark: a dependent-free stream of min-plus
// terms per thread, measured on the device the context runs on.  Two variants: plain add + min (the
// compiler's choice of IADD3/VIMNMX or VIADDMNMX) and the DPX intrinsic __viaddmin_s32.
#include "mirfold_internal.cuh"

template <int MODE>
__global__ void __launch_bounds__(256) k_int_peak(int *out, int iters, int c)
{
    int acc[16];
#pragma unroll
    for (int k = 0; k < 16; k++) acc[k] = (int)threadIdx.x * (k + 3) - c * k;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            if (MODE == 0) {  // keep add and min as two instructions (IADD3 + VIMNMX)
                int t;
                asm volatile("add.s32 %0, %1, %2;" : "=r"(t) : "r"(acc[(k + 1) & 15]), "r"(c));
                asm volatile("min.s32 %0, %1, %2;" : "=r"(acc[k]) : "r"(t), "r"(acc[k]));
            }
            else if (MODE == 1) acc[k] = __viaddmin_s32(acc[(k + 1) & 15], c, acc[k]);
            else acc[k] = (int)__viaddmin_s16x2((unsigned)acc[(k + 1) & 15], (unsigned)c, (unsigned)acc[k]);   // VIADDMNMX.S16x2: two terms
        }
    }
    int r = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) r ^= acc[k];
    if (r == 0x7fffffff) out[0] = r;  // never true in practice; keeps the loop alive
}

// returns terms/s for the variants (best of `reps`); s16x2 (may be NULL) counts two terms per instruction
cudaError_t run_int_peak(cudaStream_t st, int sm_count, double *addmin, double *dpx, double *s16x2)
{
    int *d = nullptr;
    cudaError_t e = cudaMalloc(&d, 4);
    if (e != cudaSuccess) return e;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const int blocks = sm_count * 8, iters = 4096;
    double best[3] = {0, 0, 0};
    for (int mode = 0; mode < (s16x2 ? 3 : 2); mode++)
        for (int rep = 0; rep < 5; rep++) {
            cudaEventRecord(a, st);
            if (mode == 0) k_int_peak<0><<<blocks, 256, 0, st>>>(d, iters, 3 + rep);
            else if (mode == 1) k_int_peak<1><<<blocks, 256, 0, st>>>(d, iters, 3 + rep);
            else k_int_peak<2><<<blocks, 256, 0, st>>>(d, iters, 3 + rep);
            cudaEventRecord(b, st);
            e = cudaEventSynchronize(b);
            if (e != cudaSuccess) break;
            float ms = 0;
            cudaEventElapsedTime(&ms, a, b);
            const double terms = (double)blocks * 256.0 * 16.0 * iters * (mode == 2 ? 2.0 : 1.0);
            if (rep > 0) best[mode] = fmax(best[mode], terms / (ms * 1e-3));
        }
    cudaEventDestroy(a); cudaEventDestroy(b);
    cudaFree(d);
    *addmin = best[0];
    *dpx = best[1];
    if (s16x2) *s16x2 = best[2];
    return e;
}
