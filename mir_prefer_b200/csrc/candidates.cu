// candidates.cu -- stages 1 and 3 of the hot path fused on the device after traceback (K6-K8).
//
// Replaces, for every hit the traceback kernels just packed (still in HBM):
//   stage 1: the structure classification of get_structures_next_extendregion (/root/reference/miR_PREFeR.py:1566-1589)
//            with is_stem_loop (:1602), filter_ss (:1685) and has_one_good_bifurcation (:1611);
//   stage 3: get_maturestar_info (:1876-1999) for every (candidate structure x size-admissible mature of the record),
//            the superset of pairs check_loci (:2246-2262) can ask for.
// Only the candidate structures (a few percent of the hit text) and the verdict table cross PCIe.
//
// k_cand_classify : one thread per hit; the reference's stack walks become depth-counter scans (an outermost stem is a
//                   0 -> 1 -> 0 excursion of the nesting depth; the partner of a bracket is found by a directed scan),
//                   so a thread needs no scratch memory.  Run twice: count, then (after a device scan) emit.
// k_cand_duplex   : one warp per (structure, mature) pair, shared-memory scratch (duplex_dev.cuh).
// k_cand_strings  : one warp per structure, copies its dot-bracket string into the compact arena.
#include "../../include/mirfold.h"
#include "mirfold_internal.cuh"
#include "duplex_dev.cuh"

namespace {

__device__ __forceinline__ bool d_stem_loop(const char *__restrict__ ss, int n, int minloop)
{   // ss.find(")") - ss.rfind("(") - 1 >= minloop, with -1 for "not found"
    int first_close = -1, last_open = -1;
    for (int k = 0; k < n; k++) if (ss[k] == ')') { first_close = k; break; }
    for (int k = n - 1; k >= 0; k--) if (ss[k] == '(') { last_open = k; break; }
    return first_close - last_open - 1 >= minloop;
}

// partner of the bracket at pos (-1: a dot or unmatched)
__device__ __forceinline__ int d_partner(const char *__restrict__ ss, int n, int pos)
{
    const char ch = ss[pos];
    int depth = 0;
    if (ch == '(') {
        for (int k = pos; k < n; k++) {
            depth += (ss[k] == '(') - (ss[k] == ')');
            if (depth == 0) return k;
        }
    } else if (ch == ')') {
        for (int k = pos; k >= 0; k--) {
            depth += (ss[k] == ')') - (ss[k] == '(');
            if (depth == 0) return k;
        }
    }
    return -1;
}

// has_one_good_bifurcation (MP:1611-1659): 1 true, 0 false, -1 where the reference raises
__device__ int d_one_good_bifurcation(const char *__restrict__ ss, int n)
{
    bool last_pop = false;
    int depth = 0, last_pos = 0, n_bif = 0, left_close = 0, right_open = 0;
    for (int k = 0; k < n; k++) {
        const char ch = ss[k];
        if (ch == '(') {
            if (k != 0 && last_pop) {
                if (depth == 0) return 0;          // ()() at the top level
                if (n_bif >= 1) return 0;
                n_bif = 1; left_close = last_pos; right_open = k;
            }
            depth++;
            last_pop = false; last_pos = k;
        } else if (ch == ')') {
            if (depth == 0) return -1;
            depth--;
            last_pop = true; last_pos = k;
        }
    }
    const int pr = d_partner(ss, n, right_open), pl = d_partner(ss, n, left_close);
    if (pr < 0 || pl < 0) return -1;               // dict_pos[...] -> KeyError
    if ((double)(pr - pl) / n < 0.5)
        if ((double)pl / n > 0.25)
            if ((double)pr / n < 0.75) return 1;
    return 0;
}

// The candidate structures of one hit, in the reference's order: emit(offset, length, sstype).  Returns their number
// or -1 for input on which the reference's functions raise.
template <class Emit>
__device__ int d_classify_hit(const char *__restrict__ ss, int n, int minloop, Emit emit)
{
    if (d_stem_loop(ss, n, minloop)) { emit(0, n, 0); return 1; }
    int count = 0, depth = 0, piece_begin = 0, cur_end = -1;
    bool have_cur = false, bad = false;
    auto piece = [&](int b, int e) {
        const int len = e - b;
        if (len <= 55) return;
        if (d_stem_loop(ss + b, len, minloop)) { emit(b, len, 0); count++; return; }
        const int g = d_one_good_bifurcation(ss + b, len);
        if (g < 0) bad = true;
        else if (g) { emit(b, len, 1); count++; }
    };
    for (int k = 0; k < n; k++) {
        const char ch = ss[k];
        if (ch == '(') {
            if (depth == 0) {                      // an outermost stem starts here: the previous one's piece ends
                if (have_cur) { piece(piece_begin, k); piece_begin = cur_end + 1; }
                have_cur = true;
            }
            depth++;
        } else if (ch == ')') {
            if (depth == 0) return -1;             // the reference pops an empty list
            if (--depth == 0) cur_end = k;
        }
    }
    if (!have_cur || depth != 0) return -1;        // no pair at all / unbalanced: dict_pair[...] -> KeyError
    piece(piece_begin, n);
    return bad ? -1 : count;
}

// hit index -> locus (largest l with bounds[l] <= h)
__device__ __forceinline__ int d_locus_of_hit(const unsigned long long *__restrict__ bounds, int nl, unsigned long long h)
{
    int lo = 0, hi = nl;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (bounds[mid] <= h) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ int d_admissible(const mirfold_mature *__restrict__ m, unsigned long long b, unsigned long long e, int lo, int hi)
{
    int c = 0;
    for (unsigned long long k = b; k < e; k++) { const int len = m[k].end - m[k].start; c += (len >= lo && len <= hi); }
    return c;
}

}  // namespace

// pass 0: counts[h] = candidate structures of hit h;  pass 1: write them at soff[h]..
template <int PASS>
__global__ void __launch_bounds__(128) k_cand_classify(CandLaunch a)
{
    const unsigned long long h = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (h > a.nhits) return;
    if (h == a.nhits) { if (PASS == 0) a.counts[h] = 0; return; }
    const mirfold_hit hit = a.hits[h];
    if (hit.len < a.minlen) { if (PASS == 0) a.counts[h] = 0; return; }
    const char *ss = a.arena + (hit.ss_off - a.arena_base);
    if (PASS == 0) {
        const int c = d_classify_hit(ss, hit.len, a.minloop, [](int, int, int) {});
        if (c < 0) { atomicExch(a.fail_flag, 2); a.counts[h] = 0; }
        else a.counts[h] = (unsigned long long)c;
    } else {
        const int l = d_locus_of_hit(a.bounds, a.nloci, h);
        const int rec = a.loci[l].rec;
        const double ne = ((double)hit.mfe_dcal / 100.) / (double)hit.len;   // float("%.2f" % E) / len(ss): the same double
        unsigned long long s = a.soff[h];
        const unsigned long long mb = a.mature_off[rec], me = a.mature_off[rec + 1];
        const unsigned long long nadm = (unsigned long long)d_admissible(a.matures, mb, me, a.min_mature, a.max_mature);
        d_classify_hit(ss, hit.len, a.minloop, [&](int off, int len, int type) {
            mirfold_structure st;
            st.rec = (uint32_t)rec; st.fold_start = hit.start + off; st.sstype = type; st.len = len;
            st.ss_off = (hit.ss_off - a.arena_base) + (uint64_t)off;       // chunk arena; rebased by k_cand_strings
            st.norm_energy = ne;
            a.structs[s] = st;
            a.nq[s] = nadm;
            a.sbytes[s] = (unsigned long long)len + 1ULL;
            s++;
        });
    }
}

// per locus: first structure index = soff[bounds[l]]
__global__ void k_cand_locus_bounds(CandLaunch a)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k <= a.nloci) a.locus_sbegin[k] = a.soff[a.bounds[k]];
}

// one warp per structure: every size-admissible mature of its record, in input order
__global__ void __launch_bounds__(128) k_cand_duplex(CandLaunch a, int maxlen)
{
    extern __shared__ short scratch[];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned long long s = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + wib;
    if (s >= a.nstructs || lane != 0) return;
    const mirfold_structure st = a.structs[s];
    const char *ss = a.arena + st.ss_off;
    const mirfold_region rg = a.regions[st.rec];
    unsigned long long v = a.voff[s];
    for (unsigned long long k = a.mature_off[st.rec]; k < a.mature_off[st.rec + 1]; k++) {
        const mirfold_mature m = a.matures[k];
        const int len = m.end - m.start;
        if (len < a.min_mature || len > a.max_mature) continue;
        mirfold_duplex_verdict V;
        dev_duplex_eval(ss, st.len, st.fold_start, m.start, m.end, rg.start, rg.end, m.strand, scratch + (size_t)wib * 4 * maxlen, maxlen, V);
        a.verdicts[v] = V;
        a.verdict_mature[v] = (uint32_t)(k - a.mature_off[st.rec]);
        v++;
    }
}

// one warp per structure: copy its string (NUL-terminated) into the compact arena and rebase ss_off
__global__ void __launch_bounds__(128) k_cand_strings(CandLaunch a)
{
    const unsigned long long s = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (s >= a.nstructs) return;
    const unsigned long long src = a.structs[s].ss_off;
    const int len = a.structs[s].len;
    char *dst = a.out_arena + a.aoff[s];
    for (int k = lane; k <= len; k += 32) dst[k] = k < len ? a.arena[src + k] : 0;
    __syncwarp();
    if (lane == 0) a.structs_out[s] = a.structs[s], a.structs_out[s].ss_off = a.out_base + a.aoff[s];
}

cudaError_t launch_cand_classify(const CandLaunch &a, int pass, cudaStream_t st)
{
    const unsigned blocks = (unsigned)((a.nhits + 1 + 127) / 128);
    if (pass == 0) k_cand_classify<0><<<blocks, 128, 0, st>>>(a);
    else {
        k_cand_classify<1><<<blocks, 128, 0, st>>>(a);
        k_cand_locus_bounds<<<(a.nloci + 1 + 255) / 256, 256, 0, st>>>(a);
    }
    return cudaGetLastError();
}

cudaError_t launch_cand_finish(const CandLaunch &a, int max_len, cudaStream_t st)
{
    if (a.nstructs == 0) return cudaSuccess;
    const int maxlen = (max_len + 7) & ~7;
    const size_t per_warp = (size_t)4 * maxlen * sizeof(short);
    int wpb = (int)(40960 / (per_warp ? per_warp : 1));
    wpb = wpb < 1 ? 1 : (wpb > 4 ? 4 : wpb);
    const size_t smem = per_warp * wpb;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_cand_duplex, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    k_cand_duplex<<<(unsigned)((a.nstructs + wpb - 1) / wpb), wpb * 32, smem, st>>>(a, maxlen);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k_cand_strings<<<(unsigned)((a.nstructs + 3) / 4), 128, 0, st>>>(a);
    return cudaGetLastError();
}
