// host.cu -- libmirfold host runtime: contexts, memory pools, chunked pipeline, multi-GPU sharding
// and the C ABI declared in include/mirfold.h.
//
// Replaces fold_use_RNALfold()/fold() (/root/reference/miR_PREFeR.py:3047-3119): where the reference
// forks one RNALfold process per FASTA shard, this runtime shards records over GPUs by DP work
// (no collectives: loci are independent), runs K1..K4 per chunk on one stream per device, and
// gathers hit records back in input order.  There is no CPU fallback.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cub/device/device_scan.cuh>

#include "../../include/mirfold.h"
#include "mirfold_internal.cuh"
#include "turner99_v185_tables.inc"

#define MIRFOLD_VERSION "0.1.0"

namespace {

struct DBuf {  // grow-only device buffer
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { e = cudaMalloc(&p, bytes); want = bytes; }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return (T *)p; }
};

struct HBuf {  // grow-only pinned host buffer
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return (T *)p; }
};

struct Partial {  // what one device produced for its share of the records
    std::vector<uint32_t> recs;            // record ids handled (n >= 5 only)
    std::vector<uint64_t> rec_hit_begin;   // per handled record: first hit in `hits`
    std::vector<uint32_t> rec_hit_count;
    std::vector<int32_t> rec_total;
    HBuf hits;                             // pinned mirfold_hit[nhits], locus order; ss_off relative to arena
    uint64_t nhits = 0;
    HBuf arena;                            // pinned
    uint64_t arena_bytes = 0;
    mirfold_stats st{};
    int err = MIRFOLD_OK;
    std::string errmsg;
};

struct Device {
    int id = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t side = nullptr;     // small fill buckets run here, concurrently with the big one
    DevParams *dP = nullptr;
    size_t mem_budget = 0;
    // pooled buffers
    std::vector<LocusDesc> units_host;   // fill units of the current chunk (pageable: uploaded with a synchronous-staging memcpy)
    DBuf units;
    DBuf raw, loci, codes, F, C, M, Mp, ring, fillflags, tbcount, tbbase, listoff, startlist, scan_in, scan_out, scan_tmp;
    DBuf slots, tblen, tbstart, tblocus, tbflag, tbenergy, stackscr, fail, ssoff, hitidx;
    DBuf o_hits, o_arena;
    HBuf h_raw, h_loci, h_listoff, h_small, h_out;
    cudaEvent_t ev[12] = {};
    void release()
    {
        DBuf *all[] = {&units, &raw, &loci, &codes, &F, &C, &M, &Mp, &ring, &fillflags, &tbcount, &tbbase, &listoff, &startlist, &scan_in,
                       &scan_out, &scan_tmp, &slots, &tblen, &tbstart, &tblocus, &tbflag, &tbenergy, &stackscr, &fail,
                       &ssoff, &hitidx, &o_hits, &o_arena};
        for (DBuf *b : all) b->release();
        HBuf *hall[] = {&h_raw, &h_loci, &h_listoff, &h_small, &h_out};
        for (HBuf *b : hall) b->release();
        for (auto &e : ev) if (e) { cudaEventDestroy(e); e = nullptr; }
        if (dP) cudaFree(dP);
        dP = nullptr;
        if (stream) cudaStreamDestroy(stream);
        stream = nullptr;
        if (side) cudaStreamDestroy(side);
        side = nullptr;
    }
};

}  // namespace

struct mirfold_ctx {
    std::vector<Device> devs;
    std::string last_error;
    std::vector<HBuf> arena_pool;  // pinned arenas returned by mirfold_free_result
    std::vector<HBuf> hits_pool;   // pinned hit tables returned by mirfold_free_result
    std::mutex pool_mu;
};

namespace {

struct ResultOwner {  // lives right behind the public struct
    mirfold_result pub;
    mirfold_ctx *ctx;
    std::vector<uint64_t> hit_begin;
    std::vector<uint32_t> hit_count;
    HBuf hits;                 // single-device fast path: pinned hit table moved from the Partial
    mirfold_hit *hits_m = nullptr;     // multi-device path: concatenated (malloc)
    std::vector<int32_t> totals;
    HBuf arena;               // single-device fast path: pinned arena moved from the Partial
    char *arena_m = nullptr;    // multi-device path: concatenated (malloc)
};

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

// host-side phase log (MIRFOLD_HOST_TIMING=1): where the wall time outside the kernels goes
struct HostTimer {
    std::chrono::steady_clock::time_point t;
    bool on;
    HostTimer() : t(std::chrono::steady_clock::now()) { static const bool e = getenv("MIRFOLD_HOST_TIMING") != nullptr; on = e; }
    void mark(const char *what)
    {
        if (!on) return;
        const auto n = std::chrono::steady_clock::now();
        fprintf(stderr, "[mirfold host] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(n - t).count());
        t = n;
    }
};

int env_opts()
{
    static const int v = getenv("MIRFOLD_OPTS") ? atoi(getenv("MIRFOLD_OPTS")) : 0;
    return v;
}

uint64_t cells_of(int n, int L)
{   // sum_{i=1}^{n-4} (min(n, i+L*) - i - 3)
    const int Ls = std::min(L, n);
    uint64_t tot = 0;
    if (n < 5) return 0;
    // rows with i+Ls <= n contribute Ls-3, the rest n-i-3
    const int full = std::max(0, std::min(n - 4, n - Ls));
    tot += (uint64_t)full * (uint64_t)(Ls - 3);
    for (int i = full + 1; i <= n - 4; i++) tot += (uint64_t)(n - i - 3);
    return tot;
}

#define CK(call)                                                                                 \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess) {                                                                 \
            out.err = MIRFOLD_ERR_CUDA;                                                          \
            out.errmsg = std::string(#call) + ": " + cudaGetErrorString(e_);                     \
            return;                                                                              \
        }                                                                                        \
    } while (0)

__global__ void k_widen_counts(const int *__restrict__ in, unsigned long long *__restrict__ out, int n)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k <= n) out[k] = k < n ? (unsigned long long)in[k] : 0ULL;
}
__global__ void k_emit_sizes(const int *__restrict__ flag, const int *__restrict__ len,
                             unsigned long long *__restrict__ bytes, unsigned long long *__restrict__ ones,
                             unsigned long long ntb)
{
    const unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k <= ntb) {
        const bool f = k < ntb && flag[k];
        bytes[k] = f ? (unsigned long long)len[k] + 1ULL : 0ULL;
        ones[k] = f ? 1ULL : 0ULL;
    }
}

__global__ void k_gather_bounds(const unsigned long long *tb_base, const unsigned long long *hitidx,
                                unsigned long long *out, int nl)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k <= nl) out[k] = hitidx[tb_base[k]];
}
__global__ void k_gather_totals(const LocusDesc *loci, const int *F, int *out, int nl)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nl) out[k] = F[loci[k].seq_off + 1];
}

cudaError_t exclusive_scan(Device &D, const unsigned long long *in, unsigned long long *out, size_t n)
{
    size_t tmp = 0;
    cudaError_t e = cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, out, n, D.stream);
    if (e != cudaSuccess) return e;
    e = D.scan_tmp.ensure(tmp);
    if (e != cudaSuccess) return e;
    return cub::DeviceScan::ExclusiveSum(D.scan_tmp.p, tmp, in, out, n, D.stream);
}

struct Locus {
    uint32_t rec;
    int n;
    uint64_t cells;
};

// Band shape of one locus: stride bucket or, for n > MF_TILE_LEN with a span that leaves at least
// MF_TILE_MIN_STEP owned rows per tile, overlapping tiles for the shared-memory kernels.
// Returns the band elements (per int32 array) the locus occupies.
#define MF_TILE_MIN_STEP 64
unsigned long long shape_locus(LocusDesc &d, int n, int L)
{
    d.n = n; d.Ls = std::min(L, n); d.dmax = std::min(d.Ls, n - 1);
    d.stride = band_stride_for(n);
    d.tile_last = 0; d.tile_step = 1 << 30; d.tile_rcp = 0;
    static const bool no_tiles = getenv("MIRFOLD_NO_TILES") != nullptr;   // A/B runs: long loci through k_fill_generic
    if (n > MF_TILE_LEN && d.dmax >= 4 && MF_TILE_LEN - d.dmax >= MF_TILE_MIN_STEP && !no_tiles) {
        const int S = MF_TILE_LEN - d.dmax;
        d.stride = MF_TILE_LEN;
        d.tile_step = S;
        d.tile_last = (n - MF_TILE_LEN + S - 1) / S;
        d.tile_rcp = (unsigned int)((0x100000000ULL + (unsigned long long)S - 1) / (unsigned long long)S);
        return (unsigned long long)(d.tile_last + 1) * band_elems(MF_TILE_LEN, d.dmax);
    }
    return band_elems(d.stride, d.dmax);
}
unsigned long long unit_ring_elems(int n, int stride)
{
    return (unsigned long long)stride * (n > MF_TILE_LEN ? MF_RING_PER_STRIDE : MF_RING_DML);
}
// Fill units of a chunk: one per untiled locus, one per tile otherwise, sorted by descending n
// so that the stride buckets are contiguous.  Assigns ring offsets; returns ring elements.
unsigned long long build_fill_units(const LocusDesc *loci, int nl, std::vector<LocusDesc> &units, int bucket_first[5], int &max_n)
{
    units.clear();
    max_n = 0;
    for (int k = 0; k < nl; k++) {
        const LocusDesc &d = loci[k];
        if (d.dmax < 4) continue;   // n < 5 never reaches here; keeps the kernels' assumptions explicit
        if (d.tile_last == 0) { units.push_back(d); continue; }
        for (int t = 0; t <= d.tile_last; t++) {
            LocusDesc u = d;
            const int a = std::min(t * d.tile_step, d.n - MF_TILE_LEN);
            u.n = MF_TILE_LEN; u.Ls = std::min(d.Ls, MF_TILE_LEN); u.dmax = d.dmax;
            u.seq_off = d.seq_off + (unsigned long long)a;
            u.band_off = d.band_off + (unsigned long long)t * band_elems(MF_TILE_LEN, d.dmax);
            u.tile_last = 0;
            units.push_back(u);
        }
    }
    std::stable_sort(units.begin(), units.end(), [](const LocusDesc &a, const LocusDesc &b) { return a.n > b.n; });
    unsigned long long ring_acc = 0;
    int k = 0;
    const int nu = (int)units.size();
    const int lim[3] = {608, 352, 160};
    bucket_first[0] = 0;
    for (int b = 0; b < 3; b++) {
        while (k < nu && units[k].n > lim[b]) k++;
        bucket_first[b + 1] = k;
    }
    bucket_first[4] = nu;
    for (auto &u : units) {
        u.ring_off = ring_acc;
        ring_acc += unit_ring_elems(u.n, u.stride);
        max_n = std::max(max_n, u.n);
    }
    return ring_acc;
}

// Runs the full pipeline for `recs` on one device.  If d_raw != nullptr the raw sequences already
// live on the device (offsets h_off are into that buffer) and no results are downloaded.
void run_device(Device &D, const char *seqs, const uint64_t *h_off, const std::vector<uint32_t> &recs, int L,
                const char *d_raw, bool download, cudaStream_t user_stream, bool force_wide, Partial &out)
{
    out.st = mirfold_stats{};
    out.st.n_devices = 1;
    CK(cudaSetDevice(D.id));
    cudaStream_t st = user_stream ? user_stream : D.stream;
    const auto t0 = std::chrono::steady_clock::now();

    std::vector<Locus> loci;
    loci.reserve(recs.size());
    for (uint32_t r : recs) {
        const uint64_t len = h_off[r + 1] - h_off[r];
        if (len >= 5) {
            Locus l{r, (int)len, cells_of((int)len, L)};
            loci.push_back(l);
            out.st.nt += len;
            out.st.cells += l.cells;
        } else out.st.nt += len;
    }
    // largest first: the hardware CTA scheduler then behaves like LPT list scheduling
    std::stable_sort(loci.begin(), loci.end(), [](const Locus &a, const Locus &b) { return a.cells > b.cells; });

    out.recs.clear(); out.rec_hit_begin.clear(); out.rec_hit_count.clear(); out.rec_total.clear(); out.nhits = 0;
    out.arena_bytes = 0;

    HostTimer ht;
    ht.mark("sort loci");
    // ---- chunking by device memory budget
    struct Chunk { size_t begin, end; };
    std::vector<Chunk> chunks;
    auto locus_bytes = [&](const Locus &l) {
        LocusDesc d{};
        const unsigned long long be = shape_locus(d, l.n, L);
        return (size_t)be * 12 + (size_t)(d.tile_last + 1) * unit_ring_elems(std::min(l.n, d.tile_last ? MF_TILE_LEN : l.n), d.stride) * 4 +
               (size_t)(d.tile_last + 2) * sizeof(LocusDesc) + (size_t)l.n * 16 + 4096;
    };
    {
        size_t b = 0, acc = 0;
        for (size_t k = 0; k < loci.size(); k++) {
            const size_t lb = locus_bytes(loci[k]);
            if (k > b && acc + lb > D.mem_budget) { chunks.push_back({b, k}); b = k; acc = 0; }
            acc += lb;
        }
        if (b < loci.size()) chunks.push_back({b, loci.size()});
    }
    out.st.n_chunks = (int32_t)chunks.size();

    struct ChunkOut {  // device-resident results of a chunk, downloaded at the end
        uint64_t nhits, arena_bytes;
    };
    // pass 1 over chunks computes everything and downloads the small per-hit arrays + arena
    // into the pinned arena (grown as needed; chunks are few).
    std::vector<char> arena_tmp;  // only used when more than one chunk
    for (size_t ci = 0; ci < chunks.size(); ci++) {
        const size_t cb = chunks[ci].begin, ce = chunks[ci].end;
        const int nl = (int)(ce - cb);
        // ---- descriptors
        CK(D.h_loci.ensure(sizeof(LocusDesc) * nl));
        CK(D.h_listoff.ensure(sizeof(unsigned long long) * nl));
        LocusDesc *hl = D.h_loci.as<LocusDesc>();
        unsigned long long *hlo = D.h_listoff.as<unsigned long long>();
        unsigned long long seq_acc = 0, band_acc = 0, raw_acc = 0, list_acc = 0;
        int max_n = 0, max_Ls = 0;
        for (int k = 0; k < nl; k++) {
            const Locus &l = loci[cb + k];
            LocusDesc &d = hl[k];
            d = LocusDesc{};
            const unsigned long long be = shape_locus(d, l.n, L);
            d.rec = (int)l.rec;
            d.seq_off = seq_acc; d.band_off = band_acc; d.ring_off = 0;
            d.raw_off = d_raw ? h_off[l.rec] : raw_acc;
            hlo[k] = list_acc;
            seq_acc += (unsigned long long)l.n + 3;
            band_acc += be;
            raw_acc += (unsigned long long)l.n;
            list_acc += (unsigned long long)l.n / 2 + 2;
            max_Ls = std::max(max_Ls, d.Ls);
        }
        std::vector<LocusDesc> &units = D.units_host;
        int bucket_first[5];
        const unsigned long long ring_acc = build_fill_units(hl, nl, units, bucket_first, max_n);
        const int nu = (int)units.size();
        out.st.fill_units += (uint64_t)nu;
        ht.mark("descriptors + fill units");
        // ---- upload
        CK(cudaEventRecord(D.ev[0], st));
        const char *raw_dev = d_raw;
        if (!d_raw) {
            CK(D.h_raw.ensure(raw_acc));
            char *hr = D.h_raw.as<char>();
            for (int k = 0; k < nl; k++) memcpy(hr + hl[k].raw_off, seqs + h_off[loci[cb + k].rec], (size_t)hl[k].n);
            CK(D.raw.ensure(raw_acc));
            CK(cudaMemcpyAsync(D.raw.p, hr, raw_acc, cudaMemcpyHostToDevice, st));
            out.st.h2d_bytes += raw_acc;
            raw_dev = D.raw.as<char>();
        }
        CK(D.loci.ensure(sizeof(LocusDesc) * nl));
        CK(cudaMemcpyAsync(D.loci.p, hl, sizeof(LocusDesc) * nl, cudaMemcpyHostToDevice, st));
        CK(D.listoff.ensure(sizeof(unsigned long long) * nl));
        CK(cudaMemcpyAsync(D.listoff.p, hlo, sizeof(unsigned long long) * nl, cudaMemcpyHostToDevice, st));
        out.st.h2d_bytes += (sizeof(LocusDesc) + 8) * (uint64_t)nl;
        CK(D.codes.ensure(seq_acc));
        CK(D.F.ensure(seq_acc * 4));
        CK(D.C.ensure(band_acc * 4));
        CK(D.M.ensure(band_acc * 4));
        if (!force_wide) CK(D.Mp.ensure(band_acc * 4));
        CK(D.ring.ensure(ring_acc * 4));
        CK(D.tbcount.ensure((size_t)nl * 4));
        CK(D.tbbase.ensure((size_t)(nl + 1) * 8));
        CK(D.scan_in.ensure((size_t)(nl + 1) * 8));
        CK(D.startlist.ensure(list_acc * 4));
        CK(D.fail.ensure(4));
        CK(cudaMemsetAsync(D.fail.p, 0, 4, st));
        CK(D.fillflags.ensure((size_t)nu * 4 + 4));
        CK(cudaMemsetAsync(D.fillflags.p, 0, (size_t)nu * 4 + 4, st));
        CK(D.units.ensure(sizeof(LocusDesc) * (size_t)nu + 64));
        CK(cudaMemcpyAsync(D.units.p, units.data(), sizeof(LocusDesc) * (size_t)nu, cudaMemcpyHostToDevice, st));
        out.st.h2d_bytes += sizeof(LocusDesc) * (uint64_t)nu;
        CK(cudaEventRecord(D.ev[1], st));
        ht.mark("stage raw + enqueue uploads");
        // ---- K1..K3
        const LocusDesc *dl = D.loci.as<LocusDesc>();
        CK(launch_prepare(raw_dev, dl, nl, seq_acc, D.codes.as<unsigned char>(), D.F.as<int>(), st));
        CK(cudaEventRecord(D.ev[2], st));
        FillLaunch fa{D.units.as<LocusDesc>(), nu, max_n, D.codes.as<unsigned char>(), D.C.as<int>(), D.M.as<int>(), D.ring.as<int>(),
                      D.Mp.as<unsigned int>(), D.dP, {0, 0, 0, 0, 0}, D.fillflags.as<int>(), force_wide ? 1 : 0, env_opts()};
        for (int b = 0; b < 5; b++) fa.bucket_first[b] = bucket_first[b];
        static const bool no_side = getenv("MIRFOLD_NO_SIDE_STREAM") != nullptr;   // A/B runs
        CK(launch_fill(fa, st, no_side ? nullptr : D.side, D.ev[10], D.ev[11]));
        CK(cudaEventRecord(D.ev[3], st));
        int n_long = 0;
        while (n_long < nl && hl[n_long].n > MF_TILE_LEN) n_long++;   // sorted by descending cells == descending n
        CK(launch_f3(dl, nl, n_long, max_Ls, D.codes.as<unsigned char>(), D.C.as<int>(), D.F.as<int>(), D.dP, st));
        {   // kernels launched so far: k_prepare, the fill kernels (16-bit + 32-bit redo per non-empty bucket, generic), k_f3 / k_f3_cta
            int nfill = bucket_first[1] > bucket_first[0] ? 1 : 0;
            for (int b = 1; b < 4; b++) if (bucket_first[b + 1] > bucket_first[b]) nfill += force_wide ? 1 : 2;
            out.st.kernel_launches += 1 + nfill + (n_long > 0 ? 1 : 0) + (n_long < nl ? 1 : 0);
        }
        CK(cudaEventRecord(D.ev[4], st));
        // ---- K4: plan
        TraceBuffers tb{};
        tb.loci = dl; tb.nloci = nl; tb.codes = D.codes.as<unsigned char>();
        tb.C = D.C.as<int>(); tb.M = D.M.as<int>(); tb.F = D.F.as<int>(); tb.P = D.dP;
        tb.tb_count = D.tbcount.as<int>(); tb.tb_base = D.tbbase.as<unsigned long long>();
        tb.list_off = D.listoff.as<unsigned long long>(); tb.tb_start_list = D.startlist.as<int>();
        tb.fail_flag = D.fail.as<int>();
        CK(launch_plan(tb, st));
        k_widen_counts<<<(nl + 1 + 255) / 256, 256, 0, st>>>(tb.tb_count, D.scan_in.as<unsigned long long>(), nl);
        CK(cudaGetLastError());
        CK(exclusive_scan(D, D.scan_in.as<unsigned long long>(), tb.tb_base, (size_t)nl + 1));
        CK(D.h_small.ensure(256));
        unsigned long long *hs = D.h_small.as<unsigned long long>();
        CK(cudaMemcpyAsync(&hs[0], tb.tb_base + nl, 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const unsigned long long ntb = hs[0];
        out.st.tracebacks += ntb;
        out.st.kernel_launches += 4;   // k_plan, k_widen_counts, cub scan (init + scan)
        ht.mark("K1-K3 enqueue + plan sync");
        // ---- traceback
        tb.ntb = ntb;
        tb.slot_stride = (max_Ls + 4 + 3) & ~3;
        tb.stack_cap = max_Ls / 4 + 16;
        tb.code_win = (max_Ls + 8 + 15) & ~15;
        CK(D.slots.ensure((size_t)ntb * tb.slot_stride + 16));
        CK(D.tblen.ensure((size_t)ntb * 4 + 16)); CK(D.tbstart.ensure((size_t)ntb * 4 + 16));
        CK(D.tblocus.ensure((size_t)ntb * 4 + 16)); CK(D.tbflag.ensure((size_t)ntb * 4 + 16));
        CK(D.tbenergy.ensure((size_t)ntb * 4 + 16));
        CK(D.stackscr.ensure((size_t)ntb * tb.stack_cap * 8 + 16));
        CK(D.ssoff.ensure((size_t)(ntb + 1) * 8)); CK(D.hitidx.ensure((size_t)(ntb + 1) * 8));
        CK(D.scan_in.ensure((size_t)(ntb + 1) * 8 + (size_t)(nl + 1) * 8));
        CK(D.scan_out.ensure((size_t)(ntb + 1) * 8));
        tb.slots = D.slots.as<char>(); tb.tb_len = D.tblen.as<int>(); tb.tb_start = D.tbstart.as<int>();
        tb.tb_locus = D.tblocus.as<int>(); tb.tb_flag = D.tbflag.as<int>(); tb.tb_energy = D.tbenergy.as<int>();
        tb.stack_scratch = D.stackscr.as<int>();
        CK(launch_traceback(tb, st));
        CK(launch_emit(tb, st));
        unsigned long long nhits = 0, abytes = 0;
        if (ntb) {
            k_emit_sizes<<<(unsigned)((ntb + 1 + 255) / 256), 256, 0, st>>>(tb.tb_flag, tb.tb_len, D.scan_in.as<unsigned long long>(),
                                                                             D.scan_out.as<unsigned long long>(), ntb);
            CK(cudaGetLastError());
            CK(exclusive_scan(D, D.scan_in.as<unsigned long long>(), D.ssoff.as<unsigned long long>(), (size_t)ntb + 1));
            CK(exclusive_scan(D, D.scan_out.as<unsigned long long>(), D.hitidx.as<unsigned long long>(), (size_t)ntb + 1));
            CK(cudaMemcpyAsync(&hs[1], D.ssoff.as<unsigned long long>() + ntb, 8, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(&hs[2], D.hitidx.as<unsigned long long>() + ntb, 8, cudaMemcpyDeviceToHost, st));
            out.st.kernel_launches += 7;
        }
        CK(cudaMemcpyAsync(&hs[3], D.fail.p, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (*(int *)&hs[3]) { out.err = MIRFOLD_ERR_BACKTRACK; out.errmsg = "traceback found no decomposition"; return; }
        if (ntb) { abytes = hs[1]; nhits = hs[2]; }
        CK(D.o_hits.ensure(nhits * sizeof(mirfold_hit) + 16)); CK(D.o_arena.ensure(abytes + 16));
        CK(launch_pack(tb, D.ssoff.as<unsigned long long>(), D.hitidx.as<unsigned long long>(), D.o_arena.as<char>(),
                       D.o_hits.as<mirfold_hit>(), out.arena_bytes, st));
        out.st.kernel_launches += 1;
        CK(cudaEventRecord(D.ev[5], st));
        ht.mark("traceback..pack (sync inside)");
        // ---- download
        if (download) {
            // per-locus first-hit index = hitidx[tb_base[l]] -> gather on host from two small arrays
            CK(D.h_out.ensure((size_t)(nl + 1) * 8 + (size_t)nl * 4 + 64));
            unsigned long long *h_tbbase = D.h_out.as<unsigned long long>();
            int *h_total = (int *)(h_tbbase + nl + 1);
            // hit table: the device wrote finished mirfold_hit records (ss_off already includes this chunk's
            // arena base); they land in the pinned table the result will own -- no per-hit host work
            const uint64_t hbase = out.nhits;
            if (out.hits.cap < (hbase + nhits) * sizeof(mirfold_hit) + 16) {
                HBuf bigger;
                CK(bigger.ensure((hbase + nhits) * sizeof(mirfold_hit) * (hbase ? 2 : 1) + 16));
                if (hbase) memcpy(bigger.p, out.hits.p, hbase * sizeof(mirfold_hit));
                out.hits.release();
                out.hits = bigger;
            }
            if (nhits)
                CK(cudaMemcpyAsync(out.hits.as<mirfold_hit>() + hbase, D.o_hits.p, nhits * sizeof(mirfold_hit), cudaMemcpyDeviceToHost, st));
            // hit index at each locus boundary: hitidx[tb_base[l]] (device gather -> reuse scan_in)
            {
                k_gather_bounds<<<(nl + 1 + 255) / 256, 256, 0, st>>>(tb.tb_base, D.hitidx.as<unsigned long long>(),
                                                                      D.scan_in.as<unsigned long long>(), nl);
                CK(cudaGetLastError());
                out.st.kernel_launches += 1;
                CK(cudaMemcpyAsync(h_tbbase, D.scan_in.p, (size_t)(nl + 1) * 8, cudaMemcpyDeviceToHost, st));
            }
            // totals: F[seq_off + 1] per locus -> strided; copy via 2D memcpy is awkward, use gather kernel
            {
                CK(D.tbcount.ensure((size_t)nl * 4));
                k_gather_totals<<<(nl + 255) / 256, 256, 0, st>>>(dl, D.F.as<int>(), D.tbcount.as<int>(), nl);
                CK(cudaGetLastError());
                out.st.kernel_launches += 1;
                CK(cudaMemcpyAsync(h_total, D.tbcount.p, (size_t)nl * 4, cudaMemcpyDeviceToHost, st));
            }
            // arena: straight into the pinned result arena
            const uint64_t abase = out.arena_bytes;
            if (chunks.size() == 1) {
                CK(out.arena.ensure(abytes + 16));
            } else {
                // multi-chunk: grow by reallocating pinned memory (rare path)
                if (out.arena.cap < abase + abytes + 16) {
                    HBuf bigger;
                    CK(bigger.ensure((abase + abytes) * 2 + 16));
                    if (abase) memcpy(bigger.p, out.arena.p, abase);
                    out.arena.release();
                    out.arena = bigger;
                }
            }
            if (abytes) CK(cudaMemcpyAsync(out.arena.as<char>() + abase, D.o_arena.p, abytes, cudaMemcpyDeviceToHost, st));
            CK(cudaEventRecord(D.ev[6], st));
            CK(cudaStreamSynchronize(st));
            out.st.d2h_bytes += nhits * sizeof(mirfold_hit) + (uint64_t)(nl + 1) * 8 + (uint64_t)nl * 4 + abytes + 32;
            ht.mark("download + sync");
            out.nhits = hbase + nhits;
            for (int k = 0; k < nl; k++) {
                out.recs.push_back(loci[cb + k].rec);
                out.rec_hit_begin.push_back(hbase + h_tbbase[k]);
                out.rec_hit_count.push_back((uint32_t)(h_tbbase[k + 1] - h_tbbase[k]));
                out.rec_total.push_back(h_total[k]);
            }
            out.arena_bytes = abase + abytes;
            ht.mark("hit records");
        } else {
            CK(cudaEventRecord(D.ev[6], st));
            CK(cudaStreamSynchronize(st));
            out.arena_bytes += abytes;
            out.st.d2h_bytes += 32;
            // count only
            out.rec_hit_begin.push_back(nhits);
        }
        float ms = 0;
        cudaEventElapsedTime(&ms, D.ev[0], D.ev[1]); out.st.ms_h2d += ms;
        cudaEventElapsedTime(&ms, D.ev[2], D.ev[3]); out.st.ms_fill += ms;
        cudaEventElapsedTime(&ms, D.ev[3], D.ev[4]); out.st.ms_f3 += ms;
        cudaEventElapsedTime(&ms, D.ev[4], D.ev[5]); out.st.ms_trace += ms;
        cudaEventElapsedTime(&ms, D.ev[5], D.ev[6]); out.st.ms_d2h += ms;
        cudaEventElapsedTime(&ms, D.ev[1], D.ev[5]); out.st.ms_device += ms;
    }
    out.st.ms_total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

}  // namespace

// ====================================================================================== C ABI
extern "C" {

const char *mirfold_version(void) { return "mirfold " MIRFOLD_VERSION " sm_100a " MIRFOLD_PARAMSET_DEFAULT; }

const char *mirfold_strerror(int code)
{
    switch (code) {
    case MIRFOLD_OK: return "ok";
    case MIRFOLD_ERR_NO_DEVICE: return "no usable CUDA device (libmirfold has no CPU fallback)";
    case MIRFOLD_ERR_CUDA: return "CUDA runtime error";
    case MIRFOLD_ERR_ARG: return "invalid argument";
    case MIRFOLD_ERR_PARAMSET: return "unknown energy parameter set";
    case MIRFOLD_ERR_BACKTRACK: return "backtrack failed";
    case MIRFOLD_ERR_NOMEM: return "out of memory";
    default: return "unknown error";
    }
}

const char *mirfold_last_error(const mirfold_ctx *ctx) { return ctx ? ctx->last_error.c_str() : ""; }

int mirfold_open(mirfold_ctx **pctx, const int *device_ids, int n_devices, const char *param_set)
{
    if (!pctx) return MIRFOLD_ERR_ARG;
    *pctx = nullptr;
    if (param_set && strcmp(param_set, MIRFOLD_PARAMSET_DEFAULT) != 0) return MIRFOLD_ERR_PARAMSET;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return MIRFOLD_ERR_NO_DEVICE; }
    std::vector<int> ids;
    if (!device_ids || n_devices <= 0) {
        int cur = 0;
        if (cudaGetDevice(&cur) != cudaSuccess) return MIRFOLD_ERR_NO_DEVICE;
        ids.push_back(cur);
    } else {
        for (int k = 0; k < n_devices; k++) {
            if (device_ids[k] < 0 || device_ids[k] >= ndev) return MIRFOLD_ERR_NO_DEVICE;
            ids.push_back(device_ids[k]);
        }
    }
    mirfold_ctx *ctx = new mirfold_ctx();
    DevParams *hp = new DevParams();
    build_params(*hp);
    int rc = MIRFOLD_OK;
    for (int id : ids) {
        Device D;
        D.id = id;
        cudaError_t e = cudaSetDevice(id);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&D.stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&D.side, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaMalloc(&D.dP, sizeof(DevParams));
        if (e == cudaSuccess) e = cudaMemcpy(D.dP, hp, sizeof(DevParams), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = fill_configure_device();
        for (auto &ev : D.ev) if (e == cudaSuccess) e = cudaEventCreate(&ev);
        size_t fr = 0, tot = 0;
        if (e == cudaSuccess) e = cudaMemGetInfo(&fr, &tot);
        if (e != cudaSuccess) { ctx->last_error = cudaGetErrorString(e); rc = MIRFOLD_ERR_CUDA; D.release(); break; }
        const char *env = getenv("MIRFOLD_MEM_BUDGET_MB");
        D.mem_budget = env ? (size_t)atoll(env) << 20 : std::min<size_t>((size_t)(fr * 0.55), (size_t)64 << 30);
        ctx->devs.push_back(D);
    }
    delete hp;
    if (rc != MIRFOLD_OK) { for (auto &D : ctx->devs) D.release(); delete ctx; return rc; }
    *pctx = ctx;
    return MIRFOLD_OK;
}

void mirfold_close(mirfold_ctx *ctx)
{
    if (!ctx) return;
    for (auto &D : ctx->devs) { cudaSetDevice(D.id); D.release(); }
    for (auto &b : ctx->arena_pool) b.release();
    for (auto &b : ctx->hits_pool) b.release();
    delete ctx;
}

static void add_stats(mirfold_stats &a, const mirfold_stats &b)
{
    a.ms_h2d = std::max(a.ms_h2d, b.ms_h2d); a.ms_fill = std::max(a.ms_fill, b.ms_fill);
    a.ms_f3 = std::max(a.ms_f3, b.ms_f3); a.ms_trace = std::max(a.ms_trace, b.ms_trace);
    a.ms_d2h = std::max(a.ms_d2h, b.ms_d2h); a.ms_device = std::max(a.ms_device, b.ms_device);
    a.nt += b.nt; a.cells += b.cells; a.tracebacks += b.tracebacks; a.kernel_launches += b.kernel_launches;
    a.h2d_bytes += b.h2d_bytes; a.d2h_bytes += b.d2h_bytes; a.n_chunks += b.n_chunks; a.fill_units += b.fill_units;
}

static int fold_impl(mirfold_ctx *ctx, const char *seqs, const uint64_t *seq_off, uint32_t nseq, int span_L,
                     const char *d_raw, bool download, void *stream, uint32_t flags, mirfold_result **out)
{
    static const bool env_wide = getenv("MIRFOLD_FORCE_WIDE") != nullptr;   // A/B runs of bench.py
    const bool force_wide = (flags & MIRFOLD_FLAG_WIDE) != 0 || env_wide || span_L > MF16_MAX_SPAN;
    if (!ctx || !out || !seq_off || (!seqs && !d_raw && nseq)) return MIRFOLD_ERR_ARG;
    if (span_L < 5 || span_L > MF_MAX_SPAN) { ctx->last_error = "span_L out of range [5, 4096]"; return MIRFOLD_ERR_ARG; }
    *out = nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    const int G = d_raw ? 1 : (int)ctx->devs.size();
    // ---- shard records over devices: greedy LPT on DP cells; no collectives
    std::vector<std::vector<uint32_t>> shard(G);
    if (G == 1) {
        shard[0].resize(nseq);
        for (uint32_t r = 0; r < nseq; r++) shard[0][r] = r;
    } else {
        std::vector<std::pair<uint64_t, uint32_t>> w(nseq);
        for (uint32_t r = 0; r < nseq; r++) w[r] = {cells_of((int)(seq_off[r + 1] - seq_off[r]), span_L), r};
        std::stable_sort(w.begin(), w.end(), [](const auto &a, const auto &b) { return a.first > b.first; });
        std::vector<uint64_t> load(G, 0);
        for (auto &x : w) {
            int g = (int)(std::min_element(load.begin(), load.end()) - load.begin());
            shard[g].push_back(x.second);
            load[g] += x.first + 1;
        }
        for (auto &s : shard) std::sort(s.begin(), s.end());
    }
    std::vector<Partial> parts(G);
    {
        std::lock_guard<std::mutex> lk(ctx->pool_mu);
        for (int g = 0; g < G && !ctx->arena_pool.empty(); g++) { parts[g].arena = ctx->arena_pool.back(); ctx->arena_pool.pop_back(); }
        for (int g = 0; g < G && !ctx->hits_pool.empty(); g++) { parts[g].hits = ctx->hits_pool.back(); ctx->hits_pool.pop_back(); }
    }
    if (G == 1) run_device(ctx->devs[0], seqs, seq_off, shard[0], span_L, d_raw, download, (cudaStream_t)stream, force_wide, parts[0]);
    else {
        std::vector<std::thread> th;
        for (int g = 0; g < G; g++)
            th.emplace_back([&, g] { run_device(ctx->devs[g], seqs, seq_off, shard[g], span_L, nullptr, download, nullptr, force_wide, parts[g]); });
        for (auto &t : th) t.join();
    }
    for (int g = 0; g < G; g++)
        if (parts[g].err != MIRFOLD_OK) {
            ctx->last_error = parts[g].errmsg;
            for (auto &p : parts) { p.arena.release(); p.hits.release(); }
            return parts[g].err;
        }
    HostTimer ht;
    // ---- gather in input order
    ResultOwner *R = new ResultOwner();
    R->ctx = ctx;
    R->hit_begin.assign((size_t)nseq + 1, 0);
    R->hit_count.assign((size_t)nseq + 1, 0);
    R->totals.assign(nseq, 0);
    mirfold_stats st{};
    uint64_t nhits = 0;
    if (download) {
        // Hits stay in the order the devices produced them (locus order inside a device, devices
        // concatenated); a record finds its run through hit_begin / hit_count.
        std::vector<uint64_t> arena_base(G, 0), hit_base(G, 0);
        uint64_t abytes = 0;
        for (int g = 0; g < G; g++) { arena_base[g] = abytes; abytes += parts[g].arena_bytes; hit_base[g] = nhits; nhits += parts[g].nhits; }
        for (int g = 0; g < G; g++)
            for (size_t k = 0; k < parts[g].recs.size(); k++) {
                const uint32_t r = parts[g].recs[k];
                R->hit_begin[r] = hit_base[g] + parts[g].rec_hit_begin[k];
                R->hit_count[r] = parts[g].rec_hit_count[k];
                R->totals[r] = parts[g].rec_total[k];
            }
        if (G == 1) {
            R->hits = parts[0].hits;    // pinned table moves into the result
            parts[0].hits = HBuf();
            R->pub.hits = R->hits.as<mirfold_hit>();
        } else {
            // multi-device: one table and one arena for the caller; every device's part is copied (and its
            // ss_off rebased) by its own thread into uninitialised buffers -- no zero fill, no serial pass
            R->hits_m = (mirfold_hit *)malloc(sizeof(mirfold_hit) * (size_t)(nhits ? nhits : 1));
            R->arena_m = (char *)malloc((size_t)abytes + 1);
            if (!R->hits_m || !R->arena_m) {
                free(R->hits_m); free(R->arena_m);
                for (auto &p : parts) { p.arena.release(); p.hits.release(); }
                delete R;
                return MIRFOLD_ERR_NOMEM;
            }
            std::vector<std::thread> th;
            for (int g = 0; g < G; g++)
                th.emplace_back([&, g] {
                    const mirfold_hit *src = parts[g].hits.as<mirfold_hit>();
                    mirfold_hit *dst = R->hits_m + hit_base[g];
                    const uint64_t ab = arena_base[g];
                    for (uint64_t h = 0; h < parts[g].nhits; h++) { dst[h] = src[h]; dst[h].ss_off += ab; }
                    if (parts[g].arena_bytes) memcpy(R->arena_m + ab, parts[g].arena.p, parts[g].arena_bytes);
                });
            for (auto &t : th) t.join();
            R->arena_m[abytes] = 0;
            for (int g = 0; g < G; g++) {
                std::lock_guard<std::mutex> lk(ctx->pool_mu);
                ctx->hits_pool.push_back(parts[g].hits);
                parts[g].hits = HBuf();
                ctx->arena_pool.push_back(parts[g].arena);
                parts[g].arena = HBuf();
            }
            R->pub.hits = R->hits_m;
            R->pub.ss_arena = R->arena_m;
        }
        if (G == 1) {
            R->arena = parts[0].arena;  // pinned buffer moves into the result
            parts[0].arena = HBuf();
            R->pub.ss_arena = R->arena.as<char>();
        }
        R->pub.ss_bytes = abytes;
    } else {
        for (int g = 0; g < G; g++) { for (uint64_t v : parts[g].rec_hit_begin) nhits += v; R->pub.ss_bytes += parts[g].arena_bytes; }
        for (auto &p : parts) {
            std::lock_guard<std::mutex> lk(ctx->pool_mu);
            if (p.arena.p) { ctx->arena_pool.push_back(p.arena); p.arena = HBuf(); }
            if (p.hits.p) { ctx->hits_pool.push_back(p.hits); p.hits = HBuf(); }
        }
        R->pub.ss_arena = nullptr;
        R->pub.hits = nullptr;
    }
    ht.mark("gather in input order");
    for (int g = 0; g < G; g++) add_stats(st, parts[g].st);
    st.n_devices = G;
    st.ms_total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    R->pub.nseq = nseq;
    R->pub.nhits = nhits;
    R->pub.hit_begin = R->hit_begin.data();
    R->pub.hit_count = R->hit_count.data();
    R->pub.total_mfe_dcal = R->totals.data();
    R->pub.stats = st;
    *out = &R->pub;
    return MIRFOLD_OK;
}

int mirfold_fold(mirfold_ctx *ctx, const char *seqs, const uint64_t *seq_off, uint32_t nseq, int span_L, uint32_t flags,
                 mirfold_result **out)
{
    return fold_impl(ctx, seqs, seq_off, nseq, span_L, nullptr, true, nullptr, flags, out);
}

int mirfold_fold_device(mirfold_ctx *ctx, const void *d_seqs, const void *d_seq_off, const uint64_t *h_seq_off,
                        uint32_t nseq, int span_L, uint32_t flags, void *stream, mirfold_result **out)
{
    (void)d_seq_off;
    if (!d_seqs) return MIRFOLD_ERR_ARG;
    return fold_impl(ctx, nullptr, h_seq_off, nseq, span_L, (const char *)d_seqs, false, stream, flags, out);
}

void mirfold_free_result(mirfold_result *res)
{
    if (!res) return;
    ResultOwner *R = reinterpret_cast<ResultOwner *>(res);
    if (R->arena.p && R->ctx) {
        std::lock_guard<std::mutex> lk(R->ctx->pool_mu);
        if (R->ctx->arena_pool.size() < 4) { R->ctx->arena_pool.push_back(R->arena); R->arena = HBuf(); }
    }
    if (R->hits.p && R->ctx) {
        std::lock_guard<std::mutex> lk(R->ctx->pool_mu);
        if (R->ctx->hits_pool.size() < 4) { R->ctx->hits_pool.push_back(R->hits); R->hits = HBuf(); }
    }
    R->arena.release();
    R->hits.release();
    free(R->hits_m);
    free(R->arena_m);
    delete R;
}

int mirfold_format_records(const mirfold_result *res, const char *seqs, const uint64_t *seq_off, uint32_t nseq, char **text,
                           uint64_t **rec_off)
{
    if (!res || !text || !rec_off || !seq_off || (!seqs && nseq) || res->nseq != nseq) return MIRFOLD_ERR_ARG;
    if (res->nhits && !res->ss_arena) return MIRFOLD_ERR_ARG;   // device-resident results carry no structures
    *text = nullptr; *rec_off = nullptr;
    uint64_t *off = (uint64_t *)malloc(sizeof(uint64_t) * ((size_t)nseq + 1));
    if (!off) return MIRFOLD_ERR_NOMEM;
    // every "(%6.2f)" field is 8 characters for |E| < 1000 kcal/mol and grows with the integer part beyond;
    // sizes are computed exactly with the same snprintf calls that fill the buffer
    auto hit_line = [&](char *dst, size_t cap, const mirfold_hit &h) {
        // dot-bracket, then " (%6.2f) %4d\n"
        if (dst) memcpy(dst, res->ss_arena + h.ss_off, (size_t)h.len);
        char tail[64];
        const int k = snprintf(tail, sizeof tail, " (%6.2f) %4d\n", h.mfe_dcal / 100., h.start);
        if (dst) memcpy(dst + h.len, tail, (size_t)k);
        (void)cap;
        return (size_t)h.len + (size_t)k;
    };
    auto total_line = [&](char *dst, uint32_t r) {
        const size_t n = (size_t)(seq_off[r + 1] - seq_off[r]);
        if (dst) {
            const char *src = seqs + seq_off[r];
            for (size_t k = 0; k < n; k++) {
                char ch = src[k];
                if (ch >= 'a' && ch <= 'z') ch = (char)(ch - 32);
                dst[k] = ch == 'T' ? 'U' : ch;
            }
            dst[n] = '\n';
        }
        char tail[64];
        const int k = snprintf(tail, sizeof tail, " (%6.2f)\n", res->total_mfe_dcal[r] / 100.);
        if (dst) memcpy(dst + n + 1, tail, (size_t)k);
        return n + 1 + (size_t)k;
    };
    const unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    const unsigned nthr = nseq < 256 ? 1u : hw;
    auto for_ranges = [&](auto &&fn) {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nthr; t++) {
            const uint32_t lo = (uint32_t)((uint64_t)nseq * t / nthr), hi = (uint32_t)((uint64_t)nseq * (t + 1) / nthr);
            if (nthr == 1) fn(lo, hi);
            else th.emplace_back(fn, lo, hi);
        }
        for (auto &x : th) x.join();
    };
    // pass 1: sizes
    for_ranges([&](uint32_t lo, uint32_t hi) {
        for (uint32_t r = lo; r < hi; r++) {
            size_t b = 0;
            for (uint64_t h = res->hit_begin[r]; h < res->hit_begin[r] + res->hit_count[r]; h++) b += hit_line(nullptr, 0, res->hits[h]);
            off[r + 1] = b + total_line(nullptr, r);
        }
    });
    off[0] = 0;
    for (uint32_t r = 0; r < nseq; r++) off[r + 1] += off[r];
    char *buf = (char *)malloc((size_t)off[nseq] + 1);
    if (!buf) { free(off); return MIRFOLD_ERR_NOMEM; }
    // pass 2: fill
    for_ranges([&](uint32_t lo, uint32_t hi) {
        for (uint32_t r = lo; r < hi; r++) {
            char *dst = buf + off[r];
            for (uint64_t h = res->hit_begin[r]; h < res->hit_begin[r] + res->hit_count[r]; h++) dst += hit_line(dst, 0, res->hits[h]);
            total_line(dst, r);
        }
    });
    buf[off[nseq]] = 0;
    *text = buf; *rec_off = off;
    return MIRFOLD_OK;
}

void mirfold_free_text(char *text, uint64_t *rec_off)
{
    free(text);
    free(rec_off);
}

int mirfold_debug_matrices(mirfold_ctx *ctx, const char *seq, uint32_t n, int span_L, uint32_t flags, int32_t *c, int32_t *m,
                           int32_t *f3)
{
    if (!ctx || !seq || n < 5 || !c || !m || !f3) return MIRFOLD_ERR_ARG;
    Device &D = ctx->devs[0];
#undef CK
#define CK(call)                                                                  \
    do {                                                                          \
        cudaError_t e_ = (call);                                                  \
        if (e_ != cudaSuccess) { ctx->last_error = cudaGetErrorString(e_); return MIRFOLD_ERR_CUDA; } \
    } while (0)
    CK(cudaSetDevice(D.id));
    cudaStream_t st = D.stream;
    LocusDesc d{};
    const unsigned long long cells = shape_locus(d, (int)n, span_L);
    std::vector<LocusDesc> units;
    int bucket_first[5], max_n = 0;
    const unsigned long long ring_elems = build_fill_units(&d, 1, units, bucket_first, max_n);
    const int nu = (int)units.size();
    CK(D.raw.ensure(n)); CK(D.loci.ensure(sizeof d)); CK(D.codes.ensure(n + 3)); CK(D.F.ensure((n + 3) * 4));
    CK(D.C.ensure(cells * 4)); CK(D.M.ensure(cells * 4)); CK(D.Mp.ensure(cells * 4));
    CK(D.ring.ensure((size_t)ring_elems * 4));
    CK(D.units.ensure(sizeof(LocusDesc) * (size_t)nu + 64));
    CK(cudaMemcpyAsync(D.raw.p, seq, n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(D.loci.p, &d, sizeof d, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(D.units.p, units.data(), sizeof(LocusDesc) * (size_t)nu, cudaMemcpyHostToDevice, st));
    const LocusDesc *dl = D.loci.as<LocusDesc>();
    CK(launch_prepare(D.raw.as<char>(), dl, 1, n + 3, D.codes.as<unsigned char>(), D.F.as<int>(), st));
    CK(D.fillflags.ensure((size_t)nu * 4 + 4));
    CK(cudaMemsetAsync(D.fillflags.p, 0, (size_t)nu * 4 + 4, st));
    FillLaunch fa{D.units.as<LocusDesc>(), nu, max_n, D.codes.as<unsigned char>(), D.C.as<int>(), D.M.as<int>(), D.ring.as<int>(),
                  D.Mp.as<unsigned int>(), D.dP, {0, 0, 0, 0, 0}, D.fillflags.as<int>(),
                  ((flags & MIRFOLD_FLAG_WIDE) || span_L > MF16_MAX_SPAN) ? 1 : 0, env_opts()};
    for (int b = 0; b < 5; b++) fa.bucket_first[b] = bucket_first[b];
    CK(launch_fill(fa, st));
    CK(launch_f3(dl, 1, d.n > MF_TILE_LEN ? 1 : 0, d.Ls, D.codes.as<unsigned char>(), D.C.as<int>(), D.F.as<int>(), D.dP, st));
    std::vector<int> hc(cells), hm(cells), hf(n + 3);
    CK(cudaMemcpyAsync(hc.data(), D.C.p, cells * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(hm.data(), D.M.p, cells * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(hf.data(), D.F.p, (n + 3) * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const int W = d.Ls + 6;
    for (size_t k = 0; k < (size_t)(n + 2) * W; k++) c[k] = m[k] = MF_INF;
    for (int dd = 4; dd <= d.dmax; dd++)
        for (int i = 1; i <= (int)n - dd; i++) {
            // owner tile of row i (same mapping as band_row_base on the device)
            unsigned long long base = (unsigned long long)(i - 1);
            if (d.tile_last) {
                const int t = std::min((i - 1) / d.tile_step, d.tile_last);
                const int a = std::min(t * d.tile_step, (int)n - MF_TILE_LEN);
                base = (unsigned long long)t * band_elems(MF_TILE_LEN, d.dmax) + (unsigned long long)(i - 1 - a);
            }
            c[(size_t)i * W + dd] = hc[base + band_doff(d.stride, dd)];
            m[(size_t)i * W + dd] = hm[base + band_doff(d.stride, dd)];
        }
    for (uint32_t k = 0; k < n + 3; k++) f3[k] = hf[k];
    f3[n + 3] = 0;
    return MIRFOLD_OK;
}

int mirfold_int_peak(mirfold_ctx *ctx, double *addmin_terms_per_s, double *dpx_terms_per_s)
{
    if (!ctx || !addmin_terms_per_s || !dpx_terms_per_s) return MIRFOLD_ERR_ARG;
    Device &D = ctx->devs[0];
    cudaSetDevice(D.id);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, D.id);
    cudaError_t e = run_int_peak(D.stream, sms, addmin_terms_per_s, dpx_terms_per_s);
    if (e != cudaSuccess) { ctx->last_error = cudaGetErrorString(e); return MIRFOLD_ERR_CUDA; }
    return MIRFOLD_OK;
}

int mirfold_duplex(mirfold_ctx *ctx, const char *ss_arena, uint64_t ss_bytes, const mirfold_duplex_query *queries,
                   uint64_t nq, mirfold_duplex_verdict *verdicts)
{
    if (!ctx || (!ss_arena && ss_bytes) || (!queries && nq) || (!verdicts && nq)) return MIRFOLD_ERR_ARG;
    if (nq == 0) return MIRFOLD_OK;
    int maxlen = 1;
    for (uint64_t k = 0; k < nq; k++) {
        if (queries[k].ss_len < 0 || queries[k].ss_off + (uint64_t)queries[k].ss_len > ss_bytes) {
            ctx->last_error = "mirfold_duplex: query structure outside the arena";
            return MIRFOLD_ERR_ARG;
        }
        maxlen = std::max(maxlen, queries[k].ss_len);
    }
    if (maxlen > 32000) { ctx->last_error = "mirfold_duplex: structure longer than 32000"; return MIRFOLD_ERR_ARG; }
    Device &D = ctx->devs[0];
    cudaStream_t st = D.stream;
    CK(cudaSetDevice(D.id));
    CK(D.o_arena.ensure(ss_bytes + 16));
    CK(D.scan_in.ensure(nq * sizeof(mirfold_duplex_query)));
    CK(D.scan_out.ensure(nq * sizeof(mirfold_duplex_verdict)));
    CK(cudaMemcpyAsync(D.o_arena.p, ss_arena, ss_bytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(D.scan_in.p, queries, nq * sizeof(mirfold_duplex_query), cudaMemcpyHostToDevice, st));
    CK(launch_duplex(D.o_arena.as<char>(), D.scan_in.as<mirfold_duplex_query>(), nq, D.scan_out.as<mirfold_duplex_verdict>(),
                     (maxlen + 7) & ~7, st));
    CK(cudaMemcpyAsync(verdicts, D.scan_out.p, nq * sizeof(mirfold_duplex_verdict), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return MIRFOLD_OK;
}

const char *mirfold_duplex_fail_name(int code)
{
    static const char *names[] = {"PASS",
                                  "FAIL_STRUCTURE_MATCHED_BASES",
                                  "FAIL_STRUCTURE_MATURE_NOT_IN_FOLD_REGION",
                                  "FAIL_STRUCTURE_MATURE_NOT_IN_ONE_ARM",
                                  "FAIL_STRUCTURE_MATURE_MATCH_SMALL_THAN_14",
                                  "FAIL_STRUCTURE_MATURE_STAR_OVERLAP",
                                  "FAIL_STRUCTURE_STAR_OUT_OF_FOLD_REGION",
                                  "FAIL_STRUCTURE_STAR_NOT_IN_ONE_ARM",
                                  "FAIL_STRUCTURE_TOO_MANY_BULGE_OR_LOOP",
                                  "FAIL_STRUCTURE_MAX_BULGE_LARGE_THAN_2",
                                  "FAIL_STRUCTURE_TOTAL_LOOP_SIZE_LARGER_THAN_5",
                                  "FAIL_STRUCTURE_NUM_BULGE_MORE_THAN_2"};
    if (code >= 0 && code < 12) return names[code];
    if (code == 100) return "EXCEPTION_UNBALANCED_STRUCTURE";
    return "";
}

}  // extern "C"
