// host.cu -- libmirfold host runtime: contexts, memory pools, the two-lane chunk pipeline, multi-GPU
// sharding and the fold entry points of the C ABI declared in include/mirfold.h.
//
// Replaces fold_use_RNALfold()/fold() (/root/reference/miR_PREFeR.py:3047-3119): where the reference
// forks one RNALfold process per FASTA shard and folds it 2*CHECKPOINT_SIZE lines at a time, this
// runtime shards records over GPUs by DP work (no collectives: loci are independent), cuts every
// shard into chunks and keeps two chunks in flight per device on two "lanes" (complete buffer sets):
// while one lane runs f3 / traceback / pack / download of chunk k on a high-priority stream, the other
// lane's band fill of chunk k+1 already occupies the SMs the last wave of chunk k left idle.  Results
// are downloaded straight into buffers shared by all devices of the call (no merge pass) or handed
// to a callback chunk by chunk (mirfold_fold_stream).  There is no CPU fallback.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cub/device/device_scan.cuh>

#include "../../include/mirfold.h"
#include "mirfold_internal.cuh"

#define MIRFOLD_VERSION "0.2.0"

void build_params(DevParams &P);   // params.cu

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
        if (e != cudaSuccess) { cudaGetLastError(); e = cudaMalloc(&p, bytes); want = bytes; }
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
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocPortable);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return (T *)p; }
};

struct Locus {
    uint32_t rec;
    int n;
    uint64_t cells;
};

// host-side description of one chunk on a lane
struct Prep {
    size_t cb = 0, ce = 0;   // [cb, ce) into the device's locus list
    int nl = 0, nu = 0, max_n = 0, max_Ls = 0, n_long = 0;
    int bucket_first[6] = {0, 0, 0, 0, 0, 0};
    unsigned long long seq_acc = 0, band_acc = 0, raw_acc = 0, list_acc = 0, ring_acc = 0;
};

struct OverflowChunk {   // a chunk that did not fit the shared result buffers (capacity is an estimate)
    HBuf hits, arena;
    ~OverflowChunk() { hits.release(); arena.release(); }
    uint64_t nhits = 0, abytes = 0;
    std::vector<uint32_t> recs, count;
    std::vector<uint64_t> begin;
};

// A lane = one complete set of pipeline buffers and streams; a device alternates chunks between two lanes.
struct Lane {
    cudaStream_t s_lo = nullptr;   // uploads, encode, band fill
    cudaStream_t s_hi = nullptr;   // f3, plan, traceback, pack, downloads (high priority: gets freed SM slots first)
    cudaStream_t side = nullptr;   // small fill buckets run here, concurrently with the big one
    DBuf units, raw, loci, codes, F, C, M, Mp, Ib, ring, fillflags, tbcount, tbbase, listoff, startlist, scan_in, scan_out, scan_tmp;
    DBuf slots, tblen, tbstart, tblocus, tbflag, tbenergy, stackscr, fail, ssoff, hitidx;
    DBuf o_hits, o_arena, bounds, totals;
    DBuf c_counts, c_soff, c_structs, c_nq, c_sbytes, c_voff, c_aoff, c_lsb, c_verd, c_vmat, c_arena, c_sout;   // fused stage 1 + 3
    HBuf hc_structs, hc_arena, hc_verd, hc_vmat, hc_voff, hc_lsb;
    uint64_t c_ns = 0, c_nv = 0, c_nb = 0;
    HBuf h_raw, h_loci, h_units, h_listoff, h_small, h_out, h_hits, h_arena;
    cudaEvent_t ev[12] = {};   // 0 h2d begin, 1 kernels begin, 2 fill begin, 3 fill end, 4 f3 end, 5 pack end, 6 d2h end, 7 done, 10/11 side fork/join
    // chunk in flight
    bool pending = false;
    Prep prep;
    TraceBuffers tb{};
    uint64_t nhits = 0, abytes = 0, hit_base = 0, arena_base = 0;
    std::unique_ptr<OverflowChunk> ovf;
    std::vector<uint32_t> s_rec, s_count;   // stream sink: per-record tables handed to the callback
    std::vector<uint64_t> s_begin;
    std::vector<int32_t> s_total;
    cudaError_t create(int prio_hi)
    {
        cudaError_t e = cudaStreamCreateWithFlags(&s_lo, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&s_hi, cudaStreamNonBlocking, prio_hi);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking);
        for (auto &x : ev) if (e == cudaSuccess) e = cudaEventCreate(&x);
        return e;
    }
    void release()
    {
        DBuf *all[] = {&units, &raw, &loci, &codes, &F, &C, &M, &Mp, &Ib, &ring, &fillflags, &tbcount, &tbbase, &listoff, &startlist, &scan_in,
                       &scan_out, &scan_tmp, &slots, &tblen, &tbstart, &tblocus, &tbflag, &tbenergy, &stackscr, &fail,
                       &ssoff, &hitidx, &o_hits, &o_arena, &bounds, &totals, &c_counts, &c_soff, &c_structs, &c_nq, &c_sbytes,
                       &c_voff, &c_aoff, &c_lsb, &c_verd, &c_vmat, &c_arena, &c_sout};
        for (DBuf *b : all) b->release();
        HBuf *hall[] = {&h_raw, &h_loci, &h_units, &h_listoff, &h_small, &h_out, &h_hits, &h_arena, &hc_structs, &hc_arena, &hc_verd,
                        &hc_vmat, &hc_voff, &hc_lsb};
        for (HBuf *b : hall) b->release();
        for (auto &e : ev) if (e) { cudaEventDestroy(e); e = nullptr; }
        if (s_lo) cudaStreamDestroy(s_lo);
        if (s_hi) cudaStreamDestroy(s_hi);
        if (side) cudaStreamDestroy(side);
        s_lo = s_hi = side = nullptr;
    }
};

struct Device {
    int id = 0;
    DevParams *dP = nullptr;
    size_t mem_budget = 0;
    Lane lane[2];
    cudaEvent_t ev_first = nullptr, ev_last = nullptr;
    DBuf duplex_arena, duplex_q, duplex_v;   // mirfold_duplex scratch
    DBuf cand_regions, cand_matures, cand_moff;   // the call's records (mirfold_fold_candidates)
    void release()
    {
        lane[0].release(); lane[1].release();
        duplex_arena.release(); duplex_q.release(); duplex_v.release();
        cand_regions.release(); cand_matures.release(); cand_moff.release();
        if (ev_first) cudaEventDestroy(ev_first);
        if (ev_last) cudaEventDestroy(ev_last);
        ev_first = ev_last = nullptr;
        if (dP) cudaFree(dP);
        dP = nullptr;
    }
};

}  // namespace

struct mirfold_ctx {
    std::vector<Device> devs;
    std::string last_error;
    std::vector<HBuf> arena_pool;  // pinned arenas returned by mirfold_free_result
    std::vector<HBuf> hits_pool;   // pinned hit tables returned by mirfold_free_result
    std::mutex pool_mu;
    int live_results = 0;          // results not yet freed (guarded by pool_mu)
    bool closed = false;           // mirfold_close() was called; the last mirfold_free_result deletes the context
    double arena_per_nt = 30.0;    // capacity estimates of the shared result buffers, raised when a call overflows them
    double hits_per_nt = 0.20;
};

struct mirfold_batch {
    mirfold_ctx *ctx = nullptr;
    uint32_t nseq = 0;
    int span_L = 0;
    std::vector<uint64_t> h_off;                   // copy of the caller's offsets
    std::vector<std::vector<uint32_t>> shard;      // per device, descending DP cells
    std::vector<uint64_t> raw_off;                 // per record: offset inside its device's buffer
    std::vector<void *> d_raw;                     // per device
};

namespace {

struct ResultOwner {  // lives right behind the public struct
    mirfold_result pub;
    mirfold_ctx *ctx;
    std::vector<uint64_t> hit_begin;
    std::vector<uint32_t> hit_count;
    std::vector<int32_t> totals;
    HBuf hits, arena;   // pinned, written directly by the devices' downloads
};

// result buffers shared by all devices of one mirfold_fold call
struct SharedOut {
    HBuf hits, arena;
    uint64_t hits_cap = 0, arena_cap = 0;     // records / bytes
    uint64_t hits_used = 0, arena_used = 0;
    std::mutex mu;
    std::vector<std::unique_ptr<OverflowChunk>> overflow;
    uint64_t *hit_begin = nullptr;
    uint32_t *hit_count = nullptr;
    int32_t *totals = nullptr;
    bool reserve(uint64_t nh, uint64_t ab, uint64_t &hb, uint64_t &abase)
    {
        std::lock_guard<std::mutex> lk(mu);
        if (hits_used + nh > hits_cap || arena_used + ab > arena_cap) return false;
        hb = hits_used; abase = arena_used;
        hits_used += nh; arena_used += ab;
        return true;
    }
};

enum { SINK_NONE = 0, SINK_SHARED = 1, SINK_STREAM = 2, SINK_CAND = 3 };

// fused stage 1 + 3: what the chunks of all devices append to (small: a few percent of the hit text)
struct CandCollector {
    std::mutex mu;
    std::vector<mirfold_structure> structs;
    std::vector<char> arena;
    std::vector<mirfold_duplex_verdict> verdicts;
    std::vector<uint32_t> verdict_mature, struct_count, verdict_count;
    std::vector<uint64_t> struct_begin, verdict_begin;
};
struct CandArgs {
    const mirfold_region *regions = nullptr;
    const mirfold_mature *matures = nullptr;
    const uint64_t *mature_off = nullptr;
    uint32_t nseq = 0;
    int minlen = 55, minloop = 3, min_mature = 18, max_mature = 24;
};

struct Job {   // one fold call; shared read-only by the device threads (sinks are synchronised)
    const char *seqs = nullptr;
    const uint64_t *h_off = nullptr;
    int L = 0;
    bool force_wide = false;
    bool serial = false;               // MIRFOLD_FLAG_SERIAL: one lane
    int sink = SINK_NONE;
    SharedOut *shared = nullptr;
    mirfold_chunk_fn fn = nullptr;
    void *user = nullptr;
    std::mutex *cb_mu = nullptr;
    std::atomic<int> *cb_abort = nullptr;
    const CandArgs *cand = nullptr;
    CandCollector *collect = nullptr;
};

struct DevOut {
    mirfold_stats st{};
    uint64_t nhits = 0, arena_bytes = 0;
    int err = MIRFOLD_OK;
    std::string errmsg;
};

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

uint64_t cells_of(int64_t n, int L)
{   // sum_{i=1}^{n-4} (min(n, i+L*) - i - 3); rows with i+L* <= n contribute L*-3, the rest n-i-3
    if (n < 5) return 0;
    const int64_t Ls = std::min<int64_t>(L, n);
    const int64_t full = std::max<int64_t>(0, std::min(n - 4, n - Ls));
    const int64_t m = n - 4 - full;
    return (uint64_t)(full * (Ls - 3) + m * (m + 1) / 2);
}

// ---- shard plan: records in descending DP cells (== descending length for a fixed span, ties by
// record index), greedy longest-processing-time assignment.  Every shard list stays in that order, which is
// also the order the device's CTA scheduler wants (largest first).
void plan_shards(const uint64_t *off, uint32_t nseq, int L, int G, std::vector<std::vector<uint32_t>> &shard, std::vector<uint64_t> &load)
{
    shard.assign((size_t)G, {});
    load.assign((size_t)G, 0);
    if (nseq == 0) return;
    std::vector<uint32_t> order(nseq);
    uint64_t maxn = 0;
    for (uint32_t r = 0; r < nseq; r++) maxn = std::max(maxn, off[r + 1] - off[r]);
    if (maxn <= 4ull * nseq + 65536) {   // counting sort by length, stable in the record index
        std::vector<uint32_t> cnt((size_t)maxn + 2, 0);
        for (uint32_t r = 0; r < nseq; r++) cnt[(size_t)(maxn - (off[r + 1] - off[r])) + 1]++;
        for (size_t k = 1; k < cnt.size(); k++) cnt[k] += cnt[k - 1];
        for (uint32_t r = 0; r < nseq; r++) order[cnt[(size_t)(maxn - (off[r + 1] - off[r]))]++] = r;
    } else {
        for (uint32_t r = 0; r < nseq; r++) order[r] = r;
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return off[a + 1] - off[a] > off[b + 1] - off[b]; });
    }
    if (G == 1) {
        for (uint32_t r : order) load[0] += cells_of((int64_t)(off[r + 1] - off[r]), L) + 1;
        shard[0].swap(order);
        return;
    }
    for (auto &s : shard) s.reserve(nseq / G + 16);
    for (uint32_t r : order) {
        int g = 0;
        for (int k = 1; k < G; k++) if (load[k] < load[g]) g = k;
        shard[g].push_back(r);
        load[g] += cells_of((int64_t)(off[r + 1] - off[r]), L) + 1;
    }
}

#define CK(call)                                                                                 \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess) {                                                                 \
            out.err = e_ == cudaErrorMemoryAllocation ? MIRFOLD_ERR_NOMEM : MIRFOLD_ERR_CUDA;    \
            out.errmsg = std::string(#call) + ": " + cudaGetErrorString(e_);                     \
            cudaGetLastError();                                                                  \
            return false;                                                                        \
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
// per-locus first-hit index = hitidx[tb_base[l]] and total line F[seq_off + 1]
__global__ void k_gather_bounds(const unsigned long long *tb_base, const unsigned long long *hitidx, const LocusDesc *loci,
                                const int *F, unsigned long long *out, int *totals, int nl)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k <= nl) out[k] = hitidx[tb_base[k]];
    if (k < nl) totals[k] = F[loci[k].seq_off + 1];
}

cudaError_t exclusive_scan(Lane &Ln, const unsigned long long *in, unsigned long long *out, size_t n, cudaStream_t st)
{
    size_t tmp = 0;
    cudaError_t e = cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, out, n, st);
    if (e != cudaSuccess) return e;
    e = Ln.scan_tmp.ensure(tmp);
    if (e != cudaSuccess) return e;
    return cub::DeviceScan::ExclusiveSum(Ln.scan_tmp.p, tmp, in, out, n, st);
}

// Band shape of one locus: stride bucket or, for a longer locus with a span that leaves at least
// MF_TILE_MIN_STEP owned rows per tile, overlapping tiles for the shared-memory kernels.  A tile owns TL - dmax of
// its TL rows, so spans from big_tile_min_span() on use the 864 bucket: at L = 500 a 608-nt tile owns 108 rows (the
// locus' cells are computed 3.3 times over), an 864-nt tile 364 (1.7 times); at L = 300 it is 1.49 against 1.27.  Loci
// of 609..864 nt are then one untiled unit of that bucket.  The 864 kernel is one CTA per SM and 13 % (L = 300) to
// 35 % (L = 500) slower per executed cell than the 608 kernel (two CTAs per SM hide each other's barriers), so narrower
// spans keep the 608-nt tiles.  Measured same-box (profiles/r02_v7_bigtile_ab.txt): long loci -6.6 % at L = 300, -29 % at L = 500.
// Returns the band elements (per int32 array) the locus occupies.
#define MF_TILE_MIN_STEP 64
#ifndef MF_BIG_TILE_MIN_SPAN
#define MF_BIG_TILE_MIN_SPAN 300
#endif
static int big_tile_min_span()
{
    static const int v = [] { const char *e = getenv("MIRFOLD_BIG_TILE_MIN_SPAN"); return e ? atoi(e) : MF_BIG_TILE_MIN_SPAN; }();   // A/B runs; a huge value disables the 864 bucket
    return v;
}
unsigned long long shape_locus(LocusDesc &d, int n, int L)
{
    d.n = n; d.Ls = std::min(L, n); d.dmax = std::min(d.Ls, n - 1);
    d.tile_last = 0; d.tile_step = 1 << 30; d.tile_rcp = 0;
    static const bool no_tiles = getenv("MIRFOLD_NO_TILES") != nullptr;   // A/B runs: long loci through k_fill_generic
    const bool big = n > MF_TILE_LEN && d.dmax >= big_tile_min_span() && !no_tiles &&
                     (n <= MF_TILE_LEN_BIG || MF_TILE_LEN_BIG - d.dmax >= MF_TILE_MIN_STEP);
    d.stride = band_stride_for(n, big);
    const int TL = big ? MF_TILE_LEN_BIG : MF_TILE_LEN;
    if (n > TL && d.dmax >= 4 && TL - d.dmax >= MF_TILE_MIN_STEP && !no_tiles) {
        const int S = TL - d.dmax;
        d.stride = TL;
        d.tile_step = S;
        d.tile_last = (n - TL + S - 1) / S;
        d.tile_rcp = (unsigned int)((0x100000000ULL + (unsigned long long)S - 1) / (unsigned long long)S);
        return (unsigned long long)(d.tile_last + 1) * band_elems(TL, d.dmax);
    }
    return band_elems(d.stride, d.dmax);
}
unsigned long long unit_ring_elems(int stride)
{
    return (unsigned long long)stride * (is_bucket_stride(stride) ? MF_RING_DML : MF_RING_PER_STRIDE);
}
static int unit_bucket(const LocusDesc &u)   // index into FillLaunch::bucket_first
{
    return !is_bucket_stride(u.stride) ? 0 : u.stride == MF_TILE_LEN_BIG ? 1 : u.stride == MF_TILE_LEN ? 2 : u.stride == 352 ? 3 : 4;
}
// Fill units of a chunk: one per untiled locus, one per tile otherwise, sorted by (bucket, descending n)
// so that the stride buckets are contiguous.  Assigns ring offsets; returns ring elements.
unsigned long long build_fill_units(const LocusDesc *loci, int nl, std::vector<LocusDesc> &units, int bucket_first[6], int &max_n)
{
    units.clear();
    max_n = 0;
    for (int k = 0; k < nl; k++) {
        const LocusDesc &d = loci[k];
        if (d.dmax < 4) continue;   // n < 5 never reaches here; keeps the kernels' assumptions explicit
        if (d.tile_last == 0) { units.push_back(d); continue; }
        for (int t = 0; t <= d.tile_last; t++) {
            LocusDesc u = d;
            const int TL = d.stride;
            const int a = std::min(t * d.tile_step, d.n - TL);
            u.n = TL; u.Ls = std::min(d.Ls, TL); u.dmax = d.dmax;
            u.seq_off = d.seq_off + (unsigned long long)a;
            u.band_off = d.band_off + (unsigned long long)t * band_elems(TL, d.dmax);
            u.tile_last = 0;
            units.push_back(u);
        }
    }
    std::stable_sort(units.begin(), units.end(), [](const LocusDesc &a, const LocusDesc &b) {
        const int ba = unit_bucket(a), bb = unit_bucket(b);
        return ba != bb ? ba < bb : a.n > b.n;
    });
    unsigned long long ring_acc = 0;
    int k = 0;
    const int nu = (int)units.size();
    bucket_first[0] = 0;
    for (int b = 0; b < 5; b++) {
        while (k < nu && unit_bucket(units[k]) <= b) k++;
        bucket_first[b + 1] = k;
    }
    for (auto &u : units) {
        u.ring_off = ring_acc;
        ring_acc += unit_ring_elems(u.stride);
        max_n = std::max(max_n, u.n);
    }
    return ring_acc;
}

// device bytes one locus needs on a lane: band (C, M, Mp), ring, descriptors, sequence-sized arrays and an
// estimate of the per-traceback buffers (slots, stacks, per-traceback ints: about one traceback per 8 nt)
size_t locus_bytes(int n, int L)
{
    LocusDesc d{};
    const unsigned long long be = shape_locus(d, n, L);
    const size_t per_tb = (size_t)((std::min(L, n) + 8) & ~3) + (size_t)(std::min(L, n) / 4 + 16) * 8 + 64;
    return (size_t)be * 13 + (size_t)(d.tile_last + 1) * unit_ring_elems(d.stride) * 4 +
           (size_t)(d.tile_last + 2) * sizeof(LocusDesc) + (size_t)n * 16 + (size_t)(n / 8 + 2) * per_tb + 4096;
}

// ------------------------------------------------------------------------------------------------------------
// One device's share of a fold call.
struct DevicePipeline {
    Device &D;
    const Job &J;
    const char *d_raw;            // != nullptr: raw sequences already on this device
    const uint64_t *raw_off;      // ... at these per-record offsets
    DevOut &out;
    std::vector<Locus> loci;      // descending DP cells
    struct Chunk { size_t begin, end; };
    std::vector<Chunk> chunks;
    int nlanes_ = 1;
    HostTimer ht;

    DevicePipeline(Device &D_, const Job &J_, const char *d_raw_, const uint64_t *raw_off_, DevOut &out_)
        : D(D_), J(J_), d_raw(d_raw_), raw_off(raw_off_), out(out_) {}

    bool run(const std::vector<uint32_t> &recs)
    {
        out.st = mirfold_stats{};
        out.st.n_devices = 1;
        CK(cudaSetDevice(D.id));
        const auto t0 = std::chrono::steady_clock::now();
        loci.reserve(recs.size());
        uint64_t total_cells = 0;
        for (uint32_t r : recs) {
            const uint64_t len = J.h_off[r + 1] - J.h_off[r];
            out.st.nt += len;
            if (len >= 5) {
                Locus l{r, (int)len, cells_of((int64_t)len, J.L)};
                loci.push_back(l);
                total_cells += l.cells;
            }
        }
        out.st.cells = total_cells;
        if (J.sink == SINK_CAND) {
            const CandArgs &c = *J.cand;
            const uint64_t nm = c.mature_off[c.nseq];
            CK(D.cand_regions.ensure(sizeof(mirfold_region) * (size_t)c.nseq + 16));
            CK(D.cand_matures.ensure(sizeof(mirfold_mature) * (size_t)nm + 16));
            CK(D.cand_moff.ensure(8 * ((size_t)c.nseq + 1)));
            CK(cudaMemcpy(D.cand_regions.p, c.regions, sizeof(mirfold_region) * (size_t)c.nseq, cudaMemcpyHostToDevice));
            if (nm) CK(cudaMemcpy(D.cand_matures.p, c.matures, sizeof(mirfold_mature) * (size_t)nm, cudaMemcpyHostToDevice));
            CK(cudaMemcpy(D.cand_moff.p, c.mature_off, 8 * ((size_t)c.nseq + 1), cudaMemcpyHostToDevice));
            out.st.h2d_bytes += sizeof(mirfold_region) * (uint64_t)c.nseq + sizeof(mirfold_mature) * nm + 8 * ((uint64_t)c.nseq + 1);
        }
        plan_chunks(total_cells);
        out.st.n_chunks = (int32_t)chunks.size();
        ht.mark("plan chunks");
        const int nlanes = nlanes_ = (J.serial || chunks.size() < 2) ? 1 : 2;
        bool ok = true;
        if (!chunks.empty()) {
            CK(cudaEventRecord(D.ev_first, D.lane[0].s_lo));
            ok = front(D.lane[0], 0);
            for (size_t k = 0; ok && k < chunks.size(); k++) {
                Lane &Ln = D.lane[k % nlanes];
                if (nlanes == 2 && k + 1 < chunks.size()) ok = front(D.lane[(k + 1) % 2], k + 1);   // its fill overlaps everything below
                ok = ok && middle(Ln) && back(Ln, k + 1 == chunks.size());
                if (nlanes == 1 && ok && k + 1 < chunks.size()) ok = front(Ln, k + 1);
                if (J.cb_abort && J.cb_abort->load()) { out.err = MIRFOLD_ERR_CALLBACK; out.errmsg = "chunk callback returned non-zero"; ok = false; }
            }
            for (int l = 0; ok && l < nlanes; l++) ok = retire(D.lane[l]);
        }
        if (!ok) {   // nothing may stay in flight into buffers the caller is about to drop
            cudaDeviceSynchronize();
            cudaGetLastError();
            D.lane[0].pending = D.lane[1].pending = false;
            D.lane[0].ovf.reset(); D.lane[1].ovf.reset();
            return false;
        }
        if (!chunks.empty()) {
            float ms = 0;
            cudaEventElapsedTime(&ms, D.ev_first, D.ev_last);
            out.st.ms_device = ms;
        }
        out.st.ms_total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        return true;
    }

    // chunks of about equal device bytes, each within a lane's budget.  A chunk boundary costs about 1.2 ms of device time
    // (launch train, two host syncs, wave tails the other lane only partly fills: 25 k loci in 1 / 2 / 4 / 9 chunks = 106.0 /
    // 108.3 / 108.0 / 111.5 ms end to end; 200 k loci in 7 / 16 / 37 chunks = 855.0 / 849.2 / 855.4) while a download runs at 55 GB/s, so a shard is only cut further than memory
    // demands when it is long enough for the last chunk's download (the only one not overlapped) to matter
    void plan_chunks(uint64_t total_cells)
    {
        if (loci.empty()) return;
        static const double cells_per_chunk = getenv("MIRFOLD_CHUNK_CELLS") ? atof(getenv("MIRFOLD_CHUNK_CELLS")) : 6.0e8;
        std::vector<size_t> lb(loci.size());
        size_t total = 0;
        int last_n = -1; size_t last_b = 0;
        for (size_t k = 0; k < loci.size(); k++) {
            if (loci[k].n != last_n) { last_n = loci[k].n; last_b = locus_bytes(last_n, J.L); }
            lb[k] = last_b; total += last_b;
        }
        const size_t lane_budget = std::max<size_t>(D.mem_budget / (J.serial ? 1 : 2), 1);
        size_t nch = (total + lane_budget - 1) / lane_budget;
        if (J.sink != SINK_NONE || nch > 1) nch = std::max<size_t>(nch, std::min<size_t>(64, (size_t)(total_cells / cells_per_chunk)));
        nch = std::max<size_t>(nch, 1);
        const size_t target = std::min(lane_budget, total / nch + 1);
        size_t b = 0, acc = 0;
        for (size_t k = 0; k < loci.size(); k++) {
            if (k > b && acc + lb[k] > target) { chunks.push_back({b, k}); b = k; acc = 0; }
            acc += lb[k];
        }
        chunks.push_back({b, loci.size()});
        // The download of a device's LAST chunk is the only one nothing overlaps (154 MB per device take 11.6 ms when eight GPUs
        // write into host memory at once): cut the last chunk 80 : 20 so that only the small tail's download is exposed.
        if (J.sink != SINK_NONE) {
            const Chunk last = chunks.back();
            uint64_t cells = 0;
            for (size_t k = last.begin; k < last.end; k++) cells += loci[k].cells;
            if (cells > 100000000ull && last.end - last.begin >= 64) {
                uint64_t acc2 = 0;
                size_t cut = last.begin;
                while (cut < last.end && acc2 < cells - cells / 5) acc2 += loci[cut++].cells;
                if (cut > last.begin && cut < last.end) {
                    chunks.back().end = cut;
                    chunks.push_back({cut, last.end});
                }
            }
        }
    }

    // ---- front: descriptors, upload, encode, band fill, f3, emission plan (everything up to the first host sync)
    bool front(Lane &Ln, size_t ci)
    {
        if (!retire(Ln)) return false;
        Prep &P = Ln.prep;
        P = Prep{};
        P.cb = chunks[ci].begin; P.ce = chunks[ci].end;
        const int nl = P.nl = (int)(P.ce - P.cb);
        CK(Ln.h_loci.ensure(sizeof(LocusDesc) * nl));
        CK(Ln.h_listoff.ensure(sizeof(unsigned long long) * nl));
        LocusDesc *hl = Ln.h_loci.as<LocusDesc>();
        unsigned long long *hlo = Ln.h_listoff.as<unsigned long long>();
        for (int k = 0; k < nl; k++) {
            const Locus &l = loci[P.cb + k];
            LocusDesc &d = hl[k];
            d = LocusDesc{};
            const unsigned long long be = shape_locus(d, l.n, J.L);
            d.rec = (int)l.rec;
            d.seq_off = P.seq_acc; d.band_off = P.band_acc; d.ring_off = 0;
            d.raw_off = d_raw ? raw_off[l.rec] : P.raw_acc;
            hlo[k] = P.list_acc;
            P.seq_acc += (unsigned long long)l.n + 3;
            P.band_acc += be;
            P.raw_acc += (unsigned long long)l.n;
            P.list_acc += (unsigned long long)l.n / 2 + 2;
            P.max_Ls = std::max(P.max_Ls, d.Ls);
        }
        while (P.n_long < nl && hl[P.n_long].n > MF_TILE_LEN) P.n_long++;   // sorted by descending cells == descending n
        std::vector<LocusDesc> units;
        P.ring_acc = build_fill_units(hl, nl, units, P.bucket_first, P.max_n);
        const int nu = P.nu = (int)units.size();
        CK(Ln.h_units.ensure(sizeof(LocusDesc) * (size_t)nu + 64));
        memcpy(Ln.h_units.p, units.data(), sizeof(LocusDesc) * (size_t)nu);
        out.st.fill_units += (uint64_t)nu;
        // ---- buffers (grow-only; cudaFree of an outgrown buffer synchronises the device, which also orders it after its last use)
        CK(Ln.loci.ensure(sizeof(LocusDesc) * nl));
        CK(Ln.listoff.ensure(sizeof(unsigned long long) * nl));
        CK(Ln.codes.ensure(P.seq_acc));
        CK(Ln.F.ensure(P.seq_acc * 4));
        CK(Ln.C.ensure(P.band_acc * 4));
        CK(Ln.M.ensure(P.band_acc * 4));
        CK(Ln.Ib.ensure(P.band_acc + 256));
        if (!J.force_wide) CK(Ln.Mp.ensure(P.band_acc * 4 + 4096));   // fML16 row pairs: NS (two copies) or NS/2 words per diagonal, by bucket
        CK(Ln.ring.ensure(P.ring_acc * 4));
        CK(Ln.tbcount.ensure((size_t)nl * 4));
        CK(Ln.tbbase.ensure((size_t)(nl + 1) * 8));
        CK(Ln.scan_in.ensure((size_t)(nl + 1) * 8));
        CK(Ln.startlist.ensure(P.list_acc * 4));
        CK(Ln.fail.ensure(4));
        CK(Ln.fillflags.ensure((size_t)nu * 4 + 4));
        CK(Ln.units.ensure(sizeof(LocusDesc) * (size_t)nu + 64));
        CK(Ln.h_small.ensure(256));
        // ---- upload
        cudaStream_t lo = Ln.s_lo, hi = Ln.s_hi;
        CK(cudaEventRecord(Ln.ev[0], lo));
        const char *raw_dev = d_raw;
        if (!d_raw) {
            CK(Ln.h_raw.ensure(P.raw_acc));
            char *hr = Ln.h_raw.as<char>();
            for (int k = 0; k < nl; k++) memcpy(hr + hl[k].raw_off, J.seqs + J.h_off[loci[P.cb + k].rec], (size_t)hl[k].n);
            CK(Ln.raw.ensure(P.raw_acc));
            CK(cudaMemcpyAsync(Ln.raw.p, hr, P.raw_acc, cudaMemcpyHostToDevice, lo));
            out.st.h2d_bytes += P.raw_acc;
            raw_dev = Ln.raw.as<char>();
        }
        CK(cudaMemcpyAsync(Ln.loci.p, hl, sizeof(LocusDesc) * nl, cudaMemcpyHostToDevice, lo));
        CK(cudaMemcpyAsync(Ln.listoff.p, hlo, sizeof(unsigned long long) * nl, cudaMemcpyHostToDevice, lo));
        CK(cudaMemcpyAsync(Ln.units.p, Ln.h_units.p, sizeof(LocusDesc) * (size_t)nu, cudaMemcpyHostToDevice, lo));
        out.st.h2d_bytes += (sizeof(LocusDesc) + 8) * (uint64_t)nl + sizeof(LocusDesc) * (uint64_t)nu;
        CK(cudaMemsetAsync(Ln.fail.p, 0, 4, lo));
        CK(cudaMemsetAsync(Ln.fillflags.p, 0, (size_t)nu * 4 + 4, lo));
        CK(cudaEventRecord(Ln.ev[1], lo));
        // The two lanes' fills normally share the SMs and finish together, which is what hides wave tails -- but it also means the
        // small last chunk (plan_chunks cuts it 80 : 20) would be done before the previous chunk's download even starts.  Its fill
        // therefore waits for the previous chunk's fill to END: it then runs next to that chunk's traceback and under its download.
        if (J.sink != SINK_NONE && nlanes_ == 2 && ci > 0 && ci + 1 == chunks.size())
            CK(cudaStreamWaitEvent(lo, D.lane[(ci + 1) % 2].ev[3], 0));
        // ---- K1, K2 on the low-priority stream
        const LocusDesc *dl = Ln.loci.as<LocusDesc>();
        CK(launch_prepare(raw_dev, dl, nl, P.seq_acc, Ln.codes.as<unsigned char>(), Ln.F.as<int>(), lo));
        CK(cudaEventRecord(Ln.ev[2], lo));
        FillLaunch fa{Ln.units.as<LocusDesc>(), nu, P.max_n, Ln.codes.as<unsigned char>(), Ln.C.as<int>(), Ln.M.as<int>(), Ln.ring.as<int>(),
                      Ln.Ib.as<unsigned char>(), Ln.Mp.as<unsigned int>(), D.dP, {0, 0, 0, 0, 0, 0}, Ln.fillflags.as<int>(), J.force_wide ? 1 : 0, env_opts() | (J.L >= MF_DYNW_MIN_SPAN ? 8 : 0)};
        for (int b = 0; b < 6; b++) fa.bucket_first[b] = P.bucket_first[b];
        static const bool no_side = getenv("MIRFOLD_NO_SIDE_STREAM") != nullptr;   // A/B runs
        CK(launch_fill(fa, lo, no_side ? nullptr : Ln.side, Ln.ev[10], Ln.ev[11]));
        CK(cudaEventRecord(Ln.ev[3], lo));
        // ---- K3 + emission plan on the high-priority stream
        CK(cudaStreamWaitEvent(hi, Ln.ev[3], 0));
        CK(launch_f3(dl, nl, P.n_long, P.max_Ls, Ln.codes.as<unsigned char>(), Ln.C.as<int>(), Ln.F.as<int>(), D.dP, hi));
        {   // kernels launched so far: k_prepare, the fill kernels (16-bit + 32-bit redo per non-empty bucket, generic), k_f3 / k_f3_cta
            int nfill = P.bucket_first[1] > P.bucket_first[0] ? 1 : 0;
            for (int b = 1; b < 5; b++) if (P.bucket_first[b + 1] > P.bucket_first[b]) nfill += J.force_wide ? 1 : 2;
            out.st.kernel_launches += 1 + nfill + (P.n_long > 0 ? 1 : 0) + (P.n_long < nl ? 1 : 0);
        }
        CK(cudaEventRecord(Ln.ev[4], hi));
        TraceBuffers &tb = Ln.tb;
        tb = TraceBuffers{};
        tb.loci = dl; tb.nloci = nl; tb.codes = Ln.codes.as<unsigned char>();
        tb.C = Ln.C.as<int>(); tb.M = Ln.M.as<int>(); tb.F = Ln.F.as<int>(); tb.Ib = Ln.Ib.as<unsigned char>(); tb.P = D.dP;
        tb.tb_count = Ln.tbcount.as<int>(); tb.tb_base = Ln.tbbase.as<unsigned long long>();
        tb.list_off = Ln.listoff.as<unsigned long long>(); tb.tb_start_list = Ln.startlist.as<int>();
        tb.fail_flag = Ln.fail.as<int>();
        CK(launch_plan(tb, hi));
        k_widen_counts<<<(nl + 1 + 255) / 256, 256, 0, hi>>>(tb.tb_count, Ln.scan_in.as<unsigned long long>(), nl);
        CK(cudaGetLastError());
        CK(exclusive_scan(Ln, Ln.scan_in.as<unsigned long long>(), tb.tb_base, (size_t)nl + 1, hi));
        CK(cudaMemcpyAsync(Ln.h_small.p, tb.tb_base + nl, 8, cudaMemcpyDeviceToHost, hi));
        out.st.kernel_launches += 4;   // k_plan, k_widen_counts, cub scan (init + scan)
        Ln.pending = true;
        ht.mark("front (descriptors, uploads, K1-K3 enqueue)");
        return true;
    }

    // ---- middle: tracebacks, emission decisions, output sizes (second host sync)
    bool middle(Lane &Ln)
    {
        cudaStream_t hi = Ln.s_hi;
        TraceBuffers &tb = Ln.tb;
        const Prep &P = Ln.prep;
        const int nl = P.nl;
        unsigned long long *hs = Ln.h_small.as<unsigned long long>();
        CK(cudaStreamSynchronize(hi));
        const unsigned long long ntb = hs[0];
        out.st.tracebacks += ntb;
        ht.mark("plan sync");
        tb.ntb = ntb;
        tb.slot_stride = (P.max_Ls + 4 + 3) & ~3;
        tb.stack_cap = P.max_Ls / 4 + 16;
        tb.code_win = (P.max_Ls + 8 + 15) & ~15;
        CK(Ln.slots.ensure((size_t)ntb * tb.slot_stride + 16));
        CK(Ln.tblen.ensure((size_t)ntb * 4 + 16)); CK(Ln.tbstart.ensure((size_t)ntb * 4 + 16));
        CK(Ln.tblocus.ensure((size_t)ntb * 4 + 16)); CK(Ln.tbflag.ensure((size_t)ntb * 4 + 16));
        CK(Ln.tbenergy.ensure((size_t)ntb * 4 + 16));
        CK(Ln.stackscr.ensure((size_t)ntb * tb.stack_cap * 8 + 16));
        CK(Ln.ssoff.ensure((size_t)(ntb + 1) * 8)); CK(Ln.hitidx.ensure((size_t)(ntb + 1) * 8));
        CK(Ln.scan_in.ensure((size_t)(ntb + 1) * 8 + (size_t)(nl + 1) * 8));
        CK(Ln.scan_out.ensure((size_t)(ntb + 1) * 8));
        tb.slots = Ln.slots.as<char>(); tb.tb_len = Ln.tblen.as<int>(); tb.tb_start = Ln.tbstart.as<int>();
        tb.tb_locus = Ln.tblocus.as<int>(); tb.tb_flag = Ln.tbflag.as<int>(); tb.tb_energy = Ln.tbenergy.as<int>();
        tb.stack_scratch = Ln.stackscr.as<int>();
        CK(launch_traceback(tb, hi));
        CK(launch_emit(tb, hi));
        if (ntb) {
            k_emit_sizes<<<(unsigned)((ntb + 1 + 255) / 256), 256, 0, hi>>>(tb.tb_flag, tb.tb_len, Ln.scan_in.as<unsigned long long>(),
                                                                             Ln.scan_out.as<unsigned long long>(), ntb);
            CK(cudaGetLastError());
            CK(exclusive_scan(Ln, Ln.scan_in.as<unsigned long long>(), Ln.ssoff.as<unsigned long long>(), (size_t)ntb + 1, hi));
            CK(exclusive_scan(Ln, Ln.scan_out.as<unsigned long long>(), Ln.hitidx.as<unsigned long long>(), (size_t)ntb + 1, hi));
            CK(cudaMemcpyAsync(&hs[1], Ln.ssoff.as<unsigned long long>() + ntb, 8, cudaMemcpyDeviceToHost, hi));
            CK(cudaMemcpyAsync(&hs[2], Ln.hitidx.as<unsigned long long>() + ntb, 8, cudaMemcpyDeviceToHost, hi));
            out.st.kernel_launches += 7;
        }
        CK(cudaMemcpyAsync(&hs[3], Ln.fail.p, 4, cudaMemcpyDeviceToHost, hi));
        CK(cudaStreamSynchronize(hi));
        if (*(int *)&hs[3]) { out.err = MIRFOLD_ERR_BACKTRACK; out.errmsg = "traceback found no decomposition"; return false; }
        Ln.abytes = ntb ? hs[1] : 0;
        Ln.nhits = ntb ? hs[2] : 0;
        ht.mark("traceback sync");
        return true;
    }

    // ---- back: pack the printed structures and start the download into the sink (asynchronous; see retire)
    bool back(Lane &Ln, bool last_chunk)
    {
        cudaStream_t hi = Ln.s_hi;
        TraceBuffers &tb = Ln.tb;
        const Prep &P = Ln.prep;
        const int nl = P.nl;
        const uint64_t nhits = Ln.nhits, abytes = Ln.abytes;
        mirfold_hit *dst_hits = nullptr;
        char *dst_arena = nullptr;
        Ln.hit_base = Ln.arena_base = 0;
        if (J.sink == SINK_SHARED) {
            if (J.shared->reserve(nhits, abytes, Ln.hit_base, Ln.arena_base)) {
                dst_hits = J.shared->hits.as<mirfold_hit>() + Ln.hit_base;
                dst_arena = J.shared->arena.as<char>() + Ln.arena_base;
            } else {
                Ln.ovf.reset(new OverflowChunk());
                CK(Ln.ovf->hits.ensure(nhits * sizeof(mirfold_hit) + 16));
                CK(Ln.ovf->arena.ensure(abytes + 16));
                Ln.ovf->nhits = nhits; Ln.ovf->abytes = abytes;
                dst_hits = Ln.ovf->hits.as<mirfold_hit>();
                dst_arena = Ln.ovf->arena.as<char>();
            }
        } else if (J.sink == SINK_STREAM) {
            CK(Ln.h_hits.ensure(nhits * sizeof(mirfold_hit) + 16));
            CK(Ln.h_arena.ensure(abytes + 16));
            dst_hits = Ln.h_hits.as<mirfold_hit>();
            dst_arena = Ln.h_arena.as<char>();
        }
        CK(Ln.o_hits.ensure(nhits * sizeof(mirfold_hit) + 16));
        CK(Ln.o_arena.ensure(abytes + 16));
        CK(Ln.bounds.ensure((size_t)(nl + 1) * 8));
        CK(Ln.totals.ensure((size_t)nl * 4 + 4));
        CK(launch_pack(tb, Ln.ssoff.as<unsigned long long>(), Ln.hitidx.as<unsigned long long>(), Ln.o_arena.as<char>(),
                       Ln.o_hits.as<mirfold_hit>(), Ln.arena_base, hi));
        out.st.kernel_launches += 1;
        if (J.sink == SINK_CAND) { if (!back_candidates(Ln, last_chunk)) return false; }
        else {
        CK(cudaEventRecord(Ln.ev[5], hi));
        if (last_chunk) CK(cudaEventRecord(D.ev_last, hi));
        }
        if (J.sink == SINK_CAND) {
        } else if (J.sink != SINK_NONE) {
            if (tb.ntb) {
                k_gather_bounds<<<(nl + 1 + 255) / 256, 256, 0, hi>>>(tb.tb_base, Ln.hitidx.as<unsigned long long>(), tb.loci, tb.F,
                                                                      Ln.bounds.as<unsigned long long>(), Ln.totals.as<int>(), nl);
                CK(cudaGetLastError());
                out.st.kernel_launches += 1;
            } else {
                CK(cudaMemsetAsync(Ln.bounds.p, 0, (size_t)(nl + 1) * 8, hi));
                CK(cudaMemsetAsync(Ln.totals.p, 0, (size_t)nl * 4, hi));
            }
            CK(Ln.h_out.ensure((size_t)(nl + 1) * 8 + (size_t)nl * 4 + 64));
            unsigned long long *h_bounds = Ln.h_out.as<unsigned long long>();
            int *h_total = (int *)(h_bounds + nl + 1);
            if (nhits) CK(cudaMemcpyAsync(dst_hits, Ln.o_hits.p, nhits * sizeof(mirfold_hit), cudaMemcpyDeviceToHost, hi));
            if (abytes) CK(cudaMemcpyAsync(dst_arena, Ln.o_arena.p, abytes, cudaMemcpyDeviceToHost, hi));
            CK(cudaMemcpyAsync(h_bounds, Ln.bounds.p, (size_t)(nl + 1) * 8, cudaMemcpyDeviceToHost, hi));
            CK(cudaMemcpyAsync(h_total, Ln.totals.p, (size_t)nl * 4, cudaMemcpyDeviceToHost, hi));
            out.st.d2h_bytes += nhits * sizeof(mirfold_hit) + (uint64_t)(nl + 1) * 8 + (uint64_t)nl * 4 + abytes + 32;
        } else out.st.d2h_bytes += 32;
        CK(cudaEventRecord(Ln.ev[6], hi));
        // the lane's next chunk (uploads on s_lo) may only start once this one has left the device
        CK(cudaStreamWaitEvent(Ln.s_lo, Ln.ev[6], 0));
        out.nhits += nhits;
        out.arena_bytes += abytes;
        ht.mark("back (pack + download enqueue)");
        return true;
    }

    // ---- fused stages 1 + 3 on the chunk's packed hits (still in HBM): classify, duplex verdicts, compact download
    bool back_candidates(Lane &Ln, bool last_chunk)
    {
        cudaStream_t hi = Ln.s_hi;
        TraceBuffers &tb = Ln.tb;
        const Prep &P = Ln.prep;
        const int nl = P.nl;
        const uint64_t nhits = Ln.nhits;
        const CandArgs &ca = *J.cand;
        unsigned long long *hs = Ln.h_small.as<unsigned long long>();
        k_gather_bounds<<<(nl + 1 + 255) / 256, 256, 0, hi>>>(tb.tb_base, Ln.hitidx.as<unsigned long long>(), tb.loci, tb.F,
                                                              Ln.bounds.as<unsigned long long>(), Ln.totals.as<int>(), nl);
        CK(cudaGetLastError());
        CK(Ln.c_counts.ensure((nhits + 2) * 8)); CK(Ln.c_soff.ensure((nhits + 2) * 8));
        CK(Ln.c_lsb.ensure((size_t)(nl + 1) * 8));
        CandLaunch c{};
        c.loci = tb.loci; c.nloci = nl; c.hits = Ln.o_hits.as<mirfold_hit>(); c.nhits = nhits;
        c.arena = Ln.o_arena.as<char>(); c.arena_base = Ln.arena_base; c.bounds = Ln.bounds.as<unsigned long long>();
        c.regions = D.cand_regions.as<mirfold_region>(); c.matures = D.cand_matures.as<mirfold_mature>();
        c.mature_off = D.cand_moff.as<unsigned long long>();
        c.minlen = ca.minlen; c.minloop = ca.minloop; c.min_mature = ca.min_mature; c.max_mature = ca.max_mature;
        c.counts = Ln.c_counts.as<unsigned long long>(); c.soff = Ln.c_soff.as<unsigned long long>();
        c.locus_sbegin = Ln.c_lsb.as<unsigned long long>();
        c.fail_flag = Ln.fail.as<int>();
        CK(launch_cand_classify(c, 0, hi));
        CK(exclusive_scan(Ln, c.counts, Ln.c_soff.as<unsigned long long>(), (size_t)nhits + 1, hi));
        CK(cudaMemcpyAsync(&hs[4], Ln.c_soff.as<unsigned long long>() + nhits, 8, cudaMemcpyDeviceToHost, hi));
        CK(cudaMemcpyAsync(&hs[3], Ln.fail.p, 4, cudaMemcpyDeviceToHost, hi));
        CK(cudaStreamSynchronize(hi));
        if (*(int *)&hs[3]) { out.err = MIRFOLD_ERR_ARG; out.errmsg = "a hit on which the reference's classifier raises (no pair / unbalanced)"; return false; }
        const uint64_t ns = Ln.c_ns = hs[4];
        CK(Ln.c_structs.ensure(ns * sizeof(mirfold_structure) + 16)); CK(Ln.c_sout.ensure(ns * sizeof(mirfold_structure) + 16));
        CK(Ln.c_nq.ensure((ns + 2) * 8)); CK(Ln.c_sbytes.ensure((ns + 2) * 8));
        CK(Ln.c_voff.ensure((ns + 2) * 8)); CK(Ln.c_aoff.ensure((ns + 2) * 8));
        CK(cudaMemsetAsync(Ln.c_nq.p, 0, (ns + 1) * 8, hi)); CK(cudaMemsetAsync(Ln.c_sbytes.p, 0, (ns + 1) * 8, hi));
        c.structs = Ln.c_structs.as<mirfold_structure>(); c.nstructs = ns;
        c.nq = Ln.c_nq.as<unsigned long long>(); c.sbytes = Ln.c_sbytes.as<unsigned long long>();
        CK(launch_cand_classify(c, 1, hi));
        CK(exclusive_scan(Ln, c.nq, Ln.c_voff.as<unsigned long long>(), (size_t)ns + 1, hi));
        CK(exclusive_scan(Ln, c.sbytes, Ln.c_aoff.as<unsigned long long>(), (size_t)ns + 1, hi));
        CK(cudaMemcpyAsync(&hs[5], Ln.c_voff.as<unsigned long long>() + ns, 8, cudaMemcpyDeviceToHost, hi));
        CK(cudaMemcpyAsync(&hs[6], Ln.c_aoff.as<unsigned long long>() + ns, 8, cudaMemcpyDeviceToHost, hi));
        CK(cudaStreamSynchronize(hi));
        const uint64_t nv = Ln.c_nv = hs[5], nb = Ln.c_nb = hs[6];
        CK(Ln.c_verd.ensure(nv * sizeof(mirfold_duplex_verdict) + 16)); CK(Ln.c_vmat.ensure(nv * 4 + 16)); CK(Ln.c_arena.ensure(nb + 16));
        c.voff = Ln.c_voff.as<unsigned long long>(); c.aoff = Ln.c_aoff.as<unsigned long long>();
        c.verdicts = Ln.c_verd.as<mirfold_duplex_verdict>(); c.verdict_mature = Ln.c_vmat.as<unsigned int>();
        c.out_arena = Ln.c_arena.as<char>(); c.out_base = 0; c.structs_out = Ln.c_sout.as<mirfold_structure>();
        CK(launch_cand_finish(c, P.max_Ls + 4, hi));
        out.st.kernel_launches += 11;   // gather, 2 x classify, locus bounds, 3 x cub scan (2 kernels each), duplex, strings
        CK(cudaEventRecord(Ln.ev[5], hi));
        if (last_chunk) CK(cudaEventRecord(D.ev_last, hi));
        CK(Ln.hc_structs.ensure(ns * sizeof(mirfold_structure) + 16)); CK(Ln.hc_arena.ensure(nb + 16));
        CK(Ln.hc_verd.ensure(nv * sizeof(mirfold_duplex_verdict) + 16)); CK(Ln.hc_vmat.ensure(nv * 4 + 16));
        CK(Ln.hc_voff.ensure((ns + 1) * 8)); CK(Ln.hc_lsb.ensure((size_t)(nl + 1) * 8));
        if (ns) CK(cudaMemcpyAsync(Ln.hc_structs.p, Ln.c_sout.p, ns * sizeof(mirfold_structure), cudaMemcpyDeviceToHost, hi));
        if (nb) CK(cudaMemcpyAsync(Ln.hc_arena.p, Ln.c_arena.p, nb, cudaMemcpyDeviceToHost, hi));
        if (nv) CK(cudaMemcpyAsync(Ln.hc_verd.p, Ln.c_verd.p, nv * sizeof(mirfold_duplex_verdict), cudaMemcpyDeviceToHost, hi));
        if (nv) CK(cudaMemcpyAsync(Ln.hc_vmat.p, Ln.c_vmat.p, nv * 4, cudaMemcpyDeviceToHost, hi));
        CK(cudaMemcpyAsync(Ln.hc_voff.p, Ln.c_voff.p, (ns + 1) * 8, cudaMemcpyDeviceToHost, hi));
        CK(cudaMemcpyAsync(Ln.hc_lsb.p, Ln.c_lsb.p, (size_t)(nl + 1) * 8, cudaMemcpyDeviceToHost, hi));
        out.st.d2h_bytes += ns * sizeof(mirfold_structure) + nb + nv * (sizeof(mirfold_duplex_verdict) + 4) + (ns + 1) * 8 + (uint64_t)(nl + 1) * 8 + 64;
        return true;
    }

    void retire_candidates(Lane &Ln)
    {
        const Prep &P = Ln.prep;
        const int nl = P.nl;
        CandCollector &C = *J.collect;
        const uint64_t ns = Ln.c_ns, nv = Ln.c_nv, nb = Ln.c_nb;
        const unsigned long long *lsb = Ln.hc_lsb.as<unsigned long long>(), *voff = Ln.hc_voff.as<unsigned long long>();
        std::lock_guard<std::mutex> lk(C.mu);
        const uint64_t bs = C.structs.size(), bv = C.verdicts.size(), ba = C.arena.size();
        C.structs.insert(C.structs.end(), Ln.hc_structs.as<mirfold_structure>(), Ln.hc_structs.as<mirfold_structure>() + ns);
        for (uint64_t k = bs; k < bs + ns; k++) C.structs[k].ss_off += ba;
        C.arena.insert(C.arena.end(), Ln.hc_arena.as<char>(), Ln.hc_arena.as<char>() + nb);
        C.verdicts.insert(C.verdicts.end(), Ln.hc_verd.as<mirfold_duplex_verdict>(), Ln.hc_verd.as<mirfold_duplex_verdict>() + nv);
        C.verdict_mature.insert(C.verdict_mature.end(), Ln.hc_vmat.as<uint32_t>(), Ln.hc_vmat.as<uint32_t>() + nv);
        for (uint64_t k = 0; k < ns; k++) { C.verdict_begin.push_back(bv + voff[k]); C.verdict_count.push_back((uint32_t)(voff[k + 1] - voff[k])); }
        for (int k = 0; k < nl; k++) {
            const uint32_t r = loci[P.cb + k].rec;
            C.struct_begin[r] = bs + lsb[k];
            C.struct_count[r] = (uint32_t)(lsb[k + 1] - lsb[k]);
        }
    }

    // ---- retire: wait for the lane's download, publish the chunk's per-record tables, collect stage times
    bool retire(Lane &Ln)
    {
        if (!Ln.pending) return true;
        Ln.pending = false;
        CK(cudaEventSynchronize(Ln.ev[6]));
        const Prep &P = Ln.prep;
        const int nl = P.nl;
        float ms = 0;
        cudaEventElapsedTime(&ms, Ln.ev[0], Ln.ev[1]); out.st.ms_h2d += ms;
        cudaEventElapsedTime(&ms, Ln.ev[2], Ln.ev[3]); out.st.ms_fill += ms;
        cudaEventElapsedTime(&ms, Ln.ev[3], Ln.ev[4]); out.st.ms_f3 += ms;
        cudaEventElapsedTime(&ms, Ln.ev[4], Ln.ev[5]); out.st.ms_trace += ms;
        cudaEventElapsedTime(&ms, Ln.ev[5], Ln.ev[6]); out.st.ms_d2h += ms;
        static const bool trace_events = getenv("MIRFOLD_TRACE_EVENTS") != nullptr;
        if (trace_events) {   // device timeline of the chunk relative to the call's first event (ms)
            float t[7];
            for (int k = 0; k < 7; k++) cudaEventElapsedTime(&t[k], D.ev_first, Ln.ev[k]);
            fprintf(stderr, "[mirfold events] dev %d lane %d loci %6d  h2d %.2f-%.2f  fill %.2f-%.2f  f3 -%.2f  pack -%.2f  d2h -%.2f\n", D.id,
                    (int)(&Ln - D.lane), nl, t[0], t[1], t[2], t[3], t[4], t[5], t[6]);
        }
        if (J.sink == SINK_NONE) return true;
        if (J.sink == SINK_CAND) { retire_candidates(Ln); return true; }
        const unsigned long long *h_bounds = Ln.h_out.as<unsigned long long>();
        const int *h_total = (const int *)(h_bounds + nl + 1);
        if (J.sink == SINK_SHARED) {
            SharedOut &S = *J.shared;
            if (!Ln.ovf) {
                for (int k = 0; k < nl; k++) {
                    const uint32_t r = loci[P.cb + k].rec;
                    S.hit_begin[r] = Ln.hit_base + h_bounds[k];
                    S.hit_count[r] = (uint32_t)(h_bounds[k + 1] - h_bounds[k]);
                    S.totals[r] = h_total[k];
                }
            } else {
                OverflowChunk &O = *Ln.ovf;
                O.recs.resize(nl); O.begin.resize(nl); O.count.resize(nl);
                for (int k = 0; k < nl; k++) {
                    const uint32_t r = loci[P.cb + k].rec;
                    O.recs[k] = r; O.begin[k] = h_bounds[k]; O.count[k] = (uint32_t)(h_bounds[k + 1] - h_bounds[k]);
                    S.totals[r] = h_total[k];
                }
                std::lock_guard<std::mutex> lk(S.mu);
                S.overflow.push_back(std::move(Ln.ovf));
            }
        } else {
            Ln.s_rec.resize(nl); Ln.s_begin.resize(nl); Ln.s_count.resize(nl); Ln.s_total.resize(nl);
            for (int k = 0; k < nl; k++) {
                Ln.s_rec[k] = loci[P.cb + k].rec;
                Ln.s_begin[k] = h_bounds[k];
                Ln.s_count[k] = (uint32_t)(h_bounds[k + 1] - h_bounds[k]);
                Ln.s_total[k] = h_total[k];
            }
            mirfold_chunk c{};
            c.n_records = (uint32_t)nl; c.device = D.id;
            c.record = Ln.s_rec.data(); c.hit_begin = Ln.s_begin.data(); c.hit_count = Ln.s_count.data();
            c.total_mfe_dcal = Ln.s_total.data();
            c.nhits = Ln.nhits; c.hits = Ln.h_hits.as<mirfold_hit>();
            c.ss_arena = Ln.h_arena.as<char>(); c.ss_bytes = Ln.abytes;
            std::lock_guard<std::mutex> lk(*J.cb_mu);
            if (!J.cb_abort->load() && J.fn(J.user, &c) != 0) J.cb_abort->store(1);
        }
        ht.mark("retire (download wait + record tables)");
        return true;
    }
};
#undef CK

bool run_device(Device &D, const Job &J, const std::vector<uint32_t> &recs, const char *d_raw, const uint64_t *raw_off, DevOut &out)
{
    DevicePipeline p(D, J, d_raw, raw_off, out);
    return p.run(recs);
}

HBuf pool_take(mirfold_ctx *ctx, std::vector<HBuf> &pool, size_t bytes)
{   // smallest pooled buffer that is large enough, else the largest one (it will be re-grown by the caller)
    std::lock_guard<std::mutex> lk(ctx->pool_mu);
    int best = -1;
    for (int k = 0; k < (int)pool.size(); k++)
        if (pool[k].cap >= bytes && (best < 0 || pool[k].cap < pool[best].cap)) best = k;
    HBuf b;
    if (best >= 0) { b = pool[best]; pool.erase(pool.begin() + best); }
    return b;
}
void pool_give(mirfold_ctx *ctx, std::vector<HBuf> &pool, HBuf &b)
{   // caller holds ctx->pool_mu
    if (!b.p) return;
    if (ctx->closed) { b.release(); return; }
    if (pool.size() >= 4) {
        int small = 0;
        for (int k = 1; k < (int)pool.size(); k++) if (pool[k].cap < pool[small].cap) small = k;
        if (pool[small].cap < b.cap) std::swap(pool[small], b);
        b.release();
        return;
    }
    pool.push_back(b);
    b = HBuf();
}

void add_stats(mirfold_stats &a, const mirfold_stats &b)
{
    a.ms_h2d = std::max(a.ms_h2d, b.ms_h2d); a.ms_fill = std::max(a.ms_fill, b.ms_fill);
    a.ms_f3 = std::max(a.ms_f3, b.ms_f3); a.ms_trace = std::max(a.ms_trace, b.ms_trace);
    a.ms_d2h = std::max(a.ms_d2h, b.ms_d2h); a.ms_device = std::max(a.ms_device, b.ms_device);
    a.nt += b.nt; a.cells += b.cells; a.tracebacks += b.tracebacks; a.kernel_launches += b.kernel_launches;
    a.h2d_bytes += b.h2d_bytes; a.d2h_bytes += b.d2h_bytes; a.n_chunks += b.n_chunks; a.fill_units += b.fill_units;
}

int validate_offsets(mirfold_ctx *ctx, const uint64_t *seq_off, uint32_t nseq)
{
    for (uint32_t r = 0; r < nseq; r++) {
        if (seq_off[r + 1] < seq_off[r]) { ctx->last_error = "seq_off is not non-decreasing"; return MIRFOLD_ERR_ARG; }
        if (seq_off[r + 1] - seq_off[r] > (uint64_t)0x3fffffff) { ctx->last_error = "a sequence is longer than 2^30 - 1 nt"; return MIRFOLD_ERR_ARG; }
    }
    return MIRFOLD_OK;
}

// the common engine behind mirfold_fold / mirfold_fold_stream / mirfold_fold_device / mirfold_batch_fold
struct FoldArgs {
    const char *seqs = nullptr;
    const uint64_t *seq_off = nullptr;
    uint32_t nseq = 0;
    int span_L = 0;
    uint32_t flags = 0;
    int sink = SINK_SHARED;
    mirfold_chunk_fn fn = nullptr;
    void *user = nullptr;
    // device-resident input: per device a buffer and per record its offset; `shard` then comes from the batch
    const std::vector<void *> *d_raw = nullptr;
    const uint64_t *raw_off = nullptr;
    const std::vector<std::vector<uint32_t>> *shard = nullptr;
    int n_devices = 0;   // 0 = all devices of the context
    const CandArgs *cand = nullptr;      // SINK_CAND
    CandCollector *collect = nullptr;
    uint64_t *nhits_out = nullptr;
};

int fold_engine(mirfold_ctx *ctx, const FoldArgs &A, mirfold_result **out, mirfold_stats *stats_out)
{
    static const bool env_wide = getenv("MIRFOLD_FORCE_WIDE") != nullptr;   // A/B runs of bench.py
    static const bool env_serial = getenv("MIRFOLD_SERIAL") != nullptr;
    if (!ctx || !A.seq_off || (!A.seqs && !A.d_raw && A.nseq)) return MIRFOLD_ERR_ARG;
    if (ctx->closed) return MIRFOLD_ERR_ARG;
    if (A.span_L < 5 || A.span_L > MF_MAX_SPAN) { ctx->last_error = "span_L out of range [5, 4096]"; return MIRFOLD_ERR_ARG; }
    if (out) *out = nullptr;
    int rc = validate_offsets(ctx, A.seq_off, A.nseq);
    if (rc != MIRFOLD_OK) return rc;
    const auto t0 = std::chrono::steady_clock::now();
    HostTimer ht;
    const int G = A.n_devices > 0 ? A.n_devices : (int)ctx->devs.size();
    const uint32_t nseq = A.nseq;
    std::vector<std::vector<uint32_t>> shard_local;
    std::vector<uint64_t> load;
    if (!A.shard) plan_shards(A.seq_off, nseq, A.span_L, G, shard_local, load);
    const std::vector<std::vector<uint32_t>> &shard = A.shard ? *A.shard : shard_local;
    ht.mark("shard plan");

    Job J;
    J.seqs = A.seqs; J.h_off = A.seq_off; J.L = A.span_L;
    J.force_wide = (A.flags & MIRFOLD_FLAG_WIDE) != 0 || env_wide || A.span_L > MF16_MAX_SPAN;
    J.serial = (A.flags & MIRFOLD_FLAG_SERIAL) != 0 || env_serial;
    J.sink = A.sink;
    std::mutex cb_mu;
    std::atomic<int> cb_abort{0};
    J.fn = A.fn; J.user = A.user; J.cb_mu = &cb_mu; J.cb_abort = &cb_abort;
    J.cand = A.cand; J.collect = A.collect;

    std::unique_ptr<ResultOwner> R;
    SharedOut S;
    const uint64_t nt_total = nseq ? A.seq_off[nseq] - A.seq_off[0] : 0;
    if (A.sink != SINK_STREAM && A.sink != SINK_CAND) {
        R.reset(new ResultOwner());
        R->ctx = ctx;
        R->hit_begin.assign((size_t)nseq + 1, 0);
        R->hit_count.assign((size_t)nseq + 1, 0);
        R->totals.assign((size_t)nseq + 1, 0);
    }
    if (A.sink == SINK_SHARED) {
        // capacity is an estimate (a chunk that does not fit takes the overflow path below); MIRFOLD_RESULT_CAP_SCALE
        // shrinks it so that tests can reach that path
        const char *cs = getenv("MIRFOLD_RESULT_CAP_SCALE");
        const double scale = cs ? atof(cs) : 1.25;
        S.hits_cap = (uint64_t)(ctx->hits_per_nt * scale * (double)nt_total) + (cs ? 1 : 4096);
        S.arena_cap = (uint64_t)(ctx->arena_per_nt * scale * (double)nt_total) + (cs ? 1 : 65536);
        S.hits = pool_take(ctx, ctx->hits_pool, S.hits_cap * sizeof(mirfold_hit));
        S.arena = pool_take(ctx, ctx->arena_pool, S.arena_cap);
        cudaSetDevice(ctx->devs[0].id);
        if (S.hits.ensure(S.hits_cap * sizeof(mirfold_hit)) != cudaSuccess || S.arena.ensure(S.arena_cap) != cudaSuccess) {
            cudaGetLastError();
            S.hits.release(); S.arena.release();
            ctx->last_error = "pinned result buffers";
            return MIRFOLD_ERR_NOMEM;
        }
        if (!cs) { S.hits_cap = S.hits.cap / sizeof(mirfold_hit); S.arena_cap = S.arena.cap; }
        S.hit_begin = R->hit_begin.data(); S.hit_count = R->hit_count.data(); S.totals = R->totals.data();
        J.shared = &S;
    }
    ht.mark("result buffers");

    std::vector<DevOut> parts(G);
    auto dev_raw = [&](int g) { return A.d_raw ? (const char *)(*A.d_raw)[g] : nullptr; };
    if (G == 1) run_device(ctx->devs[0], J, shard[0], dev_raw(0), A.raw_off, parts[0]);
    else {
        std::vector<std::thread> th;
        for (int g = 0; g < G; g++)
            th.emplace_back([&, g] { run_device(ctx->devs[g], J, shard[g], dev_raw(g), A.raw_off, parts[g]); });
        for (auto &t : th) t.join();
    }
    for (int g = 0; g < G; g++)
        if (parts[g].err != MIRFOLD_OK) {
            ctx->last_error = parts[g].errmsg;
            std::lock_guard<std::mutex> lk(ctx->pool_mu);
            pool_give(ctx, ctx->hits_pool, S.hits);
            pool_give(ctx, ctx->arena_pool, S.arena);
            return parts[g].err;
        }
    ht.mark("devices");
    mirfold_stats st{};
    uint64_t nhits = 0, abytes = 0;
    for (int g = 0; g < G; g++) { add_stats(st, parts[g].st); nhits += parts[g].nhits; abytes += parts[g].arena_bytes; }
    st.n_devices = G;

    if (A.sink == SINK_CAND) {
        st.ms_total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (stats_out) *stats_out = st;
        if (A.nhits_out) *A.nhits_out = nhits;
        return MIRFOLD_OK;
    }
    if (A.sink == SINK_STREAM) {
        // records without a DP band (shorter than 5 nt) were never part of a device chunk
        std::vector<uint32_t> rec;
        for (uint32_t r = 0; r < nseq; r++) if (A.seq_off[r + 1] - A.seq_off[r] < 5) rec.push_back(r);
        if (!rec.empty()) {
            std::vector<uint64_t> hb(rec.size(), 0);
            std::vector<uint32_t> hc(rec.size(), 0);
            std::vector<int32_t> tot(rec.size(), 0);
            mirfold_chunk c{};
            c.n_records = (uint32_t)rec.size(); c.device = -1;
            c.record = rec.data(); c.hit_begin = hb.data(); c.hit_count = hc.data(); c.total_mfe_dcal = tot.data();
            if (A.fn(A.user, &c) != 0) { ctx->last_error = "chunk callback returned non-zero"; return MIRFOLD_ERR_CALLBACK; }
        }
        st.ms_total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (stats_out) *stats_out = st;
        return MIRFOLD_OK;
    }

    if (A.sink == SINK_SHARED) {
        if (!S.overflow.empty()) {
            // the capacity estimate was too small: one exact-size copy, and a better estimate for the next call
            uint64_t th = S.hits_used, ta = S.arena_used;
            for (auto &o : S.overflow) { th += o->nhits; ta += o->abytes; }
            HBuf nh, na;
            if (nh.ensure(th * sizeof(mirfold_hit) + 16) != cudaSuccess || na.ensure(ta + 16) != cudaSuccess) {
                cudaGetLastError();
                nh.release(); na.release(); S.hits.release(); S.arena.release();
                return MIRFOLD_ERR_NOMEM;
            }
            memcpy(nh.p, S.hits.p, S.hits_used * sizeof(mirfold_hit));
            memcpy(na.p, S.arena.p, S.arena_used);
            uint64_t hb = S.hits_used, ab = S.arena_used;
            for (auto &o : S.overflow) {
                mirfold_hit *dst = nh.as<mirfold_hit>() + hb;
                memcpy(dst, o->hits.p, o->nhits * sizeof(mirfold_hit));
                for (uint64_t h = 0; h < o->nhits; h++) dst[h].ss_off += ab;
                memcpy(na.as<char>() + ab, o->arena.p, o->abytes);
                for (size_t k = 0; k < o->recs.size(); k++) { S.hit_begin[o->recs[k]] = hb + o->begin[k]; S.hit_count[o->recs[k]] = o->count[k]; }
                hb += o->nhits; ab += o->abytes;
                o->hits.release(); o->arena.release();
            }
            S.hits.release(); S.arena.release();
            S.hits = nh; S.arena = na;
            S.hits_used = th; S.arena_used = ta;
        }
        if (nt_total && !getenv("MIRFOLD_RESULT_CAP_SCALE")) {
            ctx->hits_per_nt = std::max(ctx->hits_per_nt, (double)S.hits_used / (double)nt_total);
            ctx->arena_per_nt = std::max(ctx->arena_per_nt, (double)S.arena_used / (double)nt_total);
        }
        R->hits = S.hits; S.hits = HBuf();
        R->arena = S.arena; S.arena = HBuf();
        R->pub.hits = R->hits.as<mirfold_hit>();
        R->pub.ss_arena = R->arena.as<char>();
        if (R->arena.cap > abytes) R->arena.as<char>()[abytes] = 0;
    } else {
        R->pub.hits = nullptr;
        R->pub.ss_arena = nullptr;
    }
    ht.mark("publish");
    st.ms_total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    R->pub.nseq = nseq;
    R->pub.nhits = nhits;
    R->pub.ss_bytes = abytes;
    R->pub.hit_begin = R->hit_begin.data();
    R->pub.hit_count = R->hit_count.data();
    R->pub.total_mfe_dcal = R->totals.data();
    R->pub.stats = st;
    if (stats_out) *stats_out = st;
    {
        std::lock_guard<std::mutex> lk(ctx->pool_mu);
        ctx->live_results++;
    }
    *out = &R.release()->pub;
    return MIRFOLD_OK;
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
    case MIRFOLD_ERR_CALLBACK: return "chunk callback returned non-zero";
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
    ctx->devs.resize(ids.size());
    size_t made = 0;
    for (size_t k = 0; k < ids.size() && rc == MIRFOLD_OK; k++) {
        Device &D = ctx->devs[k];
        D.id = ids[k];
        made = k + 1;
        // the same ordinal may be listed more than once (each entry is an independent pipeline sharing the GPU):
        // the memory budget is split between the entries
        int share = 0;
        for (int id : ids) share += id == D.id;
        int prio_lo = 0, prio_hi = 0;
        cudaError_t e = cudaSetDevice(D.id);
        if (e == cudaSuccess) e = cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        for (int l = 0; l < 2; l++) if (e == cudaSuccess) e = D.lane[l].create(prio_hi);
        if (e == cudaSuccess) e = cudaEventCreate(&D.ev_first);
        if (e == cudaSuccess) e = cudaEventCreate(&D.ev_last);
        if (e == cudaSuccess) e = cudaMalloc(&D.dP, sizeof(DevParams));
        if (e == cudaSuccess) e = cudaMemcpy(D.dP, hp, sizeof(DevParams), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = fill_configure_device();
        size_t fr = 0, tot = 0;
        if (e == cudaSuccess) e = cudaMemGetInfo(&fr, &tot);
        if (e != cudaSuccess) { ctx->last_error = cudaGetErrorString(e); cudaGetLastError(); rc = MIRFOLD_ERR_CUDA; break; }
        const char *env = getenv("MIRFOLD_MEM_BUDGET_MB");
        D.mem_budget = env ? (size_t)atoll(env) << 20 : std::min<size_t>((size_t)(fr * 0.60), (size_t)96 << 30) / (size_t)share;
    }
    delete hp;
    if (rc != MIRFOLD_OK) {
        for (size_t k = 0; k < made; k++) { cudaSetDevice(ctx->devs[k].id); ctx->devs[k].release(); }
        delete ctx;
        return rc;
    }
    *pctx = ctx;
    return MIRFOLD_OK;
}

void mirfold_close(mirfold_ctx *ctx)
{
    if (!ctx) return;
    bool del;
    {
        std::lock_guard<std::mutex> lk(ctx->pool_mu);
        if (ctx->closed) return;
        ctx->closed = true;
        for (auto &D : ctx->devs) { cudaSetDevice(D.id); cudaDeviceSynchronize(); D.release(); }
        ctx->devs.clear();
        for (auto &b : ctx->arena_pool) b.release();
        for (auto &b : ctx->hits_pool) b.release();
        ctx->arena_pool.clear(); ctx->hits_pool.clear();
        del = ctx->live_results == 0;
    }
    if (del) delete ctx;   // otherwise the last mirfold_free_result() deletes it
}

int mirfold_fold(mirfold_ctx *ctx, const char *seqs, const uint64_t *seq_off, uint32_t nseq, int span_L, uint32_t flags,
                 mirfold_result **out)
{
    if (!out) return MIRFOLD_ERR_ARG;
    FoldArgs A;
    A.seqs = seqs; A.seq_off = seq_off; A.nseq = nseq; A.span_L = span_L; A.flags = flags; A.sink = SINK_SHARED;
    return fold_engine(ctx, A, out, nullptr);
}

int mirfold_fold_stream(mirfold_ctx *ctx, const char *seqs, const uint64_t *seq_off, uint32_t nseq, int span_L, uint32_t flags,
                        mirfold_chunk_fn fn, void *user, mirfold_stats *stats)
{
    if (!fn) return MIRFOLD_ERR_ARG;
    FoldArgs A;
    A.seqs = seqs; A.seq_off = seq_off; A.nseq = nseq; A.span_L = span_L; A.flags = flags; A.sink = SINK_STREAM;
    A.fn = fn; A.user = user;
    return fold_engine(ctx, A, nullptr, stats);
}

int mirfold_fold_device(mirfold_ctx *ctx, const void *d_seqs, const void *d_seq_off, const uint64_t *h_seq_off,
                        uint32_t nseq, int span_L, uint32_t flags, void *stream, mirfold_result **out)
{
    (void)d_seq_off;
    if (!d_seqs || !out || !ctx || ctx->closed) return MIRFOLD_ERR_ARG;
    cudaSetDevice(ctx->devs[0].id);
    if (cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) { cudaGetLastError(); return MIRFOLD_ERR_CUDA; }   // the caller's producer
    std::vector<void *> d_raw(1, const_cast<void *>(d_seqs));
    FoldArgs A;
    A.seq_off = h_seq_off; A.nseq = nseq; A.span_L = span_L; A.flags = flags; A.sink = SINK_NONE;
    A.d_raw = &d_raw; A.raw_off = h_seq_off; A.n_devices = 1;
    return fold_engine(ctx, A, out, nullptr);
}

int mirfold_batch_upload(mirfold_ctx *ctx, const char *seqs, const uint64_t *seq_off, uint32_t nseq, int span_L, mirfold_batch **out)
{
    if (!ctx || !out || !seq_off || (!seqs && nseq) || ctx->closed) return MIRFOLD_ERR_ARG;
    *out = nullptr;
    if (span_L < 5 || span_L > MF_MAX_SPAN) { ctx->last_error = "span_L out of range [5, 4096]"; return MIRFOLD_ERR_ARG; }
    int rc = validate_offsets(ctx, seq_off, nseq);
    if (rc != MIRFOLD_OK) return rc;
    std::unique_ptr<mirfold_batch> B(new mirfold_batch());
    B->ctx = ctx; B->nseq = nseq; B->span_L = span_L;
    B->h_off.assign(seq_off, seq_off + (size_t)nseq + 1);
    const int G = (int)ctx->devs.size();
    std::vector<uint64_t> load;
    plan_shards(seq_off, nseq, span_L, G, B->shard, load);
    B->raw_off.assign((size_t)nseq + 1, 0);
    B->d_raw.assign((size_t)G, nullptr);
    for (int g = 0; g < G; g++) {
        uint64_t bytes = 0;
        for (uint32_t r : B->shard[g]) { B->raw_off[r] = bytes; bytes += seq_off[r + 1] - seq_off[r]; }
        std::vector<char> stage((size_t)bytes + 1);
        for (uint32_t r : B->shard[g]) memcpy(stage.data() + B->raw_off[r], seqs + seq_off[r], (size_t)(seq_off[r + 1] - seq_off[r]));
        cudaError_t e = cudaSetDevice(ctx->devs[g].id);
        if (e == cudaSuccess) e = cudaMalloc(&B->d_raw[g], (size_t)bytes + 16);
        if (e == cudaSuccess) e = cudaMemcpy(B->d_raw[g], stage.data(), (size_t)bytes, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            ctx->last_error = cudaGetErrorString(e);
            cudaGetLastError();
            mirfold_batch_free(B.release());
            return e == cudaErrorMemoryAllocation ? MIRFOLD_ERR_NOMEM : MIRFOLD_ERR_CUDA;
        }
    }
    *out = B.release();
    return MIRFOLD_OK;
}

int mirfold_batch_fold(mirfold_ctx *ctx, mirfold_batch *B, uint32_t flags, int download, mirfold_result **out)
{
    if (!ctx || !B || B->ctx != ctx || !out) return MIRFOLD_ERR_ARG;
    FoldArgs A;
    A.seq_off = B->h_off.data(); A.nseq = B->nseq; A.span_L = B->span_L; A.flags = flags;
    A.sink = download ? SINK_SHARED : SINK_NONE;
    A.d_raw = &B->d_raw; A.raw_off = B->raw_off.data(); A.shard = &B->shard;
    return fold_engine(ctx, A, out, nullptr);
}

void mirfold_batch_free(mirfold_batch *B)
{
    if (!B) return;
    if (B->ctx && !B->ctx->closed)
        for (size_t g = 0; g < B->d_raw.size() && g < B->ctx->devs.size(); g++)
            if (B->d_raw[g]) { cudaSetDevice(B->ctx->devs[g].id); cudaFree(B->d_raw[g]); }
    delete B;
}

namespace {
struct CandOwner {
    mirfold_candidates pub;
    CandCollector c;
};
}  // namespace

int mirfold_fold_candidates(mirfold_ctx *ctx, const char *seqs, const uint64_t *seq_off, uint32_t nseq, int span_L, uint32_t flags,
                            const mirfold_region *regions, const mirfold_mature *matures, const uint64_t *mature_off, int minlen,
                            int minloop, int min_mature_len, int max_mature_len, mirfold_candidates **out)
{
    if (!ctx || !out || !seq_off || (!regions && nseq) || !mature_off || (!matures && mature_off[nseq])) return MIRFOLD_ERR_ARG;
    *out = nullptr;
    for (uint32_t r = 0; r < nseq; r++) if (mature_off[r + 1] < mature_off[r]) { ctx->last_error = "mature_off is not non-decreasing"; return MIRFOLD_ERR_ARG; }
    std::unique_ptr<CandOwner> O(new CandOwner());
    O->c.struct_begin.assign((size_t)nseq + 1, 0);
    O->c.struct_count.assign((size_t)nseq + 1, 0);
    CandArgs ca;
    ca.regions = regions; ca.matures = matures; ca.mature_off = mature_off; ca.nseq = nseq;
    ca.minlen = minlen; ca.minloop = minloop; ca.min_mature = min_mature_len; ca.max_mature = max_mature_len;
    FoldArgs A;
    A.seqs = seqs; A.seq_off = seq_off; A.nseq = nseq; A.span_L = span_L; A.flags = flags; A.sink = SINK_CAND;
    A.cand = &ca; A.collect = &O->c;
    uint64_t nhits = 0;
    A.nhits_out = &nhits;
    mirfold_stats st{};
    const int rc = fold_engine(ctx, A, nullptr, &st);
    if (rc != MIRFOLD_OK) return rc;
    CandCollector &c = O->c;
    mirfold_candidates &p = O->pub;
    p = mirfold_candidates{};
    p.nseq = nseq;
    p.nstructs = c.structs.size();
    p.struct_begin = c.struct_begin.data(); p.struct_count = c.struct_count.data();
    p.structs = c.structs.data();
    c.arena.push_back(0);
    p.ss_arena = c.arena.data(); p.ss_bytes = c.arena.size() - 1;
    p.nverdicts = c.verdicts.size();
    p.verdict_begin = c.verdict_begin.data(); p.verdict_count = c.verdict_count.data();
    p.verdicts = c.verdicts.data(); p.verdict_mature = c.verdict_mature.data();
    p.nhits = nhits;
    p.stats = st;
    *out = &O.release()->pub;
    return MIRFOLD_OK;
}

void mirfold_free_candidates(mirfold_candidates *c)
{
    if (c) delete reinterpret_cast<CandOwner *>(c);
}

int mirfold_plan_shards(const uint64_t *seq_off, uint32_t nseq, int span_L, int n_shards, uint32_t *shard_of, uint64_t *shard_cells)
{
    if (!seq_off || n_shards < 1 || span_L < 5) return MIRFOLD_ERR_ARG;
    for (uint32_t r = 0; r < nseq; r++) if (seq_off[r + 1] < seq_off[r]) return MIRFOLD_ERR_ARG;
    std::vector<std::vector<uint32_t>> shard;
    std::vector<uint64_t> load;
    plan_shards(seq_off, nseq, span_L, n_shards, shard, load);
    for (int g = 0; g < n_shards; g++) {
        if (shard_of) for (uint32_t r : shard[g]) shard_of[r] = (uint32_t)g;
        if (shard_cells) shard_cells[g] = load[g] - shard[g].size();   // the plan weighs every record cells + 1
    }
    return MIRFOLD_OK;
}

int mirfold_plan_fill_units(uint32_t n, int span_L, mirfold_fill_plan *out)
{
    if (!out || n < 1 || n > 0x7fffffffu || span_L < 5) return MIRFOLD_ERR_ARG;
    LocusDesc d{};
    const unsigned long long be = shape_locus(d, (int)n, span_L);
    out->kernel = is_bucket_stride(d.stride) ? d.stride : 0;
    out->stride = d.stride;
    out->tile_len = d.tile_last ? d.stride : (int)n;
    out->tile_step = d.tile_last ? d.tile_step : (int)n;
    out->n_units = d.dmax >= 4 ? d.tile_last + 1 : 0;   // n < 5: nothing to fill
    out->dmax = d.dmax;
    out->band_cells = be;
    return MIRFOLD_OK;
}

void mirfold_free_result(mirfold_result *res)
{
    if (!res) return;
    ResultOwner *R = reinterpret_cast<ResultOwner *>(res);
    mirfold_ctx *ctx = R->ctx;
    bool del = false;
    {
        std::lock_guard<std::mutex> lk(ctx->pool_mu);
        pool_give(ctx, ctx->arena_pool, R->arena);
        pool_give(ctx, ctx->hits_pool, R->hits);
        ctx->live_results--;
        del = ctx->closed && ctx->live_results == 0;
    }
    R->arena.release();
    R->hits.release();
    delete R;
    if (del) delete ctx;
}

int mirfold_debug_matrices(mirfold_ctx *ctx, const char *seq, uint32_t n, int span_L, uint32_t flags, int32_t *c, int32_t *m,
                           int32_t *f3)
{
    if (!ctx || ctx->closed || !seq || n < 5 || !c || !m || !f3) return MIRFOLD_ERR_ARG;
    Device &D = ctx->devs[0];
    Lane &Ln = D.lane[0];
#define CK(call)                                                                  \
    do {                                                                          \
        cudaError_t e_ = (call);                                                  \
        if (e_ != cudaSuccess) { ctx->last_error = cudaGetErrorString(e_); cudaGetLastError(); return MIRFOLD_ERR_CUDA; } \
    } while (0)
    CK(cudaSetDevice(D.id));
    cudaStream_t st = Ln.s_lo;
    LocusDesc d{};
    const unsigned long long cells = shape_locus(d, (int)n, span_L);
    std::vector<LocusDesc> units;
    int bucket_first[6], max_n = 0;
    const unsigned long long ring_elems = build_fill_units(&d, 1, units, bucket_first, max_n);
    const int nu = (int)units.size();
    CK(Ln.raw.ensure(n)); CK(Ln.loci.ensure(sizeof d)); CK(Ln.codes.ensure(n + 3)); CK(Ln.F.ensure((n + 3) * 4));
    CK(Ln.C.ensure(cells * 4)); CK(Ln.M.ensure(cells * 4)); CK(Ln.Mp.ensure(cells * 4 + 4096)); CK(Ln.Ib.ensure(cells + 256));
    CK(Ln.ring.ensure((size_t)ring_elems * 4));
    CK(Ln.units.ensure(sizeof(LocusDesc) * (size_t)nu + 64));
    CK(cudaMemcpyAsync(Ln.raw.p, seq, n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(Ln.loci.p, &d, sizeof d, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(Ln.units.p, units.data(), sizeof(LocusDesc) * (size_t)nu, cudaMemcpyHostToDevice, st));
    const LocusDesc *dl = Ln.loci.as<LocusDesc>();
    CK(launch_prepare(Ln.raw.as<char>(), dl, 1, n + 3, Ln.codes.as<unsigned char>(), Ln.F.as<int>(), st));
    CK(Ln.fillflags.ensure((size_t)nu * 4 + 4));
    CK(cudaMemsetAsync(Ln.fillflags.p, 0, (size_t)nu * 4 + 4, st));
    FillLaunch fa{Ln.units.as<LocusDesc>(), nu, max_n, Ln.codes.as<unsigned char>(), Ln.C.as<int>(), Ln.M.as<int>(), Ln.ring.as<int>(),
                  Ln.Ib.as<unsigned char>(), Ln.Mp.as<unsigned int>(), D.dP, {0, 0, 0, 0, 0, 0}, Ln.fillflags.as<int>(),
                  ((flags & MIRFOLD_FLAG_WIDE) || span_L > MF16_MAX_SPAN) ? 1 : 0, env_opts() | (span_L >= MF_DYNW_MIN_SPAN ? 8 : 0)};
    for (int b = 0; b < 6; b++) fa.bucket_first[b] = bucket_first[b];
    CK(launch_fill(fa, st));
    CK(launch_f3(dl, 1, d.n > MF_TILE_LEN ? 1 : 0, d.Ls, Ln.codes.as<unsigned char>(), Ln.C.as<int>(), Ln.F.as<int>(), D.dP, st));
    std::vector<int> hc(cells), hm(cells), hf(n + 3);
    CK(cudaMemcpyAsync(hc.data(), Ln.C.p, cells * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(hm.data(), Ln.M.p, cells * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(hf.data(), Ln.F.p, (n + 3) * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const int W = d.Ls + 6;
    for (size_t k = 0; k < (size_t)(n + 2) * W; k++) c[k] = m[k] = MF_INF;
    for (int dd = 4; dd <= d.dmax; dd++)
        for (int i = 1; i <= (int)n - dd; i++) {
            // owner tile of row i (same mapping as band_row_base on the device)
            unsigned long long base = (unsigned long long)(i - 1);
            if (d.tile_last) {
                const int t = std::min((i - 1) / d.tile_step, d.tile_last);
                const int a = std::min(t * d.tile_step, (int)n - d.stride);
                base = (unsigned long long)t * band_elems(d.stride, d.dmax) + (unsigned long long)(i - 1 - a);
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
    if (!ctx || ctx->closed || !addmin_terms_per_s || !dpx_terms_per_s) return MIRFOLD_ERR_ARG;
    Device &D = ctx->devs[0];
    cudaSetDevice(D.id);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, D.id);
    cudaError_t e = run_int_peak(D.lane[0].s_lo, sms, addmin_terms_per_s, dpx_terms_per_s, nullptr);
    if (e != cudaSuccess) { ctx->last_error = cudaGetErrorString(e); return MIRFOLD_ERR_CUDA; }
    return MIRFOLD_OK;
}

int mirfold_int_peak2(mirfold_ctx *ctx, double *addmin_terms_per_s, double *dpx_terms_per_s, double *s16x2_terms_per_s)
{
    if (!ctx || ctx->closed || !addmin_terms_per_s || !dpx_terms_per_s || !s16x2_terms_per_s) return MIRFOLD_ERR_ARG;
    Device &D = ctx->devs[0];
    cudaSetDevice(D.id);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, D.id);
    cudaError_t e = run_int_peak(D.lane[0].s_lo, sms, addmin_terms_per_s, dpx_terms_per_s, s16x2_terms_per_s);
    if (e != cudaSuccess) { ctx->last_error = cudaGetErrorString(e); return MIRFOLD_ERR_CUDA; }
    return MIRFOLD_OK;
}

int mirfold_duplex(mirfold_ctx *ctx, const char *ss_arena, uint64_t ss_bytes, const mirfold_duplex_query *queries,
                   uint64_t nq, mirfold_duplex_verdict *verdicts)
{
    if (!ctx || ctx->closed || (!ss_arena && ss_bytes) || (!queries && nq) || (!verdicts && nq)) return MIRFOLD_ERR_ARG;
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
    cudaStream_t st = D.lane[0].s_hi;
    CK(cudaSetDevice(D.id));
    CK(D.duplex_arena.ensure(ss_bytes + 16));
    CK(D.duplex_q.ensure(nq * sizeof(mirfold_duplex_query)));
    CK(D.duplex_v.ensure(nq * sizeof(mirfold_duplex_verdict)));
    CK(cudaMemcpyAsync(D.duplex_arena.p, ss_arena, ss_bytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(D.duplex_q.p, queries, nq * sizeof(mirfold_duplex_query), cudaMemcpyHostToDevice, st));
    CK(launch_duplex(D.duplex_arena.as<char>(), D.duplex_q.as<mirfold_duplex_query>(), nq, D.duplex_v.as<mirfold_duplex_verdict>(),
                     (maxlen + 7) & ~7, st));
    CK(cudaMemcpyAsync(verdicts, D.duplex_v.p, nq * sizeof(mirfold_duplex_verdict), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return MIRFOLD_OK;
}
#undef CK

}  // extern "C"
