// text.cu -- host-only text boundary of libmirfold (no device code): RNALfold's per-record output format and
// the names of the duplex verdict codes.
//
// Replaces what RNALfold's main() prints per record (RLF .rodata "%s (%6.2f) %4d\n" / "%s\n (%6.2f)\n",
// SURVEY.md A.6) -- the text /root/reference/miR_PREFeR.py collects at :3085-3098 and parses at :1541-1599.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/mirfold.h"

extern "C" {

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
