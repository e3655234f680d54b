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

namespace {

// Formatter of RNALfold's per-record output.  sizes(): exact byte count of record r's block; fill(): writes it.
struct RecordFormatter {
    const mirfold_result *res;
    const char *seqs;
    const uint64_t *seq_off;
    // every "(%6.2f)" field is 8 characters for |E| < 1000 kcal/mol and grows with the integer part beyond;
    // sizes are computed exactly with the same snprintf calls that fill the buffer
    size_t hit_line(char *dst, const mirfold_hit &h) const
    {   // dot-bracket, then " (%6.2f) %4d\n"
        if (dst) memcpy(dst, res->ss_arena + h.ss_off, (size_t)h.len);
        char tail[64];
        const int k = snprintf(tail, sizeof tail, " (%6.2f) %4d\n", h.mfe_dcal / 100., h.start);
        if (dst) memcpy(dst + h.len, tail, (size_t)k);
        return (size_t)h.len + (size_t)k;
    }
    size_t total_line(char *dst, uint32_t r) const
    {   // the sequence token upper-cased with T -> U (RNALfold main()), then " (%6.2f)\n"
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
    }
    size_t block(char *dst, uint32_t r) const
    {
        size_t b = 0;
        for (uint64_t h = res->hit_begin[r]; h < res->hit_begin[r] + res->hit_count[r]; h++) b += hit_line(dst ? dst + b : nullptr, res->hits[h]);
        return b + total_line(dst ? dst + b : nullptr, r);
    }
};

template <class Fn>
void for_ranges(uint64_t n, Fn &&fn)
{
    const unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    const unsigned nthr = n < 256 ? 1u : hw;
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nthr; t++) {
        const uint64_t lo = n * t / nthr, hi = n * (t + 1) / nthr;
        if (nthr == 1) fn(lo, hi);
        else th.emplace_back(fn, lo, hi);
    }
    for (auto &x : th) x.join();
}

inline bool c_space(char ch) { return ch == ' ' || ch == '\t' || ch == '\r' || ch == '\v' || ch == '\f' || ch == '\n'; }   // isspace, "C" locale

}  // namespace

extern "C" {

int mirfold_format_records(const mirfold_result *res, const char *seqs, const uint64_t *seq_off, uint32_t nseq, char **text,
                           uint64_t **rec_off)
{
    if (!res || !text || !rec_off || !seq_off || (!seqs && nseq) || res->nseq != nseq) return MIRFOLD_ERR_ARG;
    if (res->nhits && !res->ss_arena) return MIRFOLD_ERR_ARG;   // device-resident results carry no structures
    *text = nullptr; *rec_off = nullptr;
    uint64_t *off = (uint64_t *)malloc(sizeof(uint64_t) * ((size_t)nseq + 1));
    if (!off) return MIRFOLD_ERR_NOMEM;
    const RecordFormatter F{res, seqs, seq_off};
    for_ranges(nseq, [&](uint64_t lo, uint64_t hi) { for (uint64_t r = lo; r < hi; r++) off[r + 1] = F.block(nullptr, (uint32_t)r); });
    off[0] = 0;
    for (uint32_t r = 0; r < nseq; r++) off[r + 1] += off[r];
    char *buf = (char *)malloc((size_t)off[nseq] + 1);
    if (!buf) { free(off); return MIRFOLD_ERR_NOMEM; }
    for_ranges(nseq, [&](uint64_t lo, uint64_t hi) { for (uint64_t r = lo; r < hi; r++) F.block(buf + off[r], (uint32_t)r); });
    buf[off[nseq]] = 0;
    *text = buf; *rec_off = off;
    return MIRFOLD_OK;
}

// One `RNALfold -L span_L` run, stdin text in, stdout text out (RNALfold main()): lines starting with '>' or '*'
// and empty lines are echoed, a line "@" ends the input, the first whitespace-delimited token of any other line is folded.
int mirfold_fold_text(mirfold_ctx *ctx, const char *text, uint64_t len, int span_L, uint32_t flags, char **out, uint64_t *out_len)
{
    if (!ctx || (!text && len) || !out || !out_len) return MIRFOLD_ERR_ARG;
    *out = nullptr; *out_len = 0;
    struct Item { uint64_t b, e; uint32_t rec; };    // rec == UINT32_MAX: echo line [b, e); else sequence token [b, e) = record rec
    std::vector<Item> items;
    std::vector<uint64_t> seq_off(1, 0);
    uint64_t pos = 0;
    while (pos < len) {
        const char *nl = (const char *)memchr(text + pos, '\n', (size_t)(len - pos));
        const uint64_t end = nl ? (uint64_t)(nl - text) : len;
        if (end == pos || text[pos] == '>' || text[pos] == '*') items.push_back({pos, end, UINT32_MAX});
        else if (end == pos + 1 && text[pos] == '@') break;
        else {
            uint64_t b = pos;
            while (b < end && c_space(text[b])) b++;
            uint64_t e = b;
            while (e < end && !c_space(text[e])) e++;
            if (seq_off.size() - 1 >= 0xfffffffeull) return MIRFOLD_ERR_ARG;
            items.push_back({b, e, (uint32_t)(seq_off.size() - 1)});
            seq_off.push_back(seq_off.back() + (e - b));
        }
        pos = end + 1;
    }
    const uint32_t nseq = (uint32_t)(seq_off.size() - 1);
    char *seqs = (char *)malloc((size_t)seq_off[nseq] + 1);
    if (!seqs) return MIRFOLD_ERR_NOMEM;
    for_ranges(items.size(), [&](uint64_t lo, uint64_t hi) {
        for (uint64_t k = lo; k < hi; k++)
            if (items[k].rec != UINT32_MAX) memcpy(seqs + seq_off[items[k].rec], text + items[k].b, (size_t)(items[k].e - items[k].b));
    });
    mirfold_result *res = nullptr;
    int rc = mirfold_fold(ctx, seqs, seq_off.data(), nseq, span_L, flags, &res);
    if (rc != MIRFOLD_OK) { free(seqs); return rc; }
    const RecordFormatter F{res, seqs, seq_off.data()};
    std::vector<uint64_t> off(items.size() + 1, 0);
    for_ranges(items.size(), [&](uint64_t lo, uint64_t hi) {
        for (uint64_t k = lo; k < hi; k++)
            off[k + 1] = items[k].rec == UINT32_MAX ? items[k].e - items[k].b + 1 : F.block(nullptr, items[k].rec);
    });
    for (size_t k = 0; k < items.size(); k++) off[k + 1] += off[k];
    char *buf = (char *)malloc((size_t)off[items.size()] + 1);
    if (!buf) { mirfold_free_result(res); free(seqs); return MIRFOLD_ERR_NOMEM; }
    for_ranges(items.size(), [&](uint64_t lo, uint64_t hi) {
        for (uint64_t k = lo; k < hi; k++) {
            char *dst = buf + off[k];
            if (items[k].rec == UINT32_MAX) {
                memcpy(dst, text + items[k].b, (size_t)(items[k].e - items[k].b));
                dst[items[k].e - items[k].b] = '\n';
            } else F.block(dst, items[k].rec);
        }
    });
    buf[off[items.size()]] = 0;
    mirfold_free_result(res);
    free(seqs);
    *out = buf; *out_len = off[items.size()];
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
