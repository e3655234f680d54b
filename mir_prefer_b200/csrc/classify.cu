// classify.cu -- host-side candidate-structure classifier behind mirfold_classify() (no device code).
//
// Replaces the per-line Python work of the reference's RNALfold-output parser on the hot path
// (SURVEY.md 8a row a11): get_structures_next_extendregion (miR_PREFeR.py:1541-1599), is_stem_loop
// (:1602-1608), has_one_good_bifurcation (:1611-1659) and filter_ss (:1685-1724).  Once the fold takes
// 80 ms for 10 k loci, 8 s of Python string handling over its 0.5 M hairpins is the stage's cost; here the
// same decisions are taken on the hit table in C++ on all host cores.  The decision rules are the ones
// tests/golden/stage1.json pins (produced by the reference's own functions); the Python versions in
// mir_prefer_b200/structures.py remain the readable statement of the same rules.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/mirfold.h"

namespace {

struct Piece { int begin, end; };   // [begin, end) inside the hit's dot-bracket string

// ss.find(")") - ss.rfind("(") - 1 >= minloop   (-1 for "not found", like str.find)
bool is_stem_loop(const char *ss, int n, int minloop)
{
    int first_close = -1, last_open = -1;
    for (int k = 0; k < n; k++) if (ss[k] == ')') { first_close = k; break; }
    for (int k = n - 1; k >= 0; k--) if (ss[k] == '(') { last_open = k; break; }
    return first_close - last_open - 1 >= minloop;
}

// partner index of every bracket (-1 for dots); false for an unmatched ')' (the reference pops an empty list)
bool pair_table(const char *ss, int n, std::vector<int> &partner, std::vector<int> &stack)
{
    partner.assign((size_t)n, -1);
    stack.clear();
    for (int k = 0; k < n; k++) {
        if (ss[k] == '(') stack.push_back(k);
        else if (ss[k] == ')') {
            if (stack.empty()) return false;
            const int o = stack.back();
            stack.pop_back();
            partner[o] = k; partner[k] = o;
        }
    }
    return true;
}

// filter_ss: one piece per outermost stem, from the end of the previous outermost stem to the start of the next
// one; pieces longer than 55 are kept.  Returns false where the reference raises (no pair at all / unbalanced).
bool filter_ss(const char *ss, int n, std::vector<int> &partner, std::vector<int> &stack, std::vector<Piece> &pieces)
{
    pieces.clear();
    if (!pair_table(ss, n, partner, stack)) return false;
    std::vector<Piece> stems;
    int pos = -1;
    for (int k = 0; k < n; k++) if (ss[k] == '(') { pos = k; break; }
    if (pos < 0) return false;                       // dict_pair[-1] -> KeyError
    while (pos != -1) {
        const int close = partner[pos];
        if (close < 0) return false;                 // unbalanced '('
        stems.push_back({pos, close});
        int next = -1;
        for (int k = close; k < n; k++) if (ss[k] == '(') { next = k; break; }   // ss.find('(', close)
        pos = next;
    }
    for (size_t k = 0; k < stems.size(); k++) {
        const int begin = k == 0 ? 0 : stems[k - 1].end + 1;
        const int end = k + 1 == stems.size() ? n : stems[k + 1].begin;
        if (end - begin > 55) pieces.push_back({begin, end});
    }
    return true;
}

// has_one_good_bifurcation: exactly one place where a stem opens right after another one closed inside an
// enclosing stem, the two inner stems reasonably centred.  rc: 1 true, 0 false, -1 where the reference raises.
int one_good_bifurcation(const char *ss, int n, std::vector<int> &partner, std::vector<int> &stack)
{
    partner.assign((size_t)n, -1);
    stack.clear();
    bool last_pop = false;
    int last_pos = 0, n_bif = 0, left_close = 0, right_open = 0;
    for (int k = 0; k < n; k++) {
        if (ss[k] == '(') {
            if (k != 0 && last_pop) {
                if (stack.empty()) return 0;         // ()() at the top level
                if (n_bif >= 1) return 0;
                n_bif = 1; left_close = last_pos; right_open = k;
            }
            stack.push_back(k);
            last_pop = false; last_pos = k;
        } else if (ss[k] == ')') {
            if (stack.empty()) return -1;
            const int o = stack.back();
            stack.pop_back();
            partner[k] = o; partner[o] = k;
            last_pop = true; last_pos = k;
        }
    }
    if (partner[right_open] < 0 || partner[left_close] < 0) return -1;   // dict_pos[...] -> KeyError
    if ((double)(partner[right_open] - partner[left_close]) / n < 0.5)
        if ((double)partner[left_close] / n > 0.25)
            if ((double)partner[right_open] / n < 0.75) return 1;
    return 0;
}

}  // namespace

extern "C" {

int mirfold_classify(const mirfold_result *res, int minlen, int minloop, mirfold_structure **out, uint64_t *n_out,
                     uint64_t **rec_begin)
{
    if (!res || !out || !n_out || !rec_begin) return MIRFOLD_ERR_ARG;
    if (res->nhits && !res->ss_arena) return MIRFOLD_ERR_ARG;
    *out = nullptr; *n_out = 0; *rec_begin = nullptr;
    const uint32_t nseq = res->nseq;
    const unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    const unsigned nthr = nseq < 256 ? 1u : hw;
    std::vector<std::vector<mirfold_structure>> found(nthr);
    std::vector<std::vector<uint32_t>> count(nthr);
    std::vector<int> bad(nthr, 0);
    auto work = [&](unsigned t) {
        const uint32_t lo = (uint32_t)((uint64_t)nseq * t / nthr), hi = (uint32_t)((uint64_t)nseq * (t + 1) / nthr);
        std::vector<int> partner, stack;
        std::vector<Piece> pieces;
        count[t].assign(hi - lo, 0);
        for (uint32_t r = lo; r < hi; r++) {
            for (uint64_t h = res->hit_begin[r]; h < res->hit_begin[r] + res->hit_count[r]; h++) {
                const mirfold_hit &hit = res->hits[h];
                if (hit.len < minlen) continue;
                const char *ss = res->ss_arena + hit.ss_off;
                const double ne = (hit.mfe_dcal / 100.) / hit.len;   // float("%.2f" % E) / len(ss): the same double
                auto emit = [&](int off, int len, int type) {
                    mirfold_structure s;
                    s.rec = r; s.fold_start = hit.start + off; s.sstype = type; s.len = len;
                    s.ss_off = hit.ss_off + (uint64_t)off; s.norm_energy = ne;
                    found[t].push_back(s);
                    count[t][r - lo]++;
                };
                if (is_stem_loop(ss, hit.len, minloop)) { emit(0, hit.len, 0); continue; }
                if (!filter_ss(ss, hit.len, partner, stack, pieces)) { bad[t] = 1; continue; }
                for (const Piece &p : pieces) {
                    const int len = p.end - p.begin;
                    if (is_stem_loop(ss + p.begin, len, minloop)) emit(p.begin, len, 0);
                    else {
                        const int g = one_good_bifurcation(ss + p.begin, len, partner, stack);
                        if (g < 0) bad[t] = 1;
                        else if (g) emit(p.begin, len, 1);
                    }
                }
            }
        }
    };
    if (nthr == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nthr; t++) th.emplace_back(work, t);
        for (auto &x : th) x.join();
    }
    for (unsigned t = 0; t < nthr; t++) if (bad[t]) return MIRFOLD_ERR_ARG;   // the reference raises on such input
    uint64_t total = 0;
    for (auto &f : found) total += f.size();
    uint64_t *rb = (uint64_t *)malloc(sizeof(uint64_t) * ((size_t)nseq + 1));
    mirfold_structure *arr = (mirfold_structure *)malloc(sizeof(mirfold_structure) * (size_t)(total ? total : 1));
    if (!rb || !arr) { free(rb); free(arr); return MIRFOLD_ERR_NOMEM; }
    uint64_t pos = 0;
    uint32_t r = 0;
    for (unsigned t = 0; t < nthr; t++) {
        if (!found[t].empty()) memcpy(arr + pos, found[t].data(), sizeof(mirfold_structure) * found[t].size());
        uint64_t p = pos;
        for (uint32_t c : count[t]) { rb[r++] = p; p += c; }
        pos += found[t].size();
    }
    rb[nseq] = total;
    *out = arr; *n_out = total; *rec_begin = rb;
    return MIRFOLD_OK;
}

void mirfold_free_structures(mirfold_structure *s, uint64_t *rec_begin)
{
    free(s);
    free(rec_begin);
}

}  // extern "C"
