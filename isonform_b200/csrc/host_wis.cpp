// host_wis.cpp -- host-side helpers of the front half (no GPU work here, by design).
//
// SURVEY.md section 8a scopes M6 (weighted interval scheduling, main:227-272) to the host: it is a tiny
// per-read DP whose result depends on Python-float tie-breaks.  The reference runs it as a pure-Python
// loop per read (main:496-514); once the minimizers and the span search are on the GPU that loop is the
// largest host cost of a batch, so it is restated here in C++ with the SAME IEEE-754 double arithmetic,
// evaluated in the same order (compiled with -ffp-contract=off: no fused multiply-add), for all reads of
// a batch in one call.  isof_span_instances materialises the `seqs` arrays of find_most_supported_span
// (main:324-337) for the picked intervals only.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <numeric>
#include <vector>

#include "../../include/isonform_b200.h"

extern "C" {

int isof_wis_picks(int32_t n_reads, const int64_t *read_rec_off, const int32_t *rec_p1, const int32_t *rec_p2,
                   const int32_t *rec_count, int32_t k, int32_t *pick_rec, int64_t pick_cap, int64_t *pick_off,
                   int64_t *n_picks)
{
    if (n_reads < 0 || !read_rec_off || !pick_off || !n_picks) return ISOF_ERR_ARG;
    std::vector<int32_t> g, valid, p, chosen;
    std::vector<int64_t> starts, stops;
    std::vector<double> v, opt;
    int64_t total = 0;
    int rc = ISOF_OK;
    for (int32_t r = 0; r < n_reads; ++r) {
        pick_off[r] = total;
        g.clear();
        for (int64_t x = read_rec_off[r]; x < read_rec_off[r + 1]; ++x)
            if (rec_count[x] > 0) g.push_back((int32_t)x);            // records that yield an interval, reference order
        const size_t n = g.size();
        if (n == 0) continue;
        // list.sort(key=stop) is stable (main:498)
        std::stable_sort(g.begin(), g.end(), [&](int32_t a, int32_t b) { return rec_p2[a] < rec_p2[b]; });
        starts.resize(n); stops.resize(n); v.resize(n); p.resize(n); opt.assign(n + 1, 0.0);
        valid.clear();
        int64_t max_start = 0;
        for (size_t j = 0; j < n; ++j) {
            starts[j] = (int64_t)rec_p1[g[j]] + k;
            stops[j] = rec_p2[g[j]];
            if (starts[j] < stops[j]) valid.push_back((int32_t)j);
            max_start = std::max(max_start, starts[j]);
        }
        if (max_start > stops[n - 1]) return ISOF_ERR_RANGE;          // the reference indexes past its coordinate table here
        // p[j]: index of the last valid interval with stop <= start_j, 0 when none (fill_p2, main:227-239, including
        // its use as an index into the 1-based opt table)
        for (size_t j = 0; j < n; ++j) {
            // upper_bound over stops[valid] for starts[j]
            size_t lo = 0, hi = valid.size();
            while (lo < hi) {
                const size_t mid = (lo + hi) / 2;
                if (stops[valid[mid]] <= starts[j]) lo = mid + 1; else hi = mid;
            }
            p[j] = lo > 0 ? valid[lo - 1] : 0;
            const double len = (double)(stops[j] - starts[j]) + 0.0001;
            v[j] = (double)((int64_t)rec_count[g[j]] - 1) * len;
        }
        for (size_t j = 1; j <= n; ++j) {
            const double a = v[j - 1] + opt[p[j - 1]];
            const double b = opt[j - 1];
            opt[j] = a > b ? a : b;
        }
        chosen.clear();
        size_t j = n;
        while (j > 0) {
            const double a = v[j - 1] + opt[p[j - 1]];
            if (a > opt[j - 1]) { chosen.push_back((int32_t)(j - 1)); j = (size_t)p[j - 1]; }
            else --j;
        }
        for (size_t c = chosen.size(); c-- > 0;) {
            if (total < pick_cap) pick_rec[total] = g[chosen[c]];
            ++total;
        }
    }
    pick_off[n_reads] = total;
    *n_picks = total;
    if (total > pick_cap) rc = ISOF_ERR_CAPACITY;
    return rc;
}

int isof_span_instances(int64_t n_picks, const int32_t *pick_rec, const int32_t *rec_read, const int32_t *rec_p1,
                        const int32_t *rec_p2, const int32_t *rec_bucket, const int32_t *sorted_rec,
                        const int32_t *bucket_off, const int64_t *read_id, int32_t delta_len, const int64_t *inst_off,
                        uint32_t *inst)
{
    if (n_picks < 0 || (n_picks && (!pick_rec || !inst_off || !inst))) return ISOF_ERR_ARG;
    for (int64_t x = 0; x < n_picks; ++x) {
        const int32_t g = pick_rec[x];
        uint32_t *out = inst + inst_off[x];
        const uint32_t *end = inst + inst_off[x + 1];
        const int32_t span = rec_p2[g] - rec_p1[g], rd = rec_read[g];
        if (out + 3 > end) return ISOF_ERR_CAPACITY;
        *out++ = (uint32_t)read_id[rd]; *out++ = (uint32_t)rec_p1[g]; *out++ = (uint32_t)rec_p2[g];
        const int32_t b = rec_bucket[g];
        for (int32_t s = bucket_off[b]; s < bucket_off[b + 1]; ++s) {
            const int32_t h = sorted_rec[s];
            if (rec_read[h] == rd) continue;
            if (std::abs((rec_p2[h] - rec_p1[h]) - span) >= delta_len) continue;
            if (out + 3 > end) return ISOF_ERR_CAPACITY;
            *out++ = (uint32_t)read_id[rec_read[h]]; *out++ = (uint32_t)rec_p1[h]; *out++ = (uint32_t)rec_p2[h];
        }
        if (out != end) return ISOF_ERR_CAPACITY;          // rec_count[g] is exactly the number of triples
    }
    return ISOF_OK;
}

}  // extern "C"
