// host_poa.cpp -- in-process partial-order-alignment consensus (host C++).
//
// The reference obtains every consensus by spawning the external `spoa` binary on two temporary files
// (IsoformGeneration.py:143-156: `spoa <reads.fa> -l 0 -r 0 -g -2`; 887-1503 spawns per batch).
// spoa is not vendored by the reference and its source is not available here, so its output cannot be reproduced
// bit for bit; this file is an independent POA with the same role and the same scoring as that command line
// (-l 0: local alignment; match 5, mismatch -4; -g -2 with spoa's default -e -6 selects its LINEAR gap mode with
// gap -2), meant to be used on BOTH sides of a comparison: as the `spoa` executable of the unmodified reference
// (scripts/isof_spoa) and, in-process, behind `IsoformGeneration.run_spoa` (isonform_b200/poa.py) -- no subprocess,
// no temporary files.
//
// Algorithm (Lee, Grasso, Sharlow 2002; consensus = heaviest bundle): sequences are added one by one to a DAG whose
// nodes carry one base; each new sequence is aligned to the graph by Smith-Waterman over the nodes in topological
// order (a cell looks at every predecessor of its node), the traceback is merged into the graph (same base: reuse the
// node or one of the nodes aligned to it; other base: new node, linked to the aligned group), edge weights count the
// sequences that took the edge.  The consensus follows, from the best-scoring end node backwards, the heaviest
// incoming edge of each node (ties: the predecessor with the larger score, then the smaller node id).
// Deterministic: no hashing, no threads; everything is ordered by node id / input order.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/isonform_b200.h"

namespace {

struct Edge { int to; int w; };

struct Graph {
    std::vector<uint8_t> base;
    std::vector<std::vector<Edge>> out, in;        // in[v] holds {from, w}
    std::vector<std::vector<int>> aligned;         // nodes of other bases aligned to the same column

    int add_node(uint8_t b)
    {
        base.push_back(b); out.emplace_back(); in.emplace_back(); aligned.emplace_back();
        return (int)base.size() - 1;
    }
    void add_edge(int u, int v)
    {
        for (Edge &e : out[u])
            if (e.to == v) { ++e.w; for (Edge &f : in[v]) if (f.to == u) { ++f.w; break; } return; }
        out[u].push_back({v, 1});
        in[v].push_back({u, 1});
    }
    // topological order: Kahn with a min-id frontier kept as a sorted insertion (deterministic); aligned nodes carry
    // no edges between them, so any topological order of the edge DAG is admissible
    std::vector<int> topo() const
    {
        const int n = (int)base.size();
        std::vector<int> indeg(n), order;
        order.reserve(n);
        std::vector<int> heap;
        for (int v = 0; v < n; ++v) { indeg[v] = (int)in[v].size(); if (!indeg[v]) heap.push_back(v); }
        std::make_heap(heap.begin(), heap.end(), std::greater<int>());
        while (!heap.empty()) {
            std::pop_heap(heap.begin(), heap.end(), std::greater<int>());
            const int v = heap.back(); heap.pop_back();
            order.push_back(v);
            for (const Edge &e : out[v])
                if (--indeg[e.to] == 0) { heap.push_back(e.to); std::push_heap(heap.begin(), heap.end(), std::greater<int>()); }
        }
        return order;
    }
};

inline uint8_t up(uint8_t c) { return (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c; }

// local alignment of seq to the graph; returns (node, pos) pairs in sequence order, node or pos = -1 for gaps
void align(const Graph &G, const std::vector<int> &order, const uint8_t *seq, int L, int match, int mismatch, int gap,
           std::vector<std::pair<int, int>> &aln)
{
    aln.clear();
    const int N = (int)order.size();
    if (N == 0 || L == 0) return;
    std::vector<int> rank(G.base.size());
    for (int r = 0; r < N; ++r) rank[order[r]] = r + 1;                 // row 0 = virtual start (all zeros)
    const size_t W = (size_t)L + 1;
    if ((double)(N + 1) * (double)W > 1.5e9) { aln.clear(); aln.push_back({-2, -2}); return; }      // caller reports ISOF_ERR_NOMEM
    std::vector<int32_t> H((size_t)(N + 1) * W, 0);
    int32_t best = 0; int best_r = 0, best_c = 0;
    for (int r = 1; r <= N; ++r) {
        const int v = order[r - 1];
        const uint8_t b = G.base[v];
        int32_t *row = &H[(size_t)r * W];
        const std::vector<Edge> &preds = G.in[v];
        // first predecessor initialises the row, the others are max-ed in
        const int np = preds.empty() ? 1 : (int)preds.size();
        for (int k = 0; k < np; ++k) {
            const int32_t *prow = &H[(size_t)(preds.empty() ? 0 : rank[preds[k].to]) * W];
            if (k == 0) {
                for (int c = 1; c <= L; ++c) {
                    const int32_t d = prow[c - 1] + (b == seq[c - 1] ? match : mismatch);
                    const int32_t u = prow[c] + gap;
                    row[c] = d > u ? d : u;
                }
            } else {
                for (int c = 1; c <= L; ++c) {
                    const int32_t d = prow[c - 1] + (b == seq[c - 1] ? match : mismatch);
                    const int32_t u = prow[c] + gap;
                    const int32_t m = d > u ? d : u;
                    if (m > row[c]) row[c] = m;
                }
            }
        }
        for (int c = 1; c <= L; ++c) {
            const int32_t l = row[c - 1] + gap;
            if (l > row[c]) row[c] = l;
            if (row[c] < 0) row[c] = 0;
            if (row[c] > best) { best = row[c]; best_r = r; best_c = c; }     // first maximum in (row, column) order
        }
    }
    if (best <= 0) return;
    // traceback: diagonal over the predecessors in stored order, then the vertical moves, then the horizontal one
    int r = best_r, c = best_c;
    std::vector<std::pair<int, int>> rev;
    while (r > 0 && c >= 0) {
        const int32_t h = H[(size_t)r * W + c];
        if (h == 0) break;
        const int v = order[r - 1];
        const std::vector<Edge> &preds = G.in[v];
        const int np = preds.empty() ? 1 : (int)preds.size();
        bool moved = false;
        if (c > 0) {
            const int32_t s = (G.base[v] == seq[c - 1]) ? match : mismatch;
            for (int k = 0; k < np && !moved; ++k) {
                const int pr = preds.empty() ? 0 : rank[preds[k].to];
                if (H[(size_t)pr * W + (c - 1)] + s == h) { rev.push_back({v, c - 1}); r = pr; --c; moved = true; }
            }
        }
        for (int k = 0; k < np && !moved; ++k) {
            const int pr = preds.empty() ? 0 : rank[preds[k].to];
            if (H[(size_t)pr * W + c] + gap == h) { rev.push_back({v, -1}); r = pr; moved = true; }
        }
        if (!moved && c > 0 && H[(size_t)r * W + (c - 1)] + gap == h) { rev.push_back({-1, c - 1}); --c; moved = true; }
        if (!moved) break;
    }
    aln.assign(rev.rbegin(), rev.rend());
}

void add_sequence(Graph &G, const uint8_t *seq, int L, const std::vector<std::pair<int, int>> &aln)
{
    if (L == 0) return;
    int first_pos = L, last_pos = -1;
    for (const auto &pr : aln) if (pr.second >= 0) { first_pos = std::min(first_pos, pr.second); last_pos = std::max(last_pos, pr.second); }
    int prev = -1;
    auto chain = [&](int from, int to) {           // unaligned stretch seq[from, to): new nodes
        for (int q = from; q < to; ++q) {
            const int v = G.add_node(seq[q]);
            if (prev >= 0) G.add_edge(prev, v);
            prev = v;
        }
    };
    if (last_pos < 0) { chain(0, L); return; }
    chain(0, first_pos);
    for (const auto &pr : aln) {
        const int node = pr.first, pos = pr.second;
        if (pos < 0) continue;                     // graph node skipped by this sequence
        int cur;
        if (node < 0) cur = G.add_node(seq[pos]);
        else if (G.base[node] == seq[pos]) cur = node;
        else {
            cur = -1;
            for (int a : G.aligned[node]) if (G.base[a] == seq[pos]) { cur = a; break; }
            if (cur < 0) {
                cur = G.add_node(seq[pos]);
                for (int a : G.aligned[node]) { G.aligned[a].push_back(cur); G.aligned[cur].push_back(a); }
                G.aligned[node].push_back(cur);
                G.aligned[cur].push_back(node);
            }
        }
        if (prev >= 0 && prev != cur) G.add_edge(prev, cur);
        prev = cur;
    }
    chain(last_pos + 1, L);
}

void consensus(const Graph &G, std::vector<uint8_t> &out)
{
    out.clear();
    const std::vector<int> order = G.topo();
    const int n = (int)G.base.size();
    std::vector<int64_t> score(n, 0);
    std::vector<int> pred(n, -1);
    int best = -1;
    for (int v : order) {
        int bw = -1;
        for (const Edge &e : G.in[v]) {
            const int u = e.to;
            const bool better = pred[v] < 0 || e.w > bw ||
                                (e.w == bw && (score[u] > score[pred[v]] || (score[u] == score[pred[v]] && u < pred[v])));
            if (better) { bw = e.w; pred[v] = u; }
        }
        if (pred[v] >= 0) score[v] = score[pred[v]] + bw;
        if (best < 0 || score[v] > score[best]) best = v;
    }
    for (int v = best; v >= 0; v = pred[v]) out.push_back(G.base[v]);
    std::reverse(out.begin(), out.end());
}

}  // namespace

extern "C" int isof_poa_consensus(const uint8_t *seq_cat, const int64_t *seq_off, int32_t n_seqs, int32_t match,
                                  int32_t mismatch, int32_t gap, uint8_t *out, int64_t out_cap, int64_t *out_len)
{
    if (n_seqs < 0 || !seq_off || !out_len || gap > 0) return ISOF_ERR_ARG;
    Graph G;
    std::vector<std::pair<int, int>> aln;
    std::vector<uint8_t> s;
    for (int32_t q = 0; q < n_seqs; ++q) {
        const int64_t L = seq_off[q + 1] - seq_off[q];
        if (L <= 0) continue;
        s.resize((size_t)L);
        for (int64_t x = 0; x < L; ++x) s[(size_t)x] = up(seq_cat[seq_off[q] + x]);
        if (G.base.empty()) { add_sequence(G, s.data(), (int)L, aln); continue; }
        const std::vector<int> order = G.topo();
        align(G, order, s.data(), (int)L, match, mismatch, gap, aln);
        if (aln.size() == 1 && aln[0].first == -2) return ISOF_ERR_NOMEM;       // DP matrix beyond 6 GB
        add_sequence(G, s.data(), (int)L, aln);
    }
    std::vector<uint8_t> cons;
    consensus(G, cons);
    *out_len = (int64_t)cons.size();
    if ((int64_t)cons.size() > out_cap) return ISOF_ERR_CAPACITY;
    if (!cons.empty()) memcpy(out, cons.data(), cons.size());
    return ISOF_OK;
}
