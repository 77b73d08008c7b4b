// myers.cu -- kernel (b'): batched unit-cost GLOBAL edit distance with optional traceback, Myers' bit-vector algorithm in
// Hyyro's formulation (G. Myers, JACM 46(3) 1999; H. Hyyro, Nordic J. Computing 10(1) 2003) -- what edlib computes for
// edlib.align(query, target, mode="NW", task="distance" | "path").  The reference declares edlib (setup.py:81) and never
// calls it, so this kernel has no seam to drop into: it serves BASELINE config 5 ("interval-pair
// microbench vs edlib CPU") and north_star's kernel (b) wording; the decisions of the reference stay with the affine
// aligner (sg_align.cu).
//
// Layout of the computation.  A pair (query of n rows, target of m columns) is owned by a GROUP of G = 1..32 lanes, lane
// g holding rows 32g .. 32g+31 of the current stripe as one 32-bit word of the vertical-delta vectors Pv / Mv; a warp
// carries 32/G pairs ("task"), sorted so that the pairs of a task have about the same m.  The G words are advanced as ONE
// 32G-bit integer per column, not block by block: the only non-bitwise operations of the column step are
//   * the addition (Eq & Pv) + Pv -- a multi-word carry chain, resolved with two warp ballots: every lane adds its words,
//     votes "generates a carry" / "would propagate one", and the carries INTO all lanes fall out of one integer addition of
//     the two ballot masks (carry-lookahead by the ALU's own adder);
//   * the shifts Ph << 1, Mh << 1 -- bit 31 of the lane below, one shuffle for both.
// So there is no wavefront skew, no fill and no drain: a column costs the same ~30 instructions whatever G is.  Queries
// longer than 1024 rows run as stripes of 1024; the horizontal deltas of a stripe's bottom row (2 bits per column) go to
// the next stripe through a line buffer in shared memory.  The distance is m + sum of the vertical deltas of the last
// column (popcounts), no per-column score tracking.
// Symbols: the bytes of the pool are renumbered densely (only values that occur), the match vectors Eq[symbol][lane] of a
// stripe sit in shared memory, one conflict-free word load per column.
// Path: per cell two bits -- "diagonal move is optimal" (match, or mismatch and D(i,j) = D(i-1,j-1) + 1) and "vertical
// delta is +1" -- stored as one coalesced 256-byte row per warp and column (0.25 B per cell); the traceback prefers
// diagonal, then up ('I', consumes the query), then left ('D').
#include "common.cuh"

#include <cub/cub.cuh>

namespace {

constexpr int ED_WARPS = 4;                // warps per block of the DP kernel
constexpr unsigned FULL = 0xffffffffu;

struct EdLayout {                          // written on the device, read back once per call
    int32_t cls_start[7];                  // sorted position where class c starts; class c runs groups of G = 32 >> c lanes
    int32_t task_start[7];                 // first task of class c
    int32_t n_sym;                         // distinct byte values in the pool
    int32_t max_m_multi;                   // longest target among multi-stripe pairs (line-buffer size)
    int32_t err;                           // bit 0: pair index outside the pool, bit 1: sequence longer than 2^31-1
    int32_t pad;
    int64_t total_words;                   // trace size of the whole call in uint2 units
};

struct EdDesc {                            // where the traceback finds a pair's trace
    int64_t tbase;                         // uint2 index of the task's trace inside the chunk
    int32_t m_max;                         // column pitch of the task
    int32_t lane0;                         // first lane of the pair's group
};

struct EdArgs {
    const uint8_t *codes;
    const int64_t *off;
    const int32_t *pa, *pb, *order, *pn, *pm;
    const EdLayout *lay;
    const int64_t *task_off;
    uint2 *trace;
    int64_t trace_base;                    // task_off of the chunk's first task
    int64_t trace_cap;                     // uint2 elements behind `trace`
    int32_t t0, t1;                        // tasks of this launch
    int32_t n_sym, lb_words;
    int32_t *dist;
    EdDesc *desc;
    int32_t k;
};

// ---------------------------------------------------------------------------------------------------------------
// symbols: which byte values occur, dense renumbering, encoded pool
// ---------------------------------------------------------------------------------------------------------------
__global__ void ed_presence_kernel(const uint8_t *__restrict__ cat, int64_t n_bases, unsigned int *__restrict__ present)
{
    unsigned int w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_bases; i += stride) {
        const unsigned b = cat[i], bit = 1u << (b & 31), idx = b >> 5;
#pragma unroll
        for (int x = 0; x < 8; ++x) w[x] |= (idx == (unsigned)x) ? bit : 0u;
    }
#pragma unroll
    for (int x = 0; x < 8; ++x) {
        unsigned v = __reduce_or_sync(FULL, w[x]);
        if ((threadIdx.x & 31) == 0 && v) atomicOr(&present[x], v);
    }
}

__global__ void ed_lut_kernel(const unsigned int *__restrict__ present, uint8_t *__restrict__ lut, EdLayout *lay)
{
    const unsigned b = threadIdx.x;          // 256 threads
    unsigned rank = 0;
    for (unsigned x = 0; x < (b >> 5); ++x) rank += __popc(present[x]);
    rank += __popc(present[b >> 5] & ((1u << (b & 31)) - 1u));
    lut[b] = (uint8_t)rank;                  // values that do not occur are never looked up
    if (b == 255) lay->n_sym = (int32_t)(rank + ((present[7] >> 31) & 1u));
}

__global__ void ed_encode_kernel(const uint8_t *__restrict__ cat, int64_t n_bases, const uint8_t *__restrict__ lut,
                                 uint8_t *__restrict__ codes)
{
    __shared__ uint8_t s_lut[256];
    s_lut[threadIdx.x & 255] = lut[threadIdx.x & 255];
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_bases + 16; i += stride)
        codes[i] = i < n_bases ? s_lut[cat[i]] : (uint8_t)0;          // 16 bytes of padding: the column loop reads whole words
}

// ---------------------------------------------------------------------------------------------------------------
// plan: lengths, class (group size), sort key
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int ed_class_of(int n)            // class c -> G = 32 >> c
{
    if (n > 512) return 0;
    if (n > 256) return 1;
    if (n > 128) return 2;
    if (n > 64) return 3;
    if (n > 32) return 4;
    return 5;
}

__global__ void ed_plan_kernel(const int64_t *__restrict__ off, int32_t n_seqs, const int32_t *__restrict__ pa,
                               const int32_t *__restrict__ pb, int64_t P, int32_t *__restrict__ pn, int32_t *__restrict__ pm,
                               unsigned long long *__restrict__ keys, int32_t *__restrict__ idx, unsigned int *__restrict__ cnt,
                               EdLayout *lay)
{
    __shared__ unsigned int s_cnt[6];
    if (threadIdx.x < 6) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < P) {
        const int32_t a = pa[p], b = pb[p];
        int64_t n = 0, m = 0;
        if (a < 0 || a >= n_seqs || b < 0 || b >= n_seqs) atomicOr(&lay->err, 1);
        else { n = off[a + 1] - off[a]; m = off[b + 1] - off[b]; }
        if (n < 0 || m < 0 || n > 0x7fffffff || m > 0x7fffffff) { atomicOr(&lay->err, 2); n = m = 0; }
        pn[p] = (int32_t)n;
        pm[p] = (int32_t)m;
        const int c = ed_class_of((int)n);
        const unsigned long long stripes = c == 0 ? (unsigned long long)((n + 1023) >> 10) : 1ull;
        const unsigned long long cost = stripes * (unsigned long long)m;                 // < 2^53
        keys[p] = ((unsigned long long)c << 56) | (0x00ffffffffffffffull - cost);        // ascending: big groups, long targets first
        idx[p] = (int32_t)p;
        atomicAdd(&s_cnt[c], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 6 && s_cnt[threadIdx.x]) atomicAdd(&cnt[threadIdx.x], s_cnt[threadIdx.x]);
}

__global__ void ed_layout_kernel(const unsigned int *__restrict__ cnt, EdLayout *lay)
{
    int32_t cs = 0, ts = 0;
    for (int c = 0; c < 6; ++c) {
        lay->cls_start[c] = cs;
        lay->task_start[c] = ts;
        const int32_t per = 1 << c;                                                       // pairs per task = 32 / G
        cs += (int32_t)cnt[c];
        ts += ((int32_t)cnt[c] + per - 1) / per;
    }
    lay->cls_start[6] = cs;
    lay->task_start[6] = ts;
}

__device__ __forceinline__ int ed_task_class(const EdLayout *lay, int t)
{
    int c = 0;
#pragma unroll
    for (int x = 1; x < 6; ++x) c += (t >= lay->task_start[x]) ? 1 : 0;
    return c;
}

// trace size per task (uint2 units); a task's column pitch is the m of its first pair (the largest: sorted descending)
__global__ void ed_task_kernel(EdLayout *lay, const int32_t *__restrict__ order, const int32_t *__restrict__ pn,
                               const int32_t *__restrict__ pm, int64_t *__restrict__ task_words, int64_t n_slots, int want_trace)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_slots) return;
    int64_t words = 0;
    if (t < lay->task_start[6]) {
        const int c = ed_task_class(lay, (int)t);
        const int pair = order[lay->cls_start[c] + ((int)t - lay->task_start[c]) * (1 << c)];
        const int64_t stripes = c == 0 ? (((int64_t)pn[pair] + 1023) >> 10) : 1;
        if (stripes > 1) atomicMax(&lay->max_m_multi, pm[pair]);
        if (want_trace) words = stripes * (int64_t)pm[pair] * 32;
    }
    task_words[t] = words;
}

// ---------------------------------------------------------------------------------------------------------------
// the DP
// ---------------------------------------------------------------------------------------------------------------
template <int G, bool TRACE>
__device__ __forceinline__ void ed_run_task(const EdArgs &a, const int t, const int c, uint32_t *__restrict__ peq,
                                            uint32_t *__restrict__ lb, const int lane)
{
    constexpr int PPT = 32 / G;
    const int gl = lane % G, slot = lane / G;
    const EdLayout *lay = a.lay;
    const int sidx = lay->cls_start[c] + (t - lay->task_start[c]) * PPT + slot;
    const bool valid = sidx < lay->cls_start[c + 1];
    int pair = -1, n = 0, m = 0;
    int64_t qb = 0, tb = 0;
    if (valid) {
        pair = a.order[sidx];
        n = a.pn[pair];
        m = a.pm[pair];
        qb = a.off[a.pa[pair]];
        tb = a.off[a.pb[pair]];
    }
    const int m_max = (int)__reduce_max_sync(FULL, (unsigned)m);
    const int ns = (G == 32) ? max(1, (int)(((int64_t)n + 1023) >> 10)) : 1;       // G == 32: one pair per warp, uniform
    const int64_t tbase = TRACE ? a.task_off[t] - a.trace_base : 0;
    const uint32_t *codes32 = reinterpret_cast<const uint32_t *>(a.codes);
    int acc = 0;
    for (int s = 0; s < ns; ++s) {
        // ---- match vectors of this stripe: Eq[symbol][lane], each lane fills its own column of the table
        for (int y = 0; y < a.n_sym; ++y) peq[y * 32 + lane] = 0;
        const int64_t r0 = (int64_t)s * 1024 + gl * 32;
        const int nv = (int)min((int64_t)32, max((int64_t)0, (int64_t)n - r0));
        for (int r = 0; r < nv; ++r) peq[(int)a.codes[qb + r0 + r] * 32 + lane] |= 1u << r;
        __syncwarp();
        uint32_t pv = ~0u, mv = 0u, hist = 0u;
        const bool rd = s > 0, wr = s + 1 < ns;
        for (int j0 = 0; j0 < m_max; j0 += 16) {
            const uint32_t hw = rd ? lb[j0 >> 4] : 0x55555555u;        // horizontal delta entering row 0: +1 per column at the top
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
                const int j4 = j0 + q4 * 4;
                if (j4 >= m_max) break;
                uint32_t c4 = 0;
                if (j4 < m) {                                          // four target symbols from two aligned words
                    const int64_t ad = tb + j4;
                    const uint32_t w0 = codes32[ad >> 2], w1 = codes32[(ad >> 2) + 1];
                    c4 = __funnelshift_r(w0, w1, ((unsigned)ad & 3u) * 8u);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = j4 + u;
                    const bool act = j < m;
                    const uint32_t ch = act ? ((c4 >> (8 * u)) & 0xffu) : 0u;
                    const uint32_t eq0 = peq[ch * 32 + lane];
                    const uint32_t inb = (hw >> (2 * (j & 15))) & 3u;  // bit 0: +1, bit 1: -1 (group lane 0 only)
                    uint32_t eq = eq0;
                    if (G == 1 || gl == 0) eq |= inb >> 1;
                    const uint32_t xv = eq0 | mv;
                    const uint32_t x = eq & pv;
                    uint32_t sum = x + pv;
                    if (G > 1) {
                        // carries between the words of the group: generate / propagate votes, lookahead by one addition
                        const bool top = gl == G - 1;
                        const unsigned gm = __ballot_sync(FULL, !top && sum < x);
                        const unsigned pm_ = __ballot_sync(FULL, !top && sum == 0xffffffffu);
                        const unsigned cy = (((pm_ | gm) + gm) ^ pm_);
                        sum += (cy >> lane) & 1u;
                    }
                    const uint32_t xh = (sum ^ pv) | eq;
                    const uint32_t ph = mv | ~(xh | pv);
                    const uint32_t mh = pv & xh;
                    const uint32_t t2 = (ph >> 31) | ((mh >> 31) << 1);
                    uint32_t up = inb;
                    if (G > 1) {
                        const uint32_t below = __shfl_up_sync(FULL, t2, 1);
                        if (gl != 0) up = below;
                    }
                    const uint32_t ph2 = (ph << 1) | (up & 1u);
                    const uint32_t mh2 = (mh << 1) | (up >> 1);
                    const uint32_t pvn = mh2 | ~(xv | ph2);
                    const uint32_t mvn = ph2 & xv;
                    if (TRACE && act) {
                        // diagonal move optimal: match, or D(i,j) - D(i-1,j-1) = dh(i,j) + dv(i,j-1) = 1
                        const uint32_t dg = eq0 | (ph & ~(pv | mv)) | (pv & ~(ph | mh));
                        ISOF_ASSERT(tbase + ((int64_t)s * m_max + j) * 32 + lane < a.trace_cap, CHK_TRACE_STORE);
                        a.trace[tbase + ((int64_t)s * m_max + j) * 32 + lane] = make_uint2(dg, pvn);
                    }
                    if (G == 32) hist |= t2 << (2 * (j & 15));
                    pv = act ? pvn : pv;
                    mv = act ? mvn : mv;
                }
            }
            if (G == 32) {
                if (wr && lane == 31) lb[j0 >> 4] = hist;              // bottom row of the stripe: the next stripe's carry-in
                hist = 0u;
            }
        }
        const uint32_t rows = nv >= 32 ? ~0u : ((1u << nv) - 1u);
        int d = __popc(pv & rows) - __popc(mv & rows);
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) d += __shfl_xor_sync(FULL, d, o);
        acc += d;
        __syncwarp();
    }
    if (valid && gl == 0) {
        int dist = m + acc;                                            // D(0,m) = m, then down the last column
        if (a.k >= 0 && dist > a.k) dist = -1;
        a.dist[pair] = dist;
        if (TRACE) a.desc[pair] = EdDesc{tbase, m_max, slot * G};
    }
}

template <bool TRACE>
__global__ void __launch_bounds__(ED_WARPS * 32) ed_dp_kernel(const EdArgs a)
{
    extern __shared__ uint32_t ed_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = a.t0 + blockIdx.x * ED_WARPS + warp;
    if (t >= a.t1) return;
    uint32_t *peq = ed_smem + (size_t)warp * (a.n_sym * 32 + a.lb_words);
    uint32_t *lb = peq + a.n_sym * 32;
    const int c = ed_task_class(a.lay, t);
    switch (c) {
    case 0: ed_run_task<32, TRACE>(a, t, c, peq, lb, lane); break;
    case 1: ed_run_task<16, TRACE>(a, t, c, peq, lb, lane); break;
    case 2: ed_run_task<8, TRACE>(a, t, c, peq, lb, lane); break;
    case 3: ed_run_task<4, TRACE>(a, t, c, peq, lb, lane); break;
    case 4: ed_run_task<2, TRACE>(a, t, c, peq, lb, lane); break;
    default: ed_run_task<1, TRACE>(a, t, c, peq, lb, lane); break;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// traceback: one thread per pair; pass 1 counts the runs, pass 2 writes them (alignment order) at their final place
// ---------------------------------------------------------------------------------------------------------------
template <bool WRITE>
__global__ void ed_traceback_kernel(const EdArgs a, int64_t *__restrict__ nruns64, const int64_t *__restrict__ cigar_off,
                                    uint32_t *__restrict__ cigar)
{
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int t = a.t0 + (int)(tid >> 5), slot = (int)(tid & 31);
    if (t >= a.t1) return;
    const EdLayout *lay = a.lay;
    const int c = ed_task_class(lay, t);
    if (slot >= (1 << c)) return;
    const int sidx = lay->cls_start[c] + (t - lay->task_start[c]) * (1 << c) + slot;
    if (sidx >= lay->cls_start[c + 1]) return;
    const int pair = a.order[sidx];
    if (a.dist[pair] < 0) {
        if (!WRITE) nruns64[pair] = 0;
        return;
    }
    const EdDesc d = a.desc[pair];
    const uint8_t *q = a.codes + a.off[a.pa[pair]], *tg = a.codes + a.off[a.pb[pair]];
    int i = a.pn[pair], j = a.pm[pair];
    int64_t count = 0, wpos = WRITE ? cigar_off[pair + 1] - 1 : 0;
    uint32_t cur_op = 0xffu, cur_len = 0;
    auto emit = [&](uint32_t op, uint32_t len) {
        if (op == cur_op) { cur_len += len; return; }
        if (cur_len) {
            if (WRITE) cigar[wpos--] = (cur_len << 4) | cur_op;
            ++count;
        }
        cur_op = op;
        cur_len = len;
    };
    while (i > 0 && j > 0) {
        const int r = i - 1;
        ISOF_ASSERT(d.tbase + ((int64_t)(r >> 10) * d.m_max + (j - 1)) * 32 + d.lane0 + ((r & 1023) >> 5) < a.trace_cap, CHK_TRACE_LOAD);
        const uint2 w = a.trace[d.tbase + ((int64_t)(r >> 10) * d.m_max + (j - 1)) * 32 + d.lane0 + ((r & 1023) >> 5)];
        const uint32_t bit = 1u << (r & 31);
        if (w.x & bit) {
            emit(q[i - 1] == tg[j - 1] ? (uint32_t)ISOF_CIGAR_EQ : (uint32_t)ISOF_CIGAR_X, 1);
            --i; --j;
        } else if (w.y & bit) {
            emit(ISOF_CIGAR_I, 1);
            --i;
        } else {
            emit(ISOF_CIGAR_D, 1);
            --j;
        }
    }
    if (i > 0) emit(ISOF_CIGAR_I, (uint32_t)i);
    if (j > 0) emit(ISOF_CIGAR_D, (uint32_t)j);
    emit(0xfeu, 0);                                                    // flush
    if (!WRITE) nruns64[pair] = count;
}

int ed_launch_dp(isof_ctx *ctx, EdArgs &a, bool trace, size_t smem)
{
    const int blocks = (a.t1 - a.t0 + ED_WARPS - 1) / ED_WARPS;
    if (blocks <= 0) return ISOF_OK;
    LaunchTimer lt(ctx, SLOT_DP);
    if (trace) ed_dp_kernel<true><<<blocks, ED_WARPS * 32, smem, ctx->stream>>>(a);
    else ed_dp_kernel<false><<<blocks, ED_WARPS * 32, smem, ctx->stream>>>(a);
    CUDA_TRY(ctx, cudaGetLastError());
    return ISOF_OK;
}

}  // namespace

unsigned isof_viol_myers() { return isof_read_violations_tu(); }

// Device-pointer core.  h_total_runs: host pointer, receives the number of CIGAR runs of the call.
int isof_edit_distance_core(isof_ctx *ctx, const uint8_t *d_cat, const int64_t *d_off, int32_t n_seqs, int64_t n_bases,
                            const int32_t *d_pa, const int32_t *d_pb, int64_t P, int32_t k, int32_t *d_dist,
                            uint32_t *d_cigar, int64_t cigar_cap, int64_t *d_cigar_off, int64_t *h_total_runs)
{
    cudaStream_t st = ctx->stream;
    if (h_total_runs) *h_total_runs = 0;
    if (P < 0 || n_seqs < 0 || n_bases < 0) return isof_fail(ctx, ISOF_ERR_ARG, "negative count");
    if (P >= ((int64_t)1 << 30)) return isof_fail(ctx, ISOF_ERR_ARG, "more than 2^30 pairs in one call");
    const bool want_path = d_cigar_off != nullptr;
    if (P == 0) {
        if (d_cigar_off) CUDA_TRY(ctx, cudaMemsetAsync(d_cigar_off, 0, sizeof(int64_t), st));
        return ISOF_OK;
    }
    DevBuf *B = ctx->b_ed;
    TRY(dev_reserve(ctx, B[0], (size_t)n_bases + 32));                       // codes
    TRY(dev_reserve(ctx, B[1], 512));                                        // present[8] | cnt[6] | lut[256]
    TRY(dev_reserve(ctx, B[2], sizeof(EdLayout)));
    TRY(dev_reserve(ctx, B[3], sizeof(int32_t) * (size_t)P));                // pn
    TRY(dev_reserve(ctx, B[4], sizeof(int32_t) * (size_t)P));                // pm
    TRY(dev_reserve(ctx, B[5], sizeof(unsigned long long) * (size_t)P * 2)); // keys in | out
    TRY(dev_reserve(ctx, B[6], sizeof(int32_t) * (size_t)P * 2));            // idx in | order
    TRY(dev_reserve(ctx, B[7], sizeof(int64_t) * (size_t)(P + 1) * 2));      // task_words | task_off
    uint8_t *codes = (uint8_t *)B[0].p;
    unsigned int *present = (unsigned int *)B[1].p, *cnt = present + 8;
    uint8_t *lut = (uint8_t *)(present + 16);
    EdLayout *lay = (EdLayout *)B[2].p;
    int32_t *pn = (int32_t *)B[3].p, *pm = (int32_t *)B[4].p;
    unsigned long long *keys = (unsigned long long *)B[5].p, *keys_out = keys + P;
    int32_t *idx = (int32_t *)B[6].p, *order = idx + P;
    int64_t *task_words = (int64_t *)B[7].p, *task_off = task_words + (P + 1);
    CUDA_TRY(ctx, cudaMemsetAsync(present, 0, 512, st));
    CUDA_TRY(ctx, cudaMemsetAsync(lay, 0, sizeof(EdLayout), st));
    {
        LaunchTimer lt(ctx, SLOT_OTHER);
        const int blocks = (int)std::min<int64_t>((n_bases + 1023) / 1024 + 1, (int64_t)ctx->sm_count * 8);
        ed_presence_kernel<<<blocks, 256, 0, st>>>(d_cat, n_bases, present);
    }
    {
        LaunchTimer lt(ctx, SLOT_OTHER);
        ed_lut_kernel<<<1, 256, 0, st>>>(present, lut, lay);
    }
    {
        LaunchTimer lt(ctx, SLOT_OTHER);
        const int blocks = (int)std::min<int64_t>((n_bases + 16 + 1023) / 1024 + 1, (int64_t)ctx->sm_count * 8);
        ed_encode_kernel<<<blocks, 256, 0, st>>>(d_cat, n_bases, lut, codes);
    }
    {
        LaunchTimer lt(ctx, SLOT_OTHER);
        ed_plan_kernel<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(d_off, n_seqs, d_pa, d_pb, P, pn, pm, keys, idx, cnt, lay);
    }
    CUDA_TRY(ctx, cudaGetLastError());
    size_t tmp_sort = 0, tmp_scan = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, keys, keys_out, idx, order, (int)P, 0, 59, st);
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan, task_words, task_off, (int)(P + 1), st);
    TRY(dev_reserve(ctx, B[8], std::max(tmp_sort, tmp_scan)));
    {
        LaunchTimer lt(ctx, SLOT_OTHER);
        CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(B[8].p, tmp_sort, keys, keys_out, idx, order, (int)P, 0, 59, st));
    }
    {
        LaunchTimer lt(ctx, SLOT_OTHER);
        ed_layout_kernel<<<1, 1, 0, st>>>(cnt, lay);
    }
    {
        LaunchTimer lt(ctx, SLOT_OTHER);
        ed_task_kernel<<<(unsigned)((P + 1 + 255) / 256), 256, 0, st>>>(lay, order, pn, pm, task_words, P + 1, want_path ? 1 : 0);
    }
    {
        LaunchTimer lt(ctx, SLOT_OTHER);
        CUDA_TRY(ctx, cub::DeviceScan::ExclusiveSum(B[8].p, tmp_scan, task_words, task_off, (int)(P + 1), st));
    }
    CUDA_TRY(ctx, cudaGetLastError());
    // ---- one read-back: layout + trace total
    EdLayout *h_lay = (EdLayout *)ctx->pinned;
    int64_t *h_total = (int64_t *)((char *)ctx->pinned + 256);
    CUDA_TRY(ctx, cudaMemcpyAsync(h_lay, lay, sizeof(EdLayout), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(h_total, task_off + P, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    if (h_lay->err & 1) return isof_fail(ctx, ISOF_ERR_ARG, "pair index outside the sequence pool");
    if (h_lay->err & 2) return isof_fail(ctx, ISOF_ERR_ARG, "sequence longer than 2^31-1");
    const int T = h_lay->task_start[6];
    const int n_sym = std::max(h_lay->n_sym, 1);
    const int lb_words = h_lay->max_m_multi > 0 ? (h_lay->max_m_multi + 15) / 16 + 1 : 0;
    const int64_t total_words = *h_total;
    const size_t smem = (size_t)ED_WARPS * ((size_t)n_sym * 32 + (size_t)lb_words) * sizeof(uint32_t);
    if (smem > 200 * 1024)
        return isof_fail(ctx, ISOF_ERR_NOMEM, "edit distance: %d symbols and a %d-column multi-stripe target need %zu bytes of "
                         "shared memory per block", n_sym, h_lay->max_m_multi, smem);
    if (smem > 48 * 1024) {
        CUDA_TRY(ctx, cudaFuncSetAttribute(ed_dp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(ctx, cudaFuncSetAttribute(ed_dp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    EdArgs a;
    a.codes = codes; a.off = d_off; a.pa = d_pa; a.pb = d_pb; a.order = order; a.pn = pn; a.pm = pm; a.lay = lay;
    a.task_off = task_off; a.trace = nullptr; a.trace_base = 0; a.trace_cap = 0; a.t0 = 0; a.t1 = T; a.n_sym = n_sym; a.lb_words = lb_words;
    a.dist = d_dist; a.desc = nullptr; a.k = k;
    if (!want_path) return ed_launch_dp(ctx, a, false, smem);

    // ---- path: trace chunks
    TRY(dev_reserve(ctx, B[9], sizeof(EdDesc) * (size_t)P));
    TRY(dev_reserve(ctx, B[10], sizeof(int64_t) * (size_t)(P + 1)));
    a.desc = (EdDesc *)B[9].p;
    int64_t *nruns64 = (int64_t *)B[10].p;
    CUDA_TRY(ctx, cudaMemsetAsync(nruns64, 0, sizeof(int64_t) * (size_t)(P + 1), st));
    size_t budget = ctx->trace_budget;
    if (!budget) {
        size_t fr = 0, tot = 0;
        CUDA_TRY(ctx, cudaMemGetInfo(&fr, &tot));
        budget = fr / 10 * 4 + B[11].cap;
    }
    const int64_t budget_words = std::max<int64_t>((int64_t)(budget / sizeof(uint2)), 1);
    std::vector<int> cuts = {0, T};
    std::vector<int64_t> chunk_base = {0, total_words};
    int64_t need = total_words;
    if (total_words > budget_words) {
        std::vector<int64_t> h_off((size_t)T + 1);
        CUDA_TRY(ctx, cudaMemcpyAsync(h_off.data(), task_off, sizeof(int64_t) * ((size_t)T + 1), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(ctx, cudaStreamSynchronize(st));
        cuts.assign(1, 0);
        chunk_base.assign(1, 0);
        need = 0;
        int t0 = 0;
        while (t0 < T) {
            int t1 = t0 + 1;
            if (h_off[t1] - h_off[t0] > budget_words)
                return isof_fail(ctx, ISOF_ERR_NOMEM, "edit-distance trace of one task (%lld bytes) exceeds the scratch budget",
                                 (long long)((h_off[t1] - h_off[t0]) * 8));
            while (t1 < T && h_off[t1 + 1] - h_off[t0] <= budget_words) ++t1;
            need = std::max(need, h_off[t1] - h_off[t0]);
            cuts.push_back(t1);
            chunk_base.push_back(h_off[t1]);
            t0 = t1;
        }
    }
    TRY(dev_reserve(ctx, B[11], sizeof(uint2) * (size_t)std::max<int64_t>(need, 1)));
    a.trace = (uint2 *)B[11].p;
    a.trace_cap = std::max<int64_t>(need, 1);
    ctx->prof.trace_bytes += total_words * (int64_t)sizeof(uint2);
    const bool single = cuts.size() == 2;
    auto run_chunk = [&](size_t x, bool write) -> int {
        a.t0 = cuts[x];
        a.t1 = cuts[x + 1];
        a.trace_base = chunk_base[x];
        if (!write || !single) TRY(ed_launch_dp(ctx, a, true, smem));
        const int64_t threads = (int64_t)(a.t1 - a.t0) * 32;
        LaunchTimer lt(ctx, SLOT_TB);
        if (write) ed_traceback_kernel<true><<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(a, nruns64, d_cigar_off, d_cigar);
        else ed_traceback_kernel<false><<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(a, nruns64, d_cigar_off, d_cigar);
        CUDA_TRY(ctx, cudaGetLastError());
        return ISOF_OK;
    };
    for (size_t x = 0; x + 1 < cuts.size(); ++x) TRY(run_chunk(x, false));
    size_t tmp2 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp2, nruns64, d_cigar_off, (int)(P + 1), st);
    TRY(dev_reserve(ctx, B[12], tmp2));
    {
        LaunchTimer lt(ctx, SLOT_OTHER);
        CUDA_TRY(ctx, cub::DeviceScan::ExclusiveSum(B[12].p, tmp2, nruns64, d_cigar_off, (int)(P + 1), st));
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(h_total, d_cigar_off + P, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    if (h_total_runs) *h_total_runs = *h_total;
    if (!d_cigar || *h_total > cigar_cap)
        return d_cigar ? isof_fail(ctx, ISOF_ERR_CAPACITY, "cigar buffer too small: need %lld runs", (long long)*h_total)
                       : ISOF_OK;
    for (size_t x = 0; x + 1 < cuts.size(); ++x) TRY(run_chunk(x, true));
    return ISOF_OK;
}
