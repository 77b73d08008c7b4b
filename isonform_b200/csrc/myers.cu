// myers.cu -- kernel (b'): batched unit-cost GLOBAL edit distance with optional traceback, Myers' bit-vector algorithm in
// Hyyro's formulation (G. Myers, JACM 46(3) 1999; H. Hyyro, Nordic J. Computing 10(1) 2003) -- what edlib computes for
// edlib.align(query, target, mode="NW", task="distance" | "path").  The reference declares edlib (setup.py:81) and never
// calls it, so this kernel has no seam to drop into: it serves BASELINE config 5 ("interval-pair
// microbench vs edlib CPU") and north_star's kernel (b) wording; the decisions of the reference stay with the affine
// aligner (sg_align.cu).
//
// Layout of the computation.  A pair (query of n rows, target of m columns) is owned by a GROUP of G = 1..32 lanes, lane
// g holding W = 1..8 words (32 rows each) of the vertical-delta vectors Pv / Mv of the current stripe; a warp carries 32/G
// pairs ("task"), sorted so that the pairs of a task have about the same m.  The G x W words are advanced as ONE wide
// integer per column, not block by block: the only non-bitwise operations of the column step are
//   * the addition (Eq & Pv) + Pv -- inside a lane an add.cc / addc.cc chain, across lanes two warp ballots: every lane
//     votes "my words generate a carry" / "my words would propagate one", and the carries INTO all lanes fall out of one
//     integer addition of the two ballot masks (carry-lookahead by the ALU's own adder);
//   * the shifts Ph << 1, Mh << 1 -- bit 31 of the lane below, one shuffle for both, funnel shifts inside the lane.
// So there is no wavefront skew, no fill and no drain, and the cross-lane part is amortised over up to 256 rows per lane.
// Queries longer than 32 x 32 x W rows run as stripes; the horizontal deltas of a stripe's bottom row (2 bits per column)
// go to the next stripe through a line buffer in shared memory.  The distance is m + sum of the vertical deltas of the
// last column (popcounts), no per-column score tracking.
// With a bound k (edlib's k) long pairs of distance-only calls run banded: a ring of lanes, one word each, slides down the
// band | i - j | <= k (ed_run_banded).
// Symbols: the bytes of the pool are renumbered densely (only values that occur), the match vectors Eq[symbol][lane][word]
// of a stripe sit in shared memory, conflict-free vector loads per column.
// Path: per cell two bits -- "diagonal move is optimal" (match, or mismatch and D(i,j) = D(i-1,j-1) + 1) and, on such a
// cell, "mismatch", elsewhere "vertical delta is +1" -- 0.25 B per cell, four columns of a lane's word per 32-byte sector, written through a shared-memory
// transpose in rows of 512 contiguous bytes; the traceback prefers diagonal, then up ('I', consumes the query), then left
// ('D'); one thread per pair, or one warp per pair (ballot over the diagonal look-ahead) for calls of few long pairs.
#include "common.cuh"

#include <cub/cub.cuh>

namespace {

constexpr int ED_WARPS = 4;                // warps per block of the DP kernel (fewer when the symbol tables are large)
constexpr unsigned FULL = 0xffffffffu;
constexpr int ED_NC = 48;                  // class ids: (5 - lg2 G) * 8 + (8 - W); lanes per pair G = 1..32, words per lane W = 1..8
constexpr int ED_MAXW_DIST = 8;            // words per lane: distances only
constexpr int ED_MAXW_PATH = 4;            // with the trace (its block of four columns is buffered in registers)

// Biggest first.  A pair of n rows holds ceil(n/32) words: up to max_w stay in ONE thread (G = 1, W = their number), more are
// spread over the smallest power-of-two group whose lanes then hold more than max_w / 2 words each.
__host__ __device__ __forceinline__ int ed_lgg(int c) { return 5 - (c >> 3); }
__host__ __device__ __forceinline__ int ed_w(int c) { return 8 - (c & 7); }
// widest W of a call: the match vectors Eq[symbol][lane][word] of a warp must fit shared memory (256 symbols x 8 words would
// be 256 KB), and the traced kernel buffers a block of four columns in registers
__host__ __device__ __forceinline__ int ed_max_w(int n_sym, bool trace)
{
    return trace ? ED_MAXW_PATH : (n_sym <= 64 ? ED_MAXW_DIST : ED_MAXW_PATH);
}

struct EdLayout {                          // written on the device, read back once per call
    int32_t cls_start[ED_NC + 1];          // sorted position where class c starts
    int32_t task_start[ED_NC + 1];         // first task of class c (a task = one warp = 32 / G pairs)
    int32_t n_sym;                         // distinct byte values in the pool
    int32_t max_m_multi;                   // longest target among multi-stripe pairs (line-buffer size)
    int32_t err;                           // bit 0: pair index outside the pool, bit 1: sequence longer than 2^31-1
    int32_t pad;
};

struct EdDesc {                            // where the traceback finds a pair's trace
    int64_t tbase;                         // uint2 index of the task's trace inside the chunk
    int32_t n_blk;                         // column pitch of the task in blocks of four columns
    int32_t lane0_cls;                     // first lane of the pair's group | class << 8
};

struct EdArgs {
    const uint8_t *codes;
    const int64_t *off;
    const int32_t *pa, *pb, *order, *pn, *pm;
    const EdLayout *lay;
    const int64_t *task_off;
    uint2 *trace;
    const uint32_t *eqg;                   // banded classes: match vectors of every pool sequence, [symbol][word], in global memory
    int64_t eq_tw;                         // words per symbol plane
    int64_t trace_base;                    // task_off of the chunk's first task
    int64_t trace_cap;                     // uint2 elements behind `trace`
    int32_t t0, t1;                        // tasks of this launch
    int32_t n_sym, eq_words, lb_words, stg_words;     // per-warp shared memory: Eq tables | line buffer | trace staging
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
// plan: lengths, class (group size, words per lane), sort key
// ---------------------------------------------------------------------------------------------------------------
// Banded class (Ukkonen): with a bound k only the cells | i - j | <= k can lie on a path of cost <= k, i.e. per column a
// window of at most (2k >> 5) + 2 words of the query.  Such a pair runs on a RING of G lanes, one word each, that slides
// down the query (ed_run_banded) -- when that is at least ~3x less work than all its words.  Class id: the unused (G, W = 1).
__host__ __device__ __forceinline__ int ed_band_class(int64_t n, int64_t m, int k, int n_sym, bool trace)
{
    if (k < 0 || trace || n_sym > 8) return -1;
    const int64_t dlen = n > m ? n - m : m - n;
    if (dlen > k) return -1;
    const int64_t bw = ((int64_t)2 * k >> 5) + 2;                      // words a column's window can touch
    if (bw > 32) return -1;
    int lg = 1;
    while ((1 << lg) < bw) ++lg;
    if (((n + 31) >> 5) < (int64_t)3 << lg) return -1;
    return (5 - lg) * 8 + 7;
}
__host__ __device__ __forceinline__ bool ed_is_band(int c) { return (c & 7) == 7 && c < 40; }

__host__ __device__ __forceinline__ int ed_class_of(int64_t n, int max_w, int64_t *stripes)
{
    const int64_t wt = (n + 31) >> 5;                                  // words of the query
    *stripes = 1;
    if (wt <= max_w) return 5 * 8 + (8 - (wt < 1 ? 1 : (int)wt));      // one thread per pair
    const int64_t ns = (wt + 32 * max_w - 1) / (32 * max_w);           // stripes of at most 32 lanes x max_w words
    const int pw = (int)((wt + ns - 1) / ns);                          // words per stripe
    int lg = 1;
    while ((max_w << lg) < pw) ++lg;                                   // smallest group with <= max_w words per lane
    const int w = (pw + (1 << lg) - 1) >> lg;
    *stripes = (n + ((int64_t)1024 * w) - 1) / ((int64_t)1024 * w);    // as the kernel counts them (G = 32 only when > 1)
    return (5 - lg) * 8 + (8 - w);
}

__global__ void ed_plan_kernel(const int64_t *__restrict__ off, int32_t n_seqs, const int32_t *__restrict__ pa,
                               const int32_t *__restrict__ pb, int64_t P, int32_t *__restrict__ pn, int32_t *__restrict__ pm,
                               unsigned long long *__restrict__ keys, int32_t *__restrict__ idx, unsigned int *__restrict__ cnt,
                               EdLayout *lay, int64_t *__restrict__ ops_words, int want_trace, int k)
{
    const int max_w = ed_max_w(lay->n_sym, want_trace != 0);
    __shared__ unsigned int s_cnt[ED_NC];
    if (threadIdx.x < ED_NC) s_cnt[threadIdx.x] = 0;
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
        int64_t stripes;
        int c = ed_band_class(n, m, k, lay->n_sym, want_trace != 0);
        if (c < 0) c = ed_class_of(n, max_w, &stripes);
        else stripes = 1;
        const unsigned long long cost = (unsigned long long)stripes * (unsigned long long)m;      // < 2^53
        keys[p] = ((unsigned long long)c << 56) | (0x00ffffffffffffffull - cost);        // ascending: big groups, long targets first
        idx[p] = (int32_t)p;
        // path scratch: 2 bits per step for the thread-per-pair walk, a byte per step for pairs that may take the warp walk
        if (ops_words) ops_words[p] = (n > 128 ? ((n + m + 3) >> 2) : ((n + m + 15) >> 4)) + 1;
        atomicAdd(&s_cnt[c], 1u);
    }
    __syncthreads();
    if (threadIdx.x < ED_NC && s_cnt[threadIdx.x]) atomicAdd(&cnt[threadIdx.x], s_cnt[threadIdx.x]);
}

__global__ void ed_layout_kernel(const unsigned int *__restrict__ cnt, EdLayout *lay)
{
    int32_t cs = 0, ts = 0;
    for (int c = 0; c < ED_NC; ++c) {
        lay->cls_start[c] = cs;
        lay->task_start[c] = ts;
        const int32_t per = 32 >> ed_lgg(c);                                              // pairs per task = 32 / G
        cs += (int32_t)cnt[c];
        ts += ((int32_t)cnt[c] + per - 1) / per;
    }
    lay->cls_start[ED_NC] = cs;
    lay->task_start[ED_NC] = ts;
}

__device__ __forceinline__ int ed_task_class(const EdLayout *lay, int t)
{
    int c = 0;
#pragma unroll
    for (int x = 1; x < ED_NC; ++x) c += (t >= lay->task_start[x]) ? 1 : 0;
    return c;
}

// trace size per task (uint2 units); a task's column pitch is the m of its first pair (the largest: sorted descending)
__global__ void ed_task_kernel(EdLayout *lay, const int32_t *__restrict__ order, const int32_t *__restrict__ pn,
                               const int32_t *__restrict__ pm, int64_t *__restrict__ task_words, int64_t n_slots, int want_trace)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_slots) return;
    int64_t words = 0;
    if (t < lay->task_start[ED_NC]) {
        const int c = ed_task_class(lay, (int)t);
        const int pair = order[lay->cls_start[c] + ((int)t - lay->task_start[c]) * (32 >> ed_lgg(c))];
        int64_t stripes = 1;
        if (!ed_is_band(c)) ed_class_of(pn[pair], ed_max_w(lay->n_sym, want_trace != 0), &stripes);
        if (stripes > 1) atomicMax(&lay->max_m_multi, pm[pair]);
        if (want_trace) words = stripes * (int64_t)((pm[pair] + 3) & ~3) * 32 * ed_w(c);      // columns in blocks of four
    }
    task_words[t] = words;
}

// ---------------------------------------------------------------------------------------------------------------
// the DP
// ---------------------------------------------------------------------------------------------------------------
// s = x + y over W words; returns the carry out of the top word (one add.cc / addc.cc chain)
template <int W>
__device__ __forceinline__ uint32_t ed_add(const uint32_t (&x)[W], const uint32_t (&y)[W], uint32_t (&s)[W])
{
    uint32_t co;
    if constexpr (W == 1)
        asm("add.cc.u32 %0, %2, %3;\n\taddc.u32 %1, 0, 0;" : "=r"(s[0]), "=r"(co) : "r"(x[0]), "r"(y[0]));
    else if constexpr (W == 2)
        asm("add.cc.u32 %0, %3, %5;\n\taddc.cc.u32 %1, %4, %6;\n\taddc.u32 %2, 0, 0;" : "=r"(s[0]), "=r"(s[1]), "=r"(co) : "r"(x[0]), "r"(x[1]), "r"(y[0]), "r"(y[1]));
    else if constexpr (W == 3)
        asm("add.cc.u32 %0, %4, %7;\n\taddc.cc.u32 %1, %5, %8;\n\taddc.cc.u32 %2, %6, %9;\n\taddc.u32 %3, 0, 0;" : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(co) : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(y[0]), "r"(y[1]), "r"(y[2]));
    else if constexpr (W == 4)
        asm("add.cc.u32 %0, %5, %9;\n\taddc.cc.u32 %1, %6, %10;\n\taddc.cc.u32 %2, %7, %11;\n\taddc.cc.u32 %3, %8, %12;\n\taddc.u32 %4, 0, 0;" : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(co) : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]));
    else if constexpr (W == 5)
        asm("add.cc.u32 %0, %6, %11;\n\taddc.cc.u32 %1, %7, %12;\n\taddc.cc.u32 %2, %8, %13;\n\taddc.cc.u32 %3, %9, %14;\n\taddc.cc.u32 %4, %10, %15;\n\taddc.u32 %5, 0, 0;" : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(s[4]), "=r"(co) : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]), "r"(y[4]));
    else if constexpr (W == 6)
        asm("add.cc.u32 %0, %7, %13;\n\taddc.cc.u32 %1, %8, %14;\n\taddc.cc.u32 %2, %9, %15;\n\taddc.cc.u32 %3, %10, %16;\n\taddc.cc.u32 %4, %11, %17;\n\taddc.cc.u32 %5, %12, %18;\n\taddc.u32 %6, 0, 0;" : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(s[4]), "=r"(s[5]), "=r"(co) : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]), "r"(y[4]), "r"(y[5]));
    else if constexpr (W == 7)
        asm("add.cc.u32 %0, %8, %15;\n\taddc.cc.u32 %1, %9, %16;\n\taddc.cc.u32 %2, %10, %17;\n\taddc.cc.u32 %3, %11, %18;\n\taddc.cc.u32 %4, %12, %19;\n\taddc.cc.u32 %5, %13, %20;\n\taddc.cc.u32 %6, %14, %21;\n\taddc.u32 %7, 0, 0;" : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(s[4]), "=r"(s[5]), "=r"(s[6]), "=r"(co) : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]), "r"(y[4]), "r"(y[5]), "r"(y[6]));
    else if constexpr (W == 8)
        asm("add.cc.u32 %0, %9, %17;\n\taddc.cc.u32 %1, %10, %18;\n\taddc.cc.u32 %2, %11, %19;\n\taddc.cc.u32 %3, %12, %20;\n\taddc.cc.u32 %4, %13, %21;\n\taddc.cc.u32 %5, %14, %22;\n\taddc.cc.u32 %6, %15, %23;\n\taddc.cc.u32 %7, %16, %24;\n\taddc.u32 %8, 0, 0;" : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(s[4]), "=r"(s[5]), "=r"(s[6]), "=r"(s[7]), "=r"(co) : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]), "r"(y[0]), "r"(y[1]), "r"(y[2]), "r"(y[3]), "r"(y[4]), "r"(y[5]), "r"(y[6]), "r"(y[7]));
    return co;
}

// s += c (0 or 1), rippling through the words
template <int W>
__device__ __forceinline__ void ed_add_carry(uint32_t (&s)[W], uint32_t c)
{
    if constexpr (W == 1)
        s[0] += c;
    else if constexpr (W == 2)
        asm("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, 0;" : "+r"(s[0]), "+r"(s[1]) : "r"(c));
    else if constexpr (W == 3)
        asm("add.cc.u32 %0, %0, %3;\n\taddc.cc.u32 %1, %1, 0;\n\taddc.u32 %2, %2, 0;" : "+r"(s[0]), "+r"(s[1]), "+r"(s[2]) : "r"(c));
    else if constexpr (W == 4)
        asm("add.cc.u32 %0, %0, %4;\n\taddc.cc.u32 %1, %1, 0;\n\taddc.cc.u32 %2, %2, 0;\n\taddc.u32 %3, %3, 0;" : "+r"(s[0]), "+r"(s[1]), "+r"(s[2]), "+r"(s[3]) : "r"(c));
    else if constexpr (W == 5)
        asm("add.cc.u32 %0, %0, %5;\n\taddc.cc.u32 %1, %1, 0;\n\taddc.cc.u32 %2, %2, 0;\n\taddc.cc.u32 %3, %3, 0;\n\taddc.u32 %4, %4, 0;" : "+r"(s[0]), "+r"(s[1]), "+r"(s[2]), "+r"(s[3]), "+r"(s[4]) : "r"(c));
    else if constexpr (W == 6)
        asm("add.cc.u32 %0, %0, %6;\n\taddc.cc.u32 %1, %1, 0;\n\taddc.cc.u32 %2, %2, 0;\n\taddc.cc.u32 %3, %3, 0;\n\taddc.cc.u32 %4, %4, 0;\n\taddc.u32 %5, %5, 0;" : "+r"(s[0]), "+r"(s[1]), "+r"(s[2]), "+r"(s[3]), "+r"(s[4]), "+r"(s[5]) : "r"(c));
    else if constexpr (W == 7)
        asm("add.cc.u32 %0, %0, %7;\n\taddc.cc.u32 %1, %1, 0;\n\taddc.cc.u32 %2, %2, 0;\n\taddc.cc.u32 %3, %3, 0;\n\taddc.cc.u32 %4, %4, 0;\n\taddc.cc.u32 %5, %5, 0;\n\taddc.u32 %6, %6, 0;" : "+r"(s[0]), "+r"(s[1]), "+r"(s[2]), "+r"(s[3]), "+r"(s[4]), "+r"(s[5]), "+r"(s[6]) : "r"(c));
    else if constexpr (W == 8)
        asm("add.cc.u32 %0, %0, %8;\n\taddc.cc.u32 %1, %1, 0;\n\taddc.cc.u32 %2, %2, 0;\n\taddc.cc.u32 %3, %3, 0;\n\taddc.cc.u32 %4, %4, 0;\n\taddc.cc.u32 %5, %5, 0;\n\taddc.cc.u32 %6, %6, 0;\n\taddc.u32 %7, %7, 0;" : "+r"(s[0]), "+r"(s[1]), "+r"(s[2]), "+r"(s[3]), "+r"(s[4]), "+r"(s[5]), "+r"(s[6]), "+r"(s[7]) : "r"(c));
}

template <int W>
__device__ __forceinline__ void ed_load_eq(const uint32_t *p, uint32_t (&e)[W])
{
    if constexpr (W % 4 == 0) {
#pragma unroll
        for (int w = 0; w < W; w += 4) {
            const uint4 v = *reinterpret_cast<const uint4 *>(p + w);
            e[w] = v.x; e[w + 1] = v.y; e[w + 2] = v.z; e[w + 3] = v.w;
        }
    } else if constexpr (W % 2 == 0) {
#pragma unroll
        for (int w = 0; w < W; w += 2) {
            const uint2 v = *reinterpret_cast<const uint2 *>(p + w);
            e[w] = v.x; e[w + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int w = 0; w < W; ++w) e[w] = p[w];
    }
}

template <int G, int W, bool TRACE>
__device__ __forceinline__ void ed_run_task(const EdArgs &a, const int t, const int c, uint32_t *__restrict__ peq,
                                            uint32_t *__restrict__ lb, uint4 *__restrict__ stg, const int lane)
{
    constexpr int PPT = 32 / G;
    constexpr int RL = 32 * W;                                         // rows per lane
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
    // k given: a pair whose lengths differ by more than k is above it whatever the sequences are (every path has at least
    // | n - m | gap columns) -- reported without running its columns
    const bool cut = a.k >= 0 && abs(n - m) > a.k;
    if (cut) m = 0;
    const int m_max = (int)__reduce_max_sync(FULL, (unsigned)m);
    const int n_blk = (m_max + 3) >> 2;                                // trace pitch: blocks of four columns
    const int ns = (G == 32) ? max(1, (int)(((int64_t)n + 32 * RL - 1) / (32 * RL))) : 1;     // G == 32: one pair per warp, uniform
    const int64_t tbase = TRACE ? a.task_off[t] - a.trace_base : 0;
    const uint32_t *codes32 = reinterpret_cast<const uint32_t *>(a.codes);
    uint32_t *my_eq = peq + lane * W;                                  // Eq[symbol][lane][word]: symbol pitch 32 * W words
    int acc = 0;
    for (int s = 0; s < ns; ++s) {
        // ---- match vectors of this stripe, each lane fills its own words of the table
        for (int y = 0; y < a.n_sym; ++y)
#pragma unroll
            for (int w = 0; w < W; ++w) my_eq[y * 32 * W + w] = 0;
        const int64_t r0 = (int64_t)s * 32 * RL + gl * RL;
        const int nv = (int)min((int64_t)RL, max((int64_t)0, (int64_t)n - r0));      // valid rows of this lane
        for (int r = 0; r < nv; ++r) my_eq[(int)a.codes[qb + r0 + r] * 32 * W + (r >> 5)] |= 1u << (r & 31);
        __syncwarp();
        uint32_t pv[W], mv[W];
#pragma unroll
        for (int w = 0; w < W; ++w) { pv[w] = ~0u; mv[w] = 0u; }
        uint32_t hist = 0u;
        const bool rd = (G == 32) && s > 0, wr = (G == 32) && s + 1 < ns;
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
                uint32_t tdg[TRACE ? 4 : 1][W], tpv[TRACE ? 4 : 1][W];      // the block's trace bits
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = j4 + u;
                    const bool act = j < m;
                    // symbols read past m belong to the next sequence or the padding: valid table rows, results discarded
                    const uint32_t ch = __byte_perm(c4, 0, 0x4440 + u);
                    uint32_t eq0[W], eq[W], x[W], sum[W], ph[W], mh[W];
                    ed_load_eq<W>(my_eq + ch * (32 * W), eq0);
                    const uint32_t inb = (G == 32) ? ((hw >> (2 * (j & 15))) & 3u) : 1u;   // bit 0: +1, bit 1: -1
#pragma unroll
                    for (int w = 0; w < W; ++w) eq[w] = eq0[w];
                    if (G == 1 || gl == 0) eq[0] |= inb >> 1;
#pragma unroll
                    for (int w = 0; w < W; ++w) x[w] = eq[w] & pv[w];
                    const uint32_t co = ed_add<W>(x, pv, sum);
                    if (G > 1) {
                        // carries between the lanes of the group: generate / propagate votes, lookahead by one addition
                        uint32_t all = sum[0];
#pragma unroll
                        for (int w = 1; w < W; ++w) all &= sum[w];
                        const bool top = gl == G - 1;
                        const unsigned gm = __ballot_sync(FULL, !top && co != 0u);
                        const unsigned pk = __ballot_sync(FULL, !top && all == 0xffffffffu);
                        const unsigned cy = (((pk | gm) + gm) ^ pk);
                        ed_add_carry<W>(sum, (cy >> lane) & 1u);
                    }
#pragma unroll
                    for (int w = 0; w < W; ++w) {
                        const uint32_t xh = (sum[w] ^ pv[w]) | eq[w];
                        ph[w] = mv[w] | ~(xh | pv[w]);
                        mh[w] = pv[w] & xh;
                    }
                    // bit 31 of the lane's top words for the lane above: byte 0 <- ph, byte 1 <- mh
                    const uint32_t t2 = __byte_perm(ph[W - 1], mh[W - 1], 0x0073);
                    uint32_t up = ((inb & 1u) << 7) | ((inb & 2u) << 14);
                    if (G > 1) {
                        const uint32_t below = __shfl_up_sync(FULL, t2, 1);
                        if (gl != 0) up = below;
                    }
                    uint32_t pcar = up << 24, mcar = up << 16;           // bit 31 = the bit shifted in
                    uint32_t dg[W], pvn[W];
#pragma unroll
                    for (int w = 0; w < W; ++w) {
                        // (tried: the in-lane shifts of words 1.. as IMAD.HI + IMAD by a run-time 2, to move them from the
                        // saturated ALU pipe to the idle FMA pipe -- 3 % slower: four instructions for two)
                        const uint32_t ph2 = __funnelshift_l(pcar, ph[w], 1);
                        const uint32_t mh2 = __funnelshift_l(mcar, mh[w], 1);
                        pcar = ph[w];
                        mcar = mh[w];
                        const uint32_t xv = eq0[w] | mv[w];
                        pvn[w] = mh2 | ~(xv | ph2);
                        const uint32_t mvn = ph2 & xv;
                        // diagonal move optimal: match, or D(i,j) - D(i-1,j-1) = dh(i,j) + dv(i,j-1) = 1
                        if (TRACE) dg[w] = eq0[w] | (ph[w] & ~(pv[w] | mv[w])) | (pv[w] & ~(ph[w] | mh[w]));
                        mv[w] = act ? mvn : mv[w];
                    }
                    if (TRACE) {
#pragma unroll
                        // second bit: on a diagonal cell "mismatch", elsewhere "vertical delta is +1" -- the two bits name the
                        // step (=, X, I, D) without a look at the sequences
                        for (int w = 0; w < W; ++w) { tdg[u][w] = dg[w]; tpv[u][w] = (dg[w] & ~eq0[w]) | (~dg[w] & pvn[w]); }
                    }
#pragma unroll
                    for (int w = 0; w < W; ++w) pv[w] = act ? pvn[w] : pv[w];
                    if (G == 32 && wr) hist |= (((t2 >> 7) & 1u) | ((t2 >> 14) & 2u)) << (2 * (j & 15));
                }
                if (TRACE) {
                    // Trace layout [stripe][block of 4 columns][lane][word][column] x (diag, up): the four columns of a
                    // lane's word share a 32-byte sector, so a diagonal walk reads a quarter of the sectors.  A lane's share
                    // of the block is 32 W contiguous bytes; stored straight from registers that is 16-byte pieces at a
                    // 32 W-byte stride (measured: DP 1.7x slower).  So the block goes through shared memory -- each lane
                    // parks its 2 W chunks of 16 bytes (swizzled by lane: conflict-free per quarter warp), then the warp
                    // copies the 1024 W bytes out in rows of 512 contiguous bytes.
                    constexpr int NC = 2 * W;                               // 16-byte chunks per lane
                    uint4 *dst = reinterpret_cast<uint4 *>(a.trace + tbase) + (((int64_t)s * n_blk + (j4 >> 2)) * 32) * NC;
                    ISOF_ASSERT(tbase + ((((int64_t)s * n_blk + (j4 >> 2)) * 32) + 32) * NC * 2 <= a.trace_cap, CHK_TRACE_STORE);
#pragma unroll
                    for (int w = 0; w < W; ++w) {
                        const int c0 = 2 * w, c1 = 2 * w + 1;
                        const int p0 = (NC & (NC - 1)) == 0 ? (c0 ^ (lane & (NC - 1))) : (c0 + lane) % NC;
                        const int p1 = (NC & (NC - 1)) == 0 ? (c1 ^ (lane & (NC - 1))) : (c1 + lane) % NC;
                        stg[lane * NC + p0] = make_uint4(tdg[0][w], tpv[0][w], tdg[1][w], tpv[1][w]);
                        stg[lane * NC + p1] = make_uint4(tdg[2][w], tpv[2][w], tdg[3][w], tpv[3][w]);
                    }
                    __syncwarp();
#pragma unroll
                    for (int rw = 0; rw < NC; ++rw) {
                        const int p = rw * 32 + lane, lsrc = p / NC, cp = p % NC;
                        const int cl = (NC & (NC - 1)) == 0 ? (cp ^ (lsrc & (NC - 1))) : (cp + NC - lsrc % NC) % NC;
                        dst[lsrc * NC + cl] = stg[p];
                    }
                    __syncwarp();
                }
            }
            if (G == 32 && wr) {
                if (lane == 31) lb[j0 >> 4] = hist;                    // bottom row of the stripe: the next stripe's carry-in
                hist = 0u;
            }
        }
        int d = 0;
#pragma unroll
        for (int w = 0; w < W; ++w) {
            const int nw = min(32, max(0, nv - 32 * w));
            const uint32_t rows = nw >= 32 ? ~0u : ((1u << nw) - 1u);
            d += __popc(pv[w] & rows) - __popc(mv[w] & rows);
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) d += __shfl_xor_sync(FULL, d, o);
        acc += d;
        __syncwarp();
    }
    if (valid && gl == 0) {
        int dist = m + acc;                                            // D(0,m) = m, then down the last column
        if (cut || (a.k >= 0 && dist > a.k)) dist = -1;
        a.dist[pair] = dist;
        if (TRACE) a.desc[pair] = EdDesc{tbase, n_blk, (slot * G) | (c << 8)};
    }
}

// ---------------------------------------------------------------------------------------------------------------
// banded classes
// ---------------------------------------------------------------------------------------------------------------
// Match vectors of every sequence of the pool, once per call: eqg[y * tw + (off[s] >> 5) + s + r] = rows 32r .. 32r+31 of
// sequence s that hold symbol y (at most 8 symbols).  One warp per sequence.
__global__ void ed_eqglobal_kernel(const uint8_t *__restrict__ codes, const int64_t *__restrict__ off, int32_t n_seqs,
                                   const EdLayout *lay, uint32_t *__restrict__ eqg, int64_t tw)
{
    if (lay->n_sym > 8) return;
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t sq = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; sq < n_seqs; sq += nwarps) {
        const int64_t b = off[sq], len = off[sq + 1] - b, o = (b >> 5) + sq;
        for (int64_t r = lane; r < ((len + 31) >> 5); r += 32) {
            uint32_t e[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            const int nv = (int)min((int64_t)32, len - 32 * r);
            for (int i = 0; i < nv; ++i) {
                const uint32_t cd = codes[b + 32 * r + i];
#pragma unroll
                for (int y = 0; y < 8; ++y) e[y] |= (cd == (uint32_t)y ? 1u : 0u) << i;
            }
#pragma unroll
            for (int y = 0; y < 8; ++y) eqg[y * tw + o + r] = e[y];
        }
    }
}

// One pair on a ring of G lanes, one 32-row word each: the window [ (j-k) >> 5, (j+k) >> 5 ] of query words slides down as
// the column j advances; word r lives in lane r mod G.  Cells outside the band count as "large" (a word enters with Pv = all
// ones, the chain's first word takes +1 from above): values never fall below the true ones and are exact wherever the
// true distance is <= k.  The score is carried along the window's bottom row.
template <int G>
__device__ __forceinline__ void ed_run_banded(const EdArgs &a, const int t, const int c, const int lane)
{
    constexpr int PPT = 32 / G;
    constexpr uint32_t GM = G == 32 ? 0xffffffffu : ((1u << G) - 1u);
    const int gl = lane % G, slot = lane / G, gbase = slot * G;
    const EdLayout *lay = a.lay;
    const int sidx = lay->cls_start[c] + (t - lay->task_start[c]) * PPT + slot;
    const bool valid = sidx < lay->cls_start[c + 1];
    int pair = -1, n = 0, m = 0;
    int64_t tb = 0, oq = 0;
    if (valid) {
        pair = a.order[sidx];
        n = a.pn[pair];
        m = a.pm[pair];
        const int qa = a.pa[pair];
        oq = (a.off[qa] >> 5) + qa;
        tb = a.off[a.pb[pair]];
    }
    const int k = a.k;
    const int nw = (n + 31) >> 5;
    const int m_max = (int)__reduce_max_sync(FULL, (unsigned)m);
    uint32_t pv = ~0u, mv = 0u;
    int whi_prev = -1, score = 0;
    for (int j = 0; j < m_max; ++j) {
        const bool act = j < m;                                        // uniform inside a group
        const int wlo = max(j - k, 0) >> 5, whi = min(nw - 1, (j + k) >> 5);
        const int r = wlo + ((gl - wlo) & (G - 1));                    // the window's word that lives in this lane
        const bool in = act && r <= whi;
        if (in && r > whi_prev) { pv = ~0u; mv = 0u; }                 // the word enters the band
        const uint32_t ch = act ? a.codes[tb + j] : 0u;
        const uint32_t eq = in ? a.eqg[ch * a.eq_tw + oq + r] : 0u;
        const int p = r - wlo;                                         // position in the chain, 0 = top of the window
        const uint32_t x = eq & pv;
        uint32_t sum = x + pv;
        const bool top = r == whi;
        const unsigned gm = __ballot_sync(FULL, in && !top && sum < x);
        const unsigned pk = __ballot_sync(FULL, in && !top && sum == 0xffffffffu);
        const int s0 = wlo & (G - 1);                                  // ring -> chain order: rotate the group's votes
        const uint32_t gs = (gm >> gbase) & GM, ps = (pk >> gbase) & GM;
        const uint32_t gr = G == 32 ? __funnelshift_r(gs, gs, s0) : (((gs >> s0) | (gs << (G - s0))) & GM);
        const uint32_t pr = G == 32 ? __funnelshift_r(ps, ps, s0) : (((ps >> s0) | (ps << (G - s0))) & GM);
        const uint32_t cy = ((pr | gr) + gr) ^ pr;
        sum += (cy >> p) & 1u;
        const uint32_t xh = (sum ^ pv) | eq;
        const uint32_t ph = mv | ~(xh | pv);
        const uint32_t mh = pv & xh;
        const uint32_t t2 = (ph >> 31) | ((mh >> 31) << 1);
        const uint32_t below = __shfl_sync(FULL, t2, gbase + ((gl - 1) & (G - 1)));
        const uint32_t tbits = __shfl_sync(FULL, t2, gbase + (whi & (G - 1)));
        const uint32_t up = p == 0 ? 1u : below;                       // above the window: +1 per column
        const uint32_t ph2 = (ph << 1) | (up & 1u), mh2 = (mh << 1) | (up >> 1);
        const uint32_t xv = eq | mv;
        if (in) {
            pv = mh2 | ~(xv | ph2);
            mv = ph2 & xv;
        }
        if (act) {
            score += 32 * (whi - whi_prev) + (int)(tbits & 1u) - (int)(tbits >> 1);
            whi_prev = whi;
        }
    }
    // the last word holds row n-1 (m + k >= n); rows of that word beyond n are padding that never matches
    const int nvl = n - 32 * (nw - 1);
    const uint32_t padm = nvl >= 32 ? 0u : ~((1u << nvl) - 1u);
    const int extra = __popc(pv & padm) - __popc(mv & padm);
    const int ex = __shfl_sync(FULL, extra, gbase + ((nw - 1) & (G - 1)));
    if (valid && gl == 0) {
        int dist = score - ex;
        if (dist > k) dist = -1;
        a.dist[pair] = dist;
    }
}

template <bool TRACE>
__global__ void __launch_bounds__(ED_WARPS * 32) ed_dp_kernel(const EdArgs a)
{
    extern __shared__ uint4 ed_smem4[];
    uint32_t *ed_smem = reinterpret_cast<uint32_t *>(ed_smem4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = a.t0 + blockIdx.x * (blockDim.x >> 5) + warp;
    if (t >= a.t1) return;
    uint32_t *peq = ed_smem + (size_t)warp * (a.eq_words + a.lb_words + a.stg_words);   // Eq sized for the widest W; all 16-byte aligned
    uint32_t *lb = peq + a.eq_words;
    uint4 *stg = reinterpret_cast<uint4 *>(lb + a.lb_words);
    const int c = ed_task_class(a.lay, t);
    // class id -> (G, W); the traced kernel only ever sees W <= ED_MAXW_PATH
#define ED_CASE(G_, W_)                                                                                        \
    case ((5 - (G_ == 32 ? 5 : G_ == 16 ? 4 : G_ == 8 ? 3 : G_ == 4 ? 2 : G_ == 2 ? 1 : 0)) * 8 + (8 - W_)):      \
        if constexpr (!TRACE || W_ <= ED_MAXW_PATH) ed_run_task<G_, W_, TRACE>(a, t, c, peq, lb, stg, lane);        \
        break;
#define ED_CASES(G_) ED_CASE(G_, 8) ED_CASE(G_, 7) ED_CASE(G_, 6) ED_CASE(G_, 5) ED_CASE(G_, 4) ED_CASE(G_, 3)
    switch (c) {
        ED_CASES(32) ED_CASES(16) ED_CASES(8) ED_CASES(4) ED_CASES(2)
        ED_CASES(1) ED_CASE(1, 2) ED_CASE(1, 1)
    case 0 * 8 + 7: if constexpr (!TRACE) ed_run_banded<32>(a, t, c, lane); break;
    case 1 * 8 + 7: if constexpr (!TRACE) ed_run_banded<16>(a, t, c, lane); break;
    case 2 * 8 + 7: if constexpr (!TRACE) ed_run_banded<8>(a, t, c, lane); break;
    case 3 * 8 + 7: if constexpr (!TRACE) ed_run_banded<4>(a, t, c, lane); break;
    case 4 * 8 + 7: if constexpr (!TRACE) ed_run_banded<2>(a, t, c, lane); break;
    default: break;
    }
#undef ED_CASES
#undef ED_CASE
}

// ---------------------------------------------------------------------------------------------------------------
// traceback: one thread per pair walks the trace ONCE and leaves the path as 2-bit steps (16 per word, walk order) in a
// small per-pair scratch plus its run count; after the scan of the counts ed_runs_kernel turns the steps into runs at
// their final place (alignment order).  The walk reads a trace that does not fit L2, so it is not repeated -- and a
// chunked call needs its DP only once.
// ---------------------------------------------------------------------------------------------------------------
__global__ void ed_traceback_kernel(const EdArgs a, const int64_t sidx0, const int64_t sidx1, int64_t *__restrict__ nruns64,
                                    const int64_t *__restrict__ ops_off, uint32_t *__restrict__ ops, int4 *__restrict__ hdr)
{
    // one thread per pair of the chunk, in sorted order: every lane of a warp walks (the trace is addressed through the
    // pair's descriptor, not through its task)
    const int64_t sidx = sidx0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (sidx >= sidx1) return;
    const int pair = a.order[sidx];
    const int c = a.desc[pair].lane0_cls >> 8;
    if (a.dist[pair] < 0) {
        nruns64[pair] = 0;
        hdr[pair] = make_int4(0, 0, 0, 0);
        return;
    }
    const EdDesc d = a.desc[pair];
    const int W = ed_w(c), lane0 = d.lane0_cls & 255;
    const int rows_stripe = (32 * W) << ed_lgg(c);
    int i = a.pn[pair], j = a.pm[pair];
    uint32_t *out = ops + ops_off[pair];
    int64_t count = 0;
    int steps = 0;
    uint32_t cur_op = 0xffu, acc = 0u;
    auto step = [&](uint32_t code) {                                   // 0 '=', 1 'X', 2 'I', 3 'D'
        count += (code != cur_op) ? 1 : 0;
        cur_op = code;
        acc |= code << (2 * (steps & 15));
        if ((++steps & 15) == 0) { out[(steps >> 4) - 1] = acc; acc = 0u; }
    };
    // A plain dependent walk: fetching the next cells along the diagonal ahead of time (look-ahead 1 -> 8) was measured
    // slower at every length, related or unrelated pairs (100 bp: 7.0 vs 2.9 ms per 2 M pairs) -- with this many walkers in
    // flight the kernel is bound by issue slots and occupancy (38 vs 64 registers), not by one walker's latency.  Calls of
    // few long pairs, where latency does bind, take ed_traceback_warp_kernel.
    const bool multi = ed_lgg(c) == 5;
    while (i > 0 && j > 0) {
        const int r = i - 1, jc = j - 1;
        const int s = multi ? r / rows_stripe : 0, wi = (r - s * rows_stripe) >> 5;           // word of the stripe
        const int64_t ti = d.tbase + ((((int64_t)s * d.n_blk + (jc >> 2)) * 32 + lane0 + wi / W) * W + wi % W) * 4 + (jc & 3);
        ISOF_ASSERT(ti < a.trace_cap, CHK_TRACE_LOAD);
        const uint2 w = a.trace[ti];
        const uint32_t bit = 1u << (r & 31);
        if (w.x & bit) {
            step((w.y & bit) ? 1u : 0u);
            --i; --j;
        } else if (w.y & bit) {
            step(2u);
            --i;
        } else {
            step(3u);
            --j;
        }
    }
    if (steps & 15) out[steps >> 4] = acc;
    // the rest of the first row / column is one run
    if (i > 0) count += (cur_op != 2u) ? 1 : 0;
    if (j > 0) count += (cur_op != 3u) ? 1 : 0;
    nruns64[pair] = count;
    hdr[pair] = make_int4(steps, i, j, 0);                           // .w = 0: 2-bit steps
}

// One WARP per pair: lane k fetches the cell k steps down the diagonal, a ballot of the "diagonal is optimal" bits gives the
// length of the diagonal stretch, the first lane that is off it holds the gap step -- one trip to DRAM covers up to 32 steps
// of the walk instead of one to eight.  The look-ahead adapts as in the thread walk.  Steps are left as one BYTE each (lane
// k writes step k of the stretch).  Taken when the chunk holds so few pairs beyond 128 rows (at most 128 x SMs) that one
// thread each would leave the GPU waiting on single walkers (a 5 kb x 5 kb pair is 10 k dependent steps).
__global__ void __launch_bounds__(128) ed_traceback_warp_kernel(const EdArgs a, const int64_t sidx0, const int64_t sidx1,
                                                                 int64_t *__restrict__ nruns64, const int64_t *__restrict__ ops_off,
                                                                 uint32_t *__restrict__ ops, int4 *__restrict__ hdr)
{
    const int64_t sidx = sidx0 + (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (sidx >= sidx1) return;
    const int pair = a.order[sidx];
    if (a.dist[pair] < 0) {
        if (lane == 0) {
            nruns64[pair] = 0;
            hdr[pair] = make_int4(0, 0, 0, 1);
        }
        return;
    }
    const EdDesc d = a.desc[pair];
    const int c = d.lane0_cls >> 8;
    const int W = ed_w(c), lane0 = d.lane0_cls & 255;
    const int rows_stripe = (32 * W) << ed_lgg(c);
    const bool multi = ed_lgg(c) == 5;
    int i = a.pn[pair], j = a.pm[pair];
    uint8_t *out = reinterpret_cast<uint8_t *>(ops + ops_off[pair]);
    int steps = 0, count = 0, kcur = 1;
    uint32_t cur = 0xffu;
    while (i > 0 && j > 0) {
        const int kmax = min(kcur, min(i, j));
        bool dg = false, up = false, eq = false;
        if (lane < kmax) {
            const int r = i - 1 - lane, jc = j - 1 - lane;
            const int s = multi ? r / rows_stripe : 0, wi = (r - s * rows_stripe) >> 5;
            const int64_t ti = d.tbase + ((((int64_t)s * d.n_blk + (jc >> 2)) * 32 + lane0 + wi / W) * W + wi % W) * 4 + (jc & 3);
            ISOF_ASSERT(ti < a.trace_cap, CHK_TRACE_LOAD);
            const uint2 w = a.trace[ti];
            const uint32_t bit = 1u << (r & 31);
            dg = (w.x & bit) != 0;
            up = (w.y & bit) != 0;
            eq = !up;                                              // on a diagonal cell the second bit says "mismatch"
        }
        const unsigned D = __ballot_sync(FULL, dg), U = __ballot_sync(FULL, up);
        const int nd = min(kmax, D == FULL ? 32 : __ffs(~D) - 1);       // diagonal steps of this stretch
        const bool gap = nd < kmax;                                      // lane nd holds a gap step
        const bool gup = gap && ((U >> nd) & 1u);
        const int nsteps = nd + (gap ? 1 : 0);
        const uint32_t code = lane < nd ? (eq ? 0u : 1u) : (up ? 2u : 3u);
        if (lane < nsteps) out[steps + lane] = (uint8_t)code;
        uint32_t prev = __shfl_up_sync(FULL, code, 1);
        if (lane == 0) prev = cur;
        count += __popc(__ballot_sync(FULL, lane < nsteps && code != prev));
        cur = __shfl_sync(FULL, code, nsteps - 1);
        steps += nsteps;
        i -= nd + (gup ? 1 : 0);
        j -= nd + ((gap && !gup) ? 1 : 0);
        kcur = gap ? 1 : min(32, kcur * 2);
    }
    if (lane == 0) {
        if (i > 0) count += (cur != 2u) ? 1 : 0;
        if (j > 0) count += (cur != 3u) ? 1 : 0;
        nruns64[pair] = count;
        hdr[pair] = make_int4(steps, i, j, 1);                           // .w = 1: a byte per step
    }
}

__global__ void ed_runs_kernel(int64_t P, const int64_t *__restrict__ ops_off, const uint32_t *__restrict__ ops,
                               const int4 *__restrict__ hdr, const int64_t *__restrict__ cigar_off, uint32_t *__restrict__ cigar)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const int4 h = hdr[p];
    const uint32_t *in = ops + ops_off[p];
    int64_t wpos = cigar_off[p + 1] - 1;
    uint32_t cur_op = 0xffu, cur_len = 0, word = 0;
    // Runs go out from the pair's last slot downwards.  One 4-byte store per run is one L2 sector operation each (ncu: 8x the
    // sectors the bytes need), so four runs are gathered and an aligned group of slots leaves as one 16-byte store; the
    // ragged groups at either end of the pair's range (shared with its neighbours) are written run by run.
    const int64_t lo = cigar_off[p];
    uint32_t g0 = 0, g1 = 0, g2 = 0, g3 = 0;
    int filled = 0;
    const bool vec = (reinterpret_cast<uintptr_t>(cigar) & 15) == 0;     // a caller's device buffer need not be 16-byte aligned
    auto put = [&](uint32_t run) {
        ISOF_ASSERT(wpos >= lo, CHK_CIGAR_COPY);
        const int idx = (int)(wpos & 3);
        g0 = idx == 0 ? run : g0; g1 = idx == 1 ? run : g1; g2 = idx == 2 ? run : g2; g3 = idx == 3 ? run : g3;
        ++filled;
        if (idx == 0) {
            if (filled == 4 && vec) *reinterpret_cast<uint4 *>(cigar + wpos) = make_uint4(g0, g1, g2, g3);
            else {                                                     // the topmost group of the pair, not full
                cigar[wpos] = g0;
                if (filled > 1) cigar[wpos + 1] = g1;
                if (filled > 2) cigar[wpos + 2] = g2;
                if (filled > 3) cigar[wpos + 3] = g3;
            }
            filled = 0;
        }
        --wpos;
    };
    auto emit = [&](uint32_t op, uint32_t len) {
        if (op == cur_op) { cur_len += len; return; }
        if (cur_len) put((cur_len << 4) | cur_op);
        cur_op = op;
        cur_len = len;
    };
    for (int st = 0; st < h.x; ++st) {
        uint32_t code;
        if (h.w) {
            if ((st & 3) == 0) word = in[st >> 2];
            code = (word >> (8 * (st & 3))) & 3u;
        } else {
            if ((st & 15) == 0) word = in[st >> 4];
            code = (word >> (2 * (st & 15))) & 3u;
        }
        emit(code == 0u ? (uint32_t)ISOF_CIGAR_EQ : code == 1u ? (uint32_t)ISOF_CIGAR_X : code == 2u ? (uint32_t)ISOF_CIGAR_I
                                                                                                      : (uint32_t)ISOF_CIGAR_D, 1u);
    }
    if (h.y > 0) emit(ISOF_CIGAR_I, (uint32_t)h.y);
    if (h.z > 0) emit(ISOF_CIGAR_D, (uint32_t)h.z);
    emit(0xfeu, 0);                                                    // flush
    // the lowest group of the pair, not reached down to its aligned start: slots wpos+1 .. wpos+filled
    if (filled > 0) {
        const int top = (int)((wpos + filled) & 3);                    // index of the group's highest filled slot
        if (top == 3) cigar[wpos + filled] = g3;
        if (top >= 2 && filled > top - 2) cigar[(wpos + filled) - (top - 2)] = g2;
        if (top >= 1 && filled > top - 1) cigar[(wpos + filled) - (top - 1)] = g1;
    }
}

int ed_launch_dp(isof_ctx *ctx, EdArgs &a, bool trace, size_t smem_warp, int wpb)
{
    const int blocks = (a.t1 - a.t0 + wpb - 1) / wpb;
    if (blocks <= 0) return ISOF_OK;
    LaunchTimer lt(ctx, SLOT_DP);
    if (trace) ed_dp_kernel<true><<<blocks, wpb * 32, smem_warp * wpb, ctx->stream>>>(a);
    else ed_dp_kernel<false><<<blocks, wpb * 32, smem_warp * wpb, ctx->stream>>>(a);
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
    TRY(dev_reserve(ctx, B[1], 512));                                        // present[8] | cnt[ED_NC] at +32 B | lut[256] at +256 B
    TRY(dev_reserve(ctx, B[2], sizeof(EdLayout)));
    TRY(dev_reserve(ctx, B[3], sizeof(int32_t) * (size_t)P));                // pn
    TRY(dev_reserve(ctx, B[4], sizeof(int32_t) * (size_t)P));                // pm
    TRY(dev_reserve(ctx, B[5], sizeof(unsigned long long) * (size_t)P * 2)); // keys in | out
    TRY(dev_reserve(ctx, B[6], sizeof(int32_t) * (size_t)P * 2));            // idx in | order
    TRY(dev_reserve(ctx, B[7], sizeof(int64_t) * (size_t)(P + 1) * 2));      // task_words | task_off
    uint8_t *codes = (uint8_t *)B[0].p;
    unsigned int *present = (unsigned int *)B[1].p, *cnt = present + 8;
    uint8_t *lut = (uint8_t *)(present + 64);
    EdLayout *lay = (EdLayout *)B[2].p;
    int32_t *pn = (int32_t *)B[3].p, *pm = (int32_t *)B[4].p;
    unsigned long long *keys = (unsigned long long *)B[5].p, *keys_out = keys + P;
    int32_t *idx = (int32_t *)B[6].p, *order = idx + P;
    int64_t *task_words = (int64_t *)B[7].p, *task_off = task_words + (P + 1);
    int64_t *ops_words = nullptr, *ops_off = nullptr;
    if (want_path) {
        TRY(dev_reserve(ctx, B[13], sizeof(int64_t) * (size_t)(P + 1) * 2));  // step words per pair | their scan
        ops_words = (int64_t *)B[13].p;
        ops_off = ops_words + (P + 1);
        CUDA_TRY(ctx, cudaMemsetAsync(ops_words + P, 0, sizeof(int64_t), st));
    }
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
        ed_plan_kernel<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(d_off, n_seqs, d_pa, d_pb, P, pn, pm, keys, idx, cnt, lay, ops_words,
                                                                    want_path ? 1 : 0, k);
    }
    const bool may_band = k >= 0 && !want_path;
    const int64_t eq_tw = (n_bases >> 5) + n_seqs + 1;
    if (may_band) {
        TRY(dev_reserve(ctx, ctx->b_ed_eqg, sizeof(uint32_t) * 8 * (size_t)eq_tw));
        LaunchTimer lt(ctx, SLOT_OTHER);
        const int blocks = (int)std::min<int64_t>(((int64_t)n_seqs * 32 + 255) / 256 + 1, (int64_t)ctx->sm_count * 16);
        ed_eqglobal_kernel<<<blocks, 256, 0, st>>>(codes, d_off, n_seqs, lay, (uint32_t *)ctx->b_ed_eqg.p, eq_tw);
    }
    CUDA_TRY(ctx, cudaGetLastError());
    size_t tmp_sort = 0, tmp_scan = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, keys, keys_out, idx, order, (int)P, 0, 62, st);
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan, task_words, task_off, (int)(P + 1), st);
    TRY(dev_reserve(ctx, B[8], std::max(tmp_sort, tmp_scan)));
    {
        LaunchTimer lt(ctx, SLOT_OTHER);
        CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(B[8].p, tmp_sort, keys, keys_out, idx, order, (int)P, 0, 62, st));
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
    if (want_path) {
        LaunchTimer lt(ctx, SLOT_OTHER);
        CUDA_TRY(ctx, cub::DeviceScan::ExclusiveSum(B[8].p, tmp_scan, ops_words, ops_off, (int)(P + 1), st));
    }
    CUDA_TRY(ctx, cudaGetLastError());
    // ---- one read-back: layout + trace total
    EdLayout *h_lay = (EdLayout *)ctx->pinned;
    int64_t *h_total = (int64_t *)((char *)ctx->pinned + 1024);
    CUDA_TRY(ctx, cudaMemcpyAsync(h_lay, lay, sizeof(EdLayout), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(h_total, task_off + P, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    if (want_path) CUDA_TRY(ctx, cudaMemcpyAsync(h_total + 1, ops_off + P, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    if (h_lay->err & 1) return isof_fail(ctx, ISOF_ERR_ARG, "pair index outside the sequence pool");
    if (h_lay->err & 2) return isof_fail(ctx, ISOF_ERR_ARG, "sequence longer than 2^31-1");
    const int T = h_lay->task_start[ED_NC];
    const int n_sym = std::max(h_lay->n_sym, 1);
    const int lb_words = h_lay->max_m_multi > 0 ? (((h_lay->max_m_multi + 15) / 16 + 1 + 3) & ~3) : 0;
    const int64_t total_words = *h_total, total_ops = want_path ? h_total[1] : 0;
    const int stg_words = want_path ? 1024 : 0;                            // 1024 W bytes of trace staging per warp, W <= 4
    const int eq_words = n_sym * 32 * ed_max_w(h_lay->n_sym, want_path);      // Eq[symbol][lane][word] at the widest W
    const size_t smem_warp = ((size_t)eq_words + (size_t)lb_words + (size_t)stg_words) * sizeof(uint32_t);
    const int wpb = (int)std::max<size_t>(1, std::min<size_t>(ED_WARPS, (200 * 1024) / smem_warp));
    const size_t smem = smem_warp * wpb;
    if (smem > 200 * 1024)
        return isof_fail(ctx, ISOF_ERR_NOMEM, "edit distance: %d symbols and a %d-column multi-stripe target need %zu bytes of "
                         "shared memory per warp", n_sym, h_lay->max_m_multi, smem_warp);
    if (smem > 48 * 1024) {
        CUDA_TRY(ctx, cudaFuncSetAttribute(ed_dp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(ctx, cudaFuncSetAttribute(ed_dp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    EdArgs a;
    a.codes = codes; a.off = d_off; a.pa = d_pa; a.pb = d_pb; a.order = order; a.pn = pn; a.pm = pm; a.lay = lay;
    a.task_off = task_off; a.trace = nullptr; a.trace_base = 0; a.trace_cap = 0; a.t0 = 0; a.t1 = T; a.n_sym = n_sym; a.eq_words = eq_words; a.lb_words = lb_words; a.stg_words = stg_words;
    a.dist = d_dist; a.desc = nullptr; a.k = k;
    a.eqg = may_band ? (const uint32_t *)ctx->b_ed_eqg.p : nullptr; a.eq_tw = eq_tw;
    if (!want_path) return ed_launch_dp(ctx, a, false, smem_warp, wpb);

    // ---- path: trace chunks
    TRY(dev_reserve(ctx, B[9], sizeof(EdDesc) * (size_t)P));
    TRY(dev_reserve(ctx, B[10], sizeof(int64_t) * (size_t)(P + 1)));
    TRY(dev_reserve(ctx, B[14], sizeof(uint32_t) * (size_t)std::max<int64_t>(total_ops, 1)));
    TRY(dev_reserve(ctx, B[15], sizeof(int4) * (size_t)P));
    a.desc = (EdDesc *)B[9].p;
    int64_t *nruns64 = (int64_t *)B[10].p;
    uint32_t *ops = (uint32_t *)B[14].p;
    int4 *hdr = (int4 *)B[15].p;
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
    for (size_t x = 0; x + 1 < cuts.size(); ++x) {                     // per chunk: DP with the trace, then the walk
        a.t0 = cuts[x];
        a.t1 = cuts[x + 1];
        a.trace_base = chunk_base[x];
        TRY(ed_launch_dp(ctx, a, true, smem_warp, wpb));
        auto first_pair = [&](int t) -> int64_t {                      // sorted position of a task's first pair
            if (t >= T) return P;
            int c = 0;
            while (c + 1 < ED_NC && t >= h_lay->task_start[c + 1]) ++c;
            return h_lay->cls_start[c] + (int64_t)(t - h_lay->task_start[c]) * (32 >> ed_lgg(c));
        };
        const int64_t s0 = first_pair(a.t0), s1 = first_pair(a.t1);
        // pairs beyond 128 rows (classes 0-9, first in the sorted order): a warp each while that still leaves the GPU
        // enough walkers in flight; the rest one thread each
        int64_t sw = std::min<int64_t>(s1, std::max<int64_t>(s0, h_lay->cls_start[10]));
        if (ctx->ed_tb_warp == 0 || (ctx->ed_tb_warp == 2 && sw - s0 > (int64_t)ctx->sm_count * 128)) sw = s0;
        LaunchTimer lt(ctx, SLOT_TB);
        if (sw > s0)
            ed_traceback_warp_kernel<<<(unsigned)((sw - s0 + 3) / 4), 128, 0, st>>>(a, s0, sw, nruns64, ops_off, ops, hdr);
        if (s1 > sw)
            ed_traceback_kernel<<<(unsigned)((s1 - sw + 127) / 128), 128, 0, st>>>(a, sw, s1, nruns64, ops_off, ops, hdr);
        CUDA_TRY(ctx, cudaGetLastError());
    }
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
    {
        LaunchTimer lt(ctx, SLOT_TB);
        ed_runs_kernel<<<(unsigned)((P + 127) / 128), 128, 0, st>>>(P, ops_off, ops, hdr, d_cigar_off, d_cigar);
        CUDA_TRY(ctx, cudaGetLastError());
    }
    return ISOF_OK;
}
