// sg_align.cu -- kernel (b): batched semi-global affine-gap alignment with full traceback,
// CIGAR-compatible with the parasail.sg_trace_scan_16/_32 call behind
// consensus.parasail_alignment (reference modules/consensus.py:55-67).
//
// DP kernel (sg_dp_kernel): persistent warps pull tasks (two consecutive pairs) from an atomic counter.
// Anti-diagonal wavefront per warp: the query (rows) is cut into stripes of 256 rows; lane l owns 8
// consecutive rows and, at step t, the two columns 2(t-l), 2(t-l)+1.  Per step a lane receives H/F of the
// row above from lane l-1 (shuffles), keeps H(i,j-1), E(i,j-1) of its rows in registers and emits one
// 32-bit trace word (8 rows x 4 bits) per column and alignment.
// Scores are held as scale*score + tag: the low bits carry the provenance of a value so that the DPX max
// instructions (VIADDMNMX / VIMNMX3) resolve every tie exactly like the restated library (DIAG > F > E;
// gap extension wins a tie against gap opening) and the trace nibble falls out of the winning operands'
// tags without a single compare instruction.  32-bit path (16*score + 4 tag bits):
//     E candidates:  open -> ..0100   extend -> ..0101      (bit0 = E came from extension)
//     F candidates:  open -> ..1000   extend -> ..1010      (bit1 = F came from extension)
//     diagonal:                ..1100
//     nibble = [src1 src0 f_ext e_ext],  src: 3 diag, 2 F (consumes query, 'I'), 1 E (consumes ref, 'D')
// The packed path (two alignments per warp in 16-bit halves, 4*score + 2 tag bits) is described at
// dp_pair16.  Trace layout per pair: word[((stripe*T + t)*32 + lane)*2 + col], T = ceil(m/2)+31 -> every
// step is one coalesced 256 B store.  The bottom row of a stripe (H,F per column) goes through a per-warp
// global line buffer.
//
// Traceback kernel: one thread per pair walks the trace from the end cell and writes the CIGAR
// runs backwards into the tail of the pair's own trace region, so the runs end up in alignment
// order and a coalesced copy moves them to the caller's buffer.
#include "common.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/block/block_scan.cuh>
#include <climits>
#include <type_traits>

namespace {

constexpr int R = 8;                 // rows per lane
constexpr int SROWS = 32 * R;        // rows per stripe
constexpr int TAG_BITS = 4;
constexpr int NEGV = -(1 << 29);     // "minus infinity", a clean multiple of 16
constexpr int CODE_SHIFT = 12;       // base codes live at bits 12..14 of a register
constexpr int TRACE_SLACK = 64;      // spare words at the end of every trace region
constexpr int DP_WARPS = 1;          // warps per CTA of the DP kernel: single-warp CTAs leave the SM one by one, so the
                                     // next chunk's CTAs fill every freed slot at once (8-warp CTAs held their registers
                                     // until the slowest of 8 tasks ended: -3.5 % on C2)
constexpr int DP_CTAS_PER_SM = 16;   // 128 registers x 32 threads x 16 = the whole register file

struct AlignParams {
    const uint8_t *cat;        // ASCII pool
    const uint8_t *codes;      // 0..3 ACGT (case-insensitive), 4 anything else
    const int64_t *seq_off;
    const int32_t *pa, *pb;
    const int32_t *pn, *pm;
    const uint8_t *pwild;
    const int64_t *ptrace;     // P+1 word offsets (global, exclusive scan)
    int64_t chunk_base;        // word offset of this chunk's first pair
    uint32_t *trace;
    int2 *bnd;
    int64_t bnd_stride;        // int2 elements per warp
    int32_t *score, *endq, *endr, *nruns;
    int c_eo, c_ee, c_fo, c_fe, c_sm, c_sx;
    // packed 16-bit path (two alignments per warp): 4*score + 2 tag bits per half
    uint32_t h_eo, h_ee, h_fo, h_fe, h_sm, h_sx;
    int h_max_min_len, h_max_max_len;      // eligibility: min(n,m) <= h_max_min_len and max(n,m) <= h_max_max_len
    uint8_t *playout;                      // per pair: columns per step of the path that wrote its trace (1 or PC)
    uint8_t *plastrr;                      // per pair: rows per lane used in the pair's LAST stripe (2,4,6,8)
    int64_t chunk_words;                   // words of this chunk's trace scratch (debug assertions)
    const int32_t *order;                  // per chunk: pair ids sorted by (packed-eligible first, reference index);
                                           // packed task t = (order[start+2t], order[start+2t+1]) -> partners share the
                                           // reference; the pairs behind the eligible ones are single 32-bit tasks
    const unsigned long long *n_elig;      // per chunk: number of packed-eligible pairs (device scalar)
};

__device__ __forceinline__ int base_code(uint8_t c)
{
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return 4;
    }
}

// warp per sequence: codes + "has a non-ACGT character" flag
__global__ void encode_kernel(const uint8_t *__restrict__ cat, const int64_t *__restrict__ off, int n_seqs,
                              uint8_t *__restrict__ codes, uint8_t *__restrict__ flags)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int s = warp; s < n_seqs; s += nwarps) {
        const int64_t b = off[s], e = off[s + 1];
        int wild = 0;
        for (int64_t x = b + lane; x < e; x += 32) {
            int c = base_code(cat[x]);
            codes[x] = (uint8_t)c;
            wild |= (c == 4);
        }
        wild = __any_sync(0xffffffffu, wild);
        if (lane == 0) flags[s] = (uint8_t)wild;
    }
}

// per pair: lengths, wildcard flag, trace words; misc[0]=error flags, misc[1]=max m, misc[2..3]=cells
__global__ void plan_kernel(const int64_t *__restrict__ off, int n_seqs, const uint8_t *__restrict__ flags,
                            const int32_t *__restrict__ pa, const int32_t *__restrict__ pb, int64_t P,
                            int32_t *__restrict__ pn, int32_t *__restrict__ pm, uint8_t *__restrict__ pwild,
                            int64_t *__restrict__ words, unsigned long long *__restrict__ misc, int h_max_min_len,
                            int h_max_max_len)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long cells = 0, padded = 0;
    int mm = 0;
    if (p < P) {
        int a = pa[p], b = pb[p];
        int n = 0, m = 0;
        int64_t w = TRACE_SLACK;
        if (a < 0 || a >= n_seqs || b < 0 || b >= n_seqs) atomicOr(&misc[0], 1ull);
        else {
            n = (int)(off[a + 1] - off[a]);
            m = (int)(off[b + 1] - off[b]);
            if (n <= 0 || m <= 0) { atomicOr(&misc[0], 2ull); n = m = 0; }
            else {
                int ns = (n + SROWS - 1) / SROWS;
                w = (int64_t)ns * (m + 64) * 32 + TRACE_SLACK;        // covers the 1- and 2-column layouts
                cells = (unsigned long long)n * m;
                // bit0: a wildcard is present (32-bit WILD path); bit1: eligible for the packed 16-bit path
                const int wild = flags[a] | flags[b];
                const int elig = !wild && min(n, m) <= h_max_min_len && max(n, m) <= h_max_max_len;
                // rows actually updated: full stripes + the compacted last stripe (2/4/6/8 rows per lane)
                const int need = n - (ns - 1) * SROWS;
                const int rr = need > 192 ? 8 : need > 128 ? 6 : need > 64 ? 4 : 2;
                padded = (unsigned long long)((ns - 1) * SROWS + (elig ? rr * 32 : SROWS)) * m;
                pwild[p] = (uint8_t)(wild | (elig << 1));
                mm = m;
            }
        }
        pn[p] = n; pm[p] = m;
        words[p] = w;
    }
    // block reduce the statistics
    for (int d = 16; d; d >>= 1) {
        cells += __shfl_down_sync(0xffffffffu, cells, d);
        padded += __shfl_down_sync(0xffffffffu, padded, d);
        mm = max(mm, __shfl_down_sync(0xffffffffu, mm, d));
    }
    if ((threadIdx.x & 31) == 0) {
        if (cells) { atomicAdd(&misc[2], cells); atomicAdd(&misc[3], padded); }
        if (mm) atomicMax(&misc[1], (unsigned long long)mm);
    }
    if (p == P) words[P] = 0;
}

// sort key of a pair inside its chunk: packed-eligible pairs first, then by reference index, so that the two
// alignments of a task share their reference whenever the batch allows it (query-profile path)
__global__ void task_key_kernel(const uint8_t *__restrict__ pwild, const int32_t *__restrict__ pb, int64_t start,
                                int64_t cnt, uint32_t *__restrict__ key, int32_t *__restrict__ val,
                                unsigned long long *__restrict__ n_elig)
{
    const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int el = 0;
    if (x < cnt) {
        const int64_t p = start + x;
        el = (pwild[p] & 2) ? 1 : 0;
        key[x] = (el ? 0u : 0x80000000u) | ((uint32_t)pb[p] & 0x7fffffffu);
        val[x] = (int32_t)p;
    }
    const unsigned m = __ballot_sync(0xffffffffu, el);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_elig, (unsigned long long)__popc(m));
}

// statistics only: cells of the pairs that take the packed path in chunk [start,end)
__global__ void packed_cells_kernel(const uint8_t *__restrict__ pwild, const int32_t *__restrict__ pn,
                                    const int32_t *__restrict__ pm, const int32_t *__restrict__ pb,
                                    const int32_t *__restrict__ order, const unsigned long long *__restrict__ n_elig,
                                    int64_t start, int64_t end, unsigned long long *__restrict__ out)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t xA = start + 2 * t, xB = xA + 1;
    unsigned long long c = 0, cp = 0;               // packed path / of those, the query-profile variant
    const int64_t pA = xA < end ? order[xA] : 0, pB = xB < end ? order[xB] : 0;
    if (xB < end && t < (int64_t)(*n_elig) / 2) {
        c = (unsigned long long)pn[pA] * pm[pA] + (unsigned long long)pn[pB] * pm[pB];
        if (pb[pA] == pb[pB]) cp = c;
    }
    for (int d = 16; d; d >>= 1) { c += __shfl_down_sync(0xffffffffu, c, d); cp += __shfl_down_sync(0xffffffffu, cp, d); }
    if ((threadIdx.x & 31) == 0 && c) { atomicAdd(out, c); if (cp) atomicAdd(out + 1, cp); }
}

// ---------------------------------------------------------------------------------------------
// Small batches (P <= SMALL_P pairs over <= SMALL_BASES bases; the per-pair seams of the reference issue their
// alignments one at a time): the whole set-up -- base codes, plan, trace offsets, task order -- is ONE
// single-block kernel, and the CIGAR scan/copy after the traceback another, instead of ~14 launches.
// ---------------------------------------------------------------------------------------------
constexpr int SMALL_P = 1024;
constexpr int64_t SMALL_BASES = 1 << 20;

__global__ void __launch_bounds__(1024)
small_plan_kernel(const uint8_t *__restrict__ cat, const int64_t *__restrict__ off, int n_seqs,
                  const int32_t *__restrict__ pa, const int32_t *__restrict__ pb, int P, uint8_t *__restrict__ codes,
                  uint8_t *__restrict__ flags, int32_t *__restrict__ pn, int32_t *__restrict__ pm,
                  uint8_t *__restrict__ pwild, int64_t *__restrict__ ptrace, unsigned long long *__restrict__ misc,
                  int h_max_min_len, int h_max_max_len, int32_t *__restrict__ order,
                  unsigned long long *__restrict__ n_elig, unsigned long long *__restrict__ ticket)
{
    typedef cub::BlockScan<long long, 1024> Scan;
    __shared__ typename Scan::TempStorage scan_tmp;
    __shared__ uint32_t s_key[SMALL_P];
    __shared__ unsigned long long s_stat[4];             // err, max m, cells, padded
    __shared__ int s_nelig;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid < 4) s_stat[tid] = 0ull;
    if (tid == 0) s_nelig = 0;
    // base codes + per-sequence wildcard flag (warp per sequence)
    for (int q = warp; q < n_seqs; q += 32) {
        const int64_t b = off[q], e = off[q + 1];
        int wild = 0;
        for (int64_t x = b + lane; x < e; x += 32) {
            const int c = base_code(cat[x]);
            codes[x] = (uint8_t)c;
            wild |= (c == 4);
        }
        wild = __any_sync(0xffffffffu, wild);
        if (lane == 0) flags[q] = (uint8_t)wild;
    }
    __syncthreads();
    // plan (same rules as plan_kernel)
    long long w = 0;
    uint32_t key = 0xffffffffu;
    if (tid < P) {
        const int a = pa[tid], b = pb[tid];
        int n = 0, m = 0;
        w = TRACE_SLACK;
        uint8_t pw = 0;
        if (a < 0 || a >= n_seqs || b < 0 || b >= n_seqs) atomicOr(&s_stat[0], 1ull);
        else {
            n = (int)(off[a + 1] - off[a]);
            m = (int)(off[b + 1] - off[b]);
            if (n <= 0 || m <= 0) { atomicOr(&s_stat[0], 2ull); n = m = 0; }
            else {
                const int ns = (n + SROWS - 1) / SROWS;
                w = (long long)ns * (m + 64) * 32 + TRACE_SLACK;
                const int wild = flags[a] | flags[b];
                const int elig = !wild && min(n, m) <= h_max_min_len && max(n, m) <= h_max_max_len;
                const int need = n - (ns - 1) * SROWS;
                const int rr = need > 192 ? 8 : need > 128 ? 6 : need > 64 ? 4 : 2;
                atomicAdd(&s_stat[2], (unsigned long long)n * m);
                atomicAdd(&s_stat[3], (unsigned long long)((ns - 1) * SROWS + (elig ? rr * 32 : SROWS)) * m);
                atomicMax(&s_stat[1], (unsigned long long)m);
                pw = (uint8_t)(wild | (elig << 1));
                if (elig) atomicAdd(&s_nelig, 1);
                key = (elig ? 0u : 0x80000000u) | ((uint32_t)b & 0x7fffffffu);
            }
        }
        pn[tid] = n; pm[tid] = m; pwild[tid] = pw;
    }
    s_key[tid] = key;
    long long excl = 0, total = 0;
    Scan(scan_tmp).ExclusiveSum(w, excl, total);         // includes a __syncthreads: s_key is complete afterwards
    if (tid < P) ptrace[tid] = excl;
    if (tid == 0) ptrace[P] = total;
    __syncthreads();
    // task order: stable sort by (eligible first, reference index) as a rank computation
    if (tid < P) {
        int rank = 0;
        for (int q = 0; q < P; ++q) {
            const uint32_t kq = s_key[q];
            rank += (kq < key) || (kq == key && q < tid);
        }
        order[rank] = tid;
    }
    if (tid == 0) {
        misc[0] = s_stat[0]; misc[1] = s_stat[1]; misc[2] = s_stat[2]; misc[3] = s_stat[3];
        misc[4] = 0ull; misc[5] = 0ull; misc[6] = 0ull;
        *n_elig = (unsigned long long)s_nelig;
        *ticket = 0ull;
    }
}

// chunk_start[c] = first pair whose trace offset is >= c*budget   (c = 0..n_chunks)
__global__ void chunk_bounds_kernel(const int64_t *__restrict__ ptrace, int64_t P, int64_t budget, int n_chunks,
                                    int64_t *__restrict__ chunk_start)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > n_chunks) return;
    if (c == n_chunks) { chunk_start[c] = P; return; }
    int64_t target = (int64_t)c * budget;
    int64_t lo = 0, hi = P;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (ptrace[mid] < target) lo = mid + 1; else hi = mid;
    }
    chunk_start[c] = lo;
}

// ---------------------------------------------------------------------------------------------
// DP
// ---------------------------------------------------------------------------------------------
// 32-bit path: one alignment per warp, 16*score + 4-bit tag, two columns per step, the step loop peeled
// into fill / steady / drain exactly like the packed path below (which documents the scheme).
// d = a * M + c on the FMA pipe (IMAD); inline PTX so the multiply is not strength-reduced to ALU-pipe shifts
__device__ __forceinline__ uint32_t mad_u32(const uint32_t a, const uint32_t mult, const uint32_t c)
{
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(mult), "r"(c));
    return d;
}

struct P32State {
    int qc[R], Hl[R], El[R];
    int Hbot[2], Fbot[2], rc_next[2];
    int2 b_next[2];
    int diag0;
};

template <bool WILD>
__device__ __forceinline__ void p32_column(P32State &S, const AlignParams &P, const int rc, int up, int fup, int diag,
                                           int &hbot, int &fbot, uint32_t &word, const uint4 *__restrict__ prof)
{
    // !WILD: the substitution scores of this lane's 8 rows against the column's base come from the lane's
    // query profile in shared memory (rc is then the plain base code 0..3); WILD keeps the arithmetic form.
    int scv[8];
    if (!WILD) {
        const uint4 v0 = prof[rc * 64], v1 = prof[rc * 64 + 32];
        scv[0] = (int)v0.x; scv[1] = (int)v0.y; scv[2] = (int)v0.z; scv[3] = (int)v0.w;
        scv[4] = (int)v1.x; scv[5] = (int)v1.y; scv[6] = (int)v1.z; scv[7] = (int)v1.w;
    }
    // Trace nibble of row r = [src1 src0 f e]: src is bits 3:2 of H's tag (H - Hc is that tag), e is bit 0 of E,
    // f is bit 1 of F (E's tag is 4|e, F's 8|2f, so (E | F) & 3 is [f e]).  Both are gathered with IMADs on the
    // FMA pipe; one LOP3 per column merges them.
    uint32_t acc_ef = 0u, acc_h = 0u;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int left = S.Hl[r];
        const int Ee = S.El[r] + P.c_ee;
        const int E = __viaddmax_s32(left, P.c_eo, Ee);
        const int Fe = fup + P.c_fe;
        const int F = __viaddmax_s32(up, P.c_fo, Fe);
        int sc;
        if (!WILD) sc = scv[r];
        else {
            const int y = S.qc[r] ^ rc ^ (3 << CODE_SHIFT);
            sc = __viaddmax_s32(y, P.c_sm, P.c_sx);
            if ((S.qc[r] | rc) & (4 << CODE_SHIFT)) sc = 12;
        }
        const int D = __viaddmax_s32(diag, sc, E);        // off the row-to-row chain (previous column only)
        const int H = max(D, F);
        const int Hc = H & ~15;
        const uint32_t ef = ((uint32_t)E | (uint32_t)F) & 3u;
        acc_ef = mad_u32(ef, 1u << (4 * r), acc_ef);
        acc_h = mad_u32((uint32_t)H, 1u << (4 * r), acc_h);
        acc_h = mad_u32((uint32_t)Hc, 0u - (1u << (4 * r)), acc_h);
        diag = left;
        S.Hl[r] = Hc;
        S.El[r] = E & ~15;
        up = Hc;
        fup = F & ~15;
    }
    hbot = up; fbot = fup;
    word = (acc_h & 0xccccccccu) | acc_ef;
}

template <bool WILD>
__device__ __forceinline__ void dp_pair(const AlignParams &P, const int64_t p, const int lane, int2 *__restrict__ bnd,
                                        uint4 *__restrict__ prof_warp)
{
    uint4 *prof = prof_warp + lane;
    constexpr int RSH = WILD ? CODE_SHIFT : 0;    // reference bases: shifted codes (arithmetic form) or plain codes (profile)
    constexpr int C2 = 2;                         // columns per step
    const unsigned FULL = 0xffffffffu;
    const int n = P.pn[p], m = P.pm[p];
    const uint8_t *__restrict__ q = P.codes + P.seq_off[P.pa[p]];
    const uint8_t *__restrict__ rf = P.codes + P.seq_off[P.pb[p]];
    uint32_t *__restrict__ trace = P.trace + (P.ptrace[p] - P.chunk_base);
    const int T = (m + C2 - 1) / C2 + 31;
    const int nstripes = (n + SROWS - 1) / SROWS;
    const int last_lane = ((n - 1) % SROWS) / R, last_r = (n - 1) % R;
    int best_col = INT_MIN, best_col_i = 0, best_row = INT_MIN, best_row_j = 0;
    if (lane == 0) { P.playout[p] = C2; P.plastrr[p] = R; }
    const int t_steady_end = (m - 2 * C2) / C2 + 1;

    for (int s = 0; s < nstripes; ++s) {
        const int i0 = s * SROWS + lane * R;
        P32State S;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            S.qc[r] = (i0 + r < n) ? ((int)q[i0 + r] << CODE_SHIFT) : 0;
            S.Hl[r] = 0;           // H(i,-1) = 0: free leading gap
            S.El[r] = NEGV;
        }
        if (!WILD) {
            const int sc_mat = P.c_sm + (3 << CODE_SHIFT), sc_mis = P.c_sx;       // 16*match+12, 16*mismatch+12
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t v[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) v[r] = (uint32_t)(((S.qc[r] >> CODE_SHIFT) == c) ? sc_mat : sc_mis);
                prof[c * 64] = make_uint4(v[0], v[1], v[2], v[3]);
                prof[c * 64 + 32] = make_uint4(v[4], v[5], v[6], v[7]);
            }
            __syncwarp();
        }
        const bool has_prev = s > 0, has_next = s + 1 < nstripes, is_last = !has_next;
        const bool track = is_last && lane == last_lane;
        S.diag0 = 0;
#pragma unroll
        for (int c = 0; c < C2; ++c) {
            S.Hbot[c] = 0; S.Fbot[c] = NEGV; S.rc_next[c] = 0;
            S.b_next[c] = make_int2(0, NEGV);
            if (lane == 0) {
                S.rc_next[c] = (c < m) ? ((int)rf[c] << RSH) : 0;
                if (has_prev && c < m) S.b_next[c] = __ldcg(&bnd[c]);
            }
        }
        uint2 *tp = reinterpret_cast<uint2 *>(trace) + (size_t)s * T * 32 + lane;

        auto general_step = [&](const int t) {
            const int jb = C2 * (t - lane);
            int rc[C2], hu[C2], fu[C2];
#pragma unroll
            for (int c = 0; c < C2; ++c) {
                rc[c] = S.rc_next[c];
                hu[c] = __shfl_up_sync(FULL, S.Hbot[c], 1);
                fu[c] = __shfl_up_sync(FULL, S.Fbot[c], 1);
                if (lane == 0) { hu[c] = S.b_next[c].x; fu[c] = S.b_next[c].y; }
            }
            {
                const int jn = jb + C2;
                if ((unsigned)jn < (unsigned)m) {
#pragma unroll
                    for (int c = 0; c < C2; ++c) {
                        S.rc_next[c] = (jn + c < m) ? ((int)rf[jn + c] << RSH) : 0;
                        if (has_prev && lane == 0 && jn + c < m) S.b_next[c] = __ldcg(&bnd[jn + c]);
                    }
                }
            }
            if ((unsigned)jb < (unsigned)m) {
                uint32_t w[C2];
#pragma unroll
                for (int c = 0; c < C2; ++c) {
                    p32_column<WILD>(S, P, rc[c], hu[c], fu[c], (c == 0) ? S.diag0 : hu[c - 1], S.Hbot[c], S.Fbot[c], w[c], prof);
                    const int j = jb + c;
                    ISOF_ASSERT(j < P.bnd_stride, CHK_BND);
                    if (has_next && lane == 31 && j < m) bnd[j] = make_int2(S.Hbot[c], S.Fbot[c]);
                    if (j == m - 1) {                      // last column: smallest i wins ties
#pragma unroll
                        for (int r = 0; r < R; ++r)
                            if (i0 + r < n && S.Hl[r] > best_col) { best_col = S.Hl[r]; best_col_i = i0 + r; }
                    }
                    if (track && j < m) {                  // last row: smallest j wins ties
                        int hv = S.Hl[0];
#pragma unroll
                        for (int r = 1; r < R; ++r) if (r == last_r) hv = S.Hl[r];
                        if (hv > best_row) { best_row = hv; best_row_j = j; }
                    }
                }
                S.diag0 = hu[C2 - 1];
                ISOF_ASSERT((uint32_t *)(tp + (size_t)t * 32 + 1) <= P.trace + (P.ptrace[p + 1] - P.chunk_base) - TRACE_SLACK, CHK_TRACE_STORE);
                tp[(size_t)t * 32] = make_uint2(w[0], w[1]);
            }
        };

        const bool can_steady = t_steady_end > 31;
        const int t1 = can_steady ? 31 : T, t2 = can_steady ? t_steady_end : T;
        for (int t = 0; t < t1; ++t) general_step(t);
        if (can_steady) {
            const uint8_t *pr = rf + C2 * (31 - lane) + C2;
            const int2 *pbn = bnd + C2 * 31 + C2;
            int2 *pbs = bnd;
            uint2 *sp = tp + (size_t)31 * 32;
            auto steady_step = [&](const int t, const bool with_track) {
                int rc[C2], hu[C2], fu[C2];
#pragma unroll
                for (int c = 0; c < C2; ++c) {
                    rc[c] = S.rc_next[c];
                    hu[c] = __shfl_up_sync(FULL, S.Hbot[c], 1);
                    fu[c] = __shfl_up_sync(FULL, S.Fbot[c], 1);
                    if (lane == 0) { hu[c] = S.b_next[c].x; fu[c] = S.b_next[c].y; }
                }
#pragma unroll
                for (int c = 0; c < C2; ++c) { ISOF_ASSERT(pr + c < rf + m, CHK_SEQ_LOAD); S.rc_next[c] = (int)pr[c] << RSH; }
                if (has_prev && lane == 0) {
#pragma unroll
                    for (int c = 0; c < C2; ++c) S.b_next[c] = __ldcg(pbn + c);
                }
                uint32_t w[C2];
#pragma unroll
                for (int c = 0; c < C2; ++c) {
                    p32_column<WILD>(S, P, rc[c], hu[c], fu[c], (c == 0) ? S.diag0 : hu[c - 1], S.Hbot[c], S.Fbot[c], w[c], prof);
                    if (with_track && track) {             // last row: smallest j wins ties
                        int hv = S.Hl[0];
#pragma unroll
                        for (int r = 1; r < R; ++r) if (r == last_r) hv = S.Hl[r];
                        if (hv > best_row) { best_row = hv; best_row_j = C2 * (t - lane) + c; }
                    }
                }
                S.diag0 = hu[C2 - 1];
                ISOF_ASSERT((uint32_t *)(sp + 1) <= P.trace + (P.ptrace[p + 1] - P.chunk_base) - TRACE_SLACK, CHK_TRACE_STORE);
                *sp = make_uint2(w[0], w[1]);
                if (has_next && lane == 31) {
#pragma unroll
                    for (int c = 0; c < C2; ++c) { ISOF_ASSERT(pbs + c < bnd + P.bnd_stride, CHK_BND); pbs[c] = make_int2(S.Hbot[c], S.Fbot[c]); }
                }
                pr += C2; pbn += C2; pbs += C2; sp += 32;
            };
            if (!is_last) {
#pragma unroll 2
                for (int t = 31; t < t2; ++t) steady_step(t, false);
            } else {
#pragma unroll 1
                for (int t = 31; t < t2; ++t) steady_step(t, true);
            }
            for (int t = t2; t < T; ++t) general_step(t);
        }
        __syncwarp();
    }
    // ---- end cell (restated library order: last row first, last column only if strictly greater)
    best_row = __shfl_sync(FULL, best_row, last_lane);
    best_row_j = __shfl_sync(FULL, best_row_j, last_lane);
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        int ov = __shfl_xor_sync(FULL, best_col, d);
        int oi = __shfl_xor_sync(FULL, best_col_i, d);
        if (ov > best_col || (ov == best_col && oi < best_col_i)) { best_col = ov; best_col_i = oi; }
    }
    if (lane == 0) {
        int sc = best_row, eq = n - 1, er = best_row_j;
        if (best_col > sc) { sc = best_col; eq = best_col_i; er = m - 1; }
        else if (best_col == sc && er == m - 1 && best_col_i < eq) eq = best_col_i;
        P.score[p] = sc >> TAG_BITS;
        P.endq[p] = eq;
        P.endr[p] = er;
    }
}

// ---------------------------------------------------------------------------------------------
// packed path: two alignments (A in the low, B in the high 16 bits of every register) per warp.
// Values are 4*score + tag (2 bits).  E and F are kept re-tagged with their class for the 3-way
// max (E: 01, F: 10, diagonal: 11, so VIMNMX3.S16x2 resolves DIAG > F > E), and that class tag
// doubles as the "came from extension" marker of the next cell: E_ext = Eh + 4*(-ext) keeps tag 01,
// F_ext = Fh + 4*(-ext) - 1 turns 10 into 01, while the opening candidates carry tag 00, so bit 0 of
// max(open, ext) is the extension flag and extension wins ties.  Every operation is per-half
// (VIADD.16x2 / VIADDMNMX.S16x2 / VIMNMX3.S16x2 / LOP3), so the two alignments never interact;
// they may differ in size (loops run to the larger one, stores and end tracking are per half).
// Trace words are unpacked per alignment into the same layout the 32-bit path writes.
// ---------------------------------------------------------------------------------------------
constexpr int CS16 = 8;                       // base codes at bits 8..9 of each half
constexpr uint32_t NEG16 = 0x83008300u;       // -32000 in both halves (a clean multiple of 4)
constexpr int PC = 2;                         // columns per step of the packed path (uint2 trace stores)
constexpr int TRACE_FMT_SPLIT = 16;           // playout flag: trace word = [src x8 (hi16) | fe x8 (lo16)], 2 bits per row

__device__ __forceinline__ int lo16(uint32_t x) { return (int)(short)(x & 0xffffu); }
__device__ __forceinline__ int hi16(uint32_t x) { return ((int)x) >> 16; }

// Lane l handles, at step t, the PC consecutive columns starting at jb = PC*(t-l): the per-step
// overhead (shuffles, prefetch, stores, loop) is paid once per PC*8 cells per alignment.
// Trace layout of a packed pair: word[((stripe*T + t)*32 + lane)*PC + c], T = ceil(m/PC)+31, column
// j = PC*(t-lane)+c; P.playout[p] = PC tells the traceback.
//
// The step loop of a stripe is peeled into fill / steady / drain.  In the steady range every lane
// is inside both alignments with a full look-ahead block, so that loop has no bounds checks, no
// end-cell bookkeeping and unconditional stores; fill, drain and the last stripes (which track the
// end cell) go through the general step.
struct P16Const {
    uint32_t h_eo, h_ee, h_fo, h_fe, h_sm, h_sx;
};

struct P16State {
    uint32_t qc[R], Hl[R], Eh[R];
    uint32_t Hbot[PC], Fbot[PC], rc_next[PC];
    int2 b_next[PC];
    uint32_t diag0;
};

// Hl[idx] for a run-time idx < RR without dynamic register indexing: a binary tree of selects on the bits of idx
// (7 SEL for RR = 8; the compare-and-select chain it replaces was 7 ISETP + 7 SEL)
template <int RR>
__device__ __forceinline__ uint32_t pick_row(const uint32_t (&h)[R], const int idx)
{
    const bool b0 = idx & 1, b1 = idx & 2, b2 = idx & 4;
    const uint32_t x0 = b0 ? h[1] : h[0];
    if (RR == 2) return x0;
    const uint32_t x1 = b0 ? h[3] : h[2];
    const uint32_t y0 = b1 ? x1 : x0;
    if (RR == 4) return y0;
    const uint32_t x2 = b0 ? h[5] : h[4];
    if (RR == 6) return b2 ? x2 : y0;
    const uint32_t x3 = b0 ? h[7] : h[6];
    const uint32_t y1 = b1 ? x3 : x2;
    return b2 ? y1 : y0;
}

// one column of RR packed rows per lane; returns the two trace words (A, B).
// PROF: both alignments share the reference, so the substitution scores of this lane's rows against the
// column's base come from the lane's query profile in shared memory (two LDS.128 on the otherwise idle
// LSU pipe) instead of one LOP3 + one VIADDMNMX per row; `rc` is then the plain base code 0..3.
template <int RR, bool PROF>
__device__ __forceinline__ void p16_column(P16State &S, const P16Const &K, const uint32_t rc, uint32_t up, uint32_t fup,
                                           uint32_t diag, uint32_t &hbot, uint32_t &fbot, uint32_t &wA, uint32_t &wB,
                                           const uint4 *__restrict__ prof)
{
    const uint32_t C3 = (3u << CS16) | (3u << (CS16 + 16));
    uint32_t scv[8];
    if (PROF) {
        const uint4 v0 = prof[rc * 64];                  // [code][half 0][lane]
        scv[0] = v0.x; scv[1] = v0.y; scv[2] = v0.z; scv[3] = v0.w;
        if (RR > 4) {
            const uint4 v1 = prof[rc * 64 + 32];         // [code][half 1][lane]
            scv[4] = v1.x; scv[5] = v1.y; scv[6] = v1.z; scv[7] = v1.w;
        }
    }
    // Trace bits are gathered with integer multiply-adds (FMA pipe), not shifts and masks (ALU pipe, the
    // bound): per half, acc_ef collects [f e] of row r at bits 2r+1:2r and acc_h the source tag of H.
    // H - (H & ~3) is the tag, so acc_h = sum_r (H_r - Hc_r) * 4^r needs no mask of its own.
    uint32_t acc_ef = 0u, acc_h = 0u;
#pragma unroll
    for (int r = 0; r < RR; ++r) {
        const uint32_t left = S.Hl[r];
        const uint32_t Ee = __vadd2(S.Eh[r], K.h_ee);               // tag 01 (extension)
        const uint32_t E = __viaddmax_s16x2(left, K.h_eo, Ee);      // open: tag 00; bit 0 = extended, ties extend
        const uint32_t Ehn = E | 0x00010001u;
        const uint32_t Fe = __vadd2(fup, K.h_fe);                   // tag 10 (extension)
        const uint32_t F = __viaddmax_s16x2(up, K.h_fo, Fe);        // open: tag 00; bit 1 = extended, ties extend
        const uint32_t Fh = F | 0x00020002u;
        uint32_t sc;
        if (PROF) sc = scv[r];
        else {
            const uint32_t y = S.qc[r] ^ rc ^ C3;
            sc = __viaddmax_s16x2(y, K.h_sm, K.h_sx);
        }
        // max(diag + sc, E) first: it depends on the previous column only, so the row-to-row chain through F is
        // F -> Fh -> H -> Hc (4 dependent ops) instead of 5
        const uint32_t D = __viaddmax_s16x2(diag, sc, Ehn);         // diagonal: tag 11
        const uint32_t H = __vmaxs2(D, Fh);                         // ties: diagonal > F > E
        const uint32_t Hc = H & 0xfffcfffcu;
        const uint32_t ef = (E | F) & 0x00030003u;
        acc_ef = mad_u32(ef, 1u << (2 * r), acc_ef);
        acc_h = mad_u32(H, 1u << (2 * r), acc_h);
        acc_h = mad_u32(Hc, 0u - (1u << (2 * r)), acc_h);
        diag = left;
        S.Hl[r] = Hc;
        S.Eh[r] = Ehn;
        up = Hc;
        fup = Fh;
    }
    hbot = up; fbot = fup;
    wA = __byte_perm(acc_ef, acc_h, 0x5410);         // A: [src x8 | fe x8]
    wB = __byte_perm(acc_ef, acc_h, 0x7632);
}

__device__ __forceinline__ void dp_pair16(const AlignParams &P, const int64_t pA, const int64_t pB, const int lane,
                                          int2 *__restrict__ bnd, uint4 *__restrict__ prof_warp)
{
    const unsigned FULL = 0xffffffffu;
    const int nA = P.pn[pA], mA = P.pm[pA], nB = P.pn[pB], mB = P.pm[pB];
    const uint8_t *__restrict__ qA = P.codes + P.seq_off[P.pa[pA]];
    const uint8_t *__restrict__ rA = P.codes + P.seq_off[P.pb[pA]];
    const uint8_t *__restrict__ qB = P.codes + P.seq_off[P.pa[pB]];
    const uint8_t *__restrict__ rB = P.codes + P.seq_off[P.pb[pB]];
    uint32_t *__restrict__ traceA = P.trace + (P.ptrace[pA] - P.chunk_base);
    uint32_t *__restrict__ traceB = P.trace + (P.ptrace[pB] - P.chunk_base);
    const int m = max(mA, mB), mmin = min(mA, mB);
    const int T = (m + PC - 1) / PC + 31, TA = (mA + PC - 1) / PC + 31, TB = (mB + PC - 1) / PC + 31;
    const int nsA = (nA + SROWS - 1) / SROWS, nsB = (nB + SROWS - 1) / SROWS, ns = max(nsA, nsB);
    int last_laneA = 0, last_rA = 0, last_laneB = 0, last_rB = 0;       // set in the alignment's last stripe
    P16Const K;
    // the F constants equal the E constants (same penalties, tags come from the stored values): one register each
    K.h_eo = P.h_eo; K.h_ee = P.h_ee; K.h_fo = K.h_eo; K.h_fe = K.h_ee; K.h_sm = P.h_sm; K.h_sx = P.h_sx;
    int bcolA = INT_MIN, bcolA_i = 0, browA = INT_MIN, browA_j = 0;
    int bcolB = INT_MIN, bcolB_i = 0, browB = INT_MIN, browB_j = 0;
    if (lane == 0) { P.playout[pA] = PC | TRACE_FMT_SPLIT; P.playout[pB] = PC | TRACE_FMT_SPLIT; }
    // steady range of t: lane 31 has started (t >= 31) and lane 0's look-ahead block is inside both
    // references: PC*t + 2*PC <= mmin
    const int t_steady_end = (mmin - 2 * PC) / PC + 1;          // exclusive; may be <= 31 (then no steady range)

    // The last stripe of the task is compacted: with `need` rows left, every lane takes RR = 2, 4, 6 or 8
    // rows (the smallest that covers them), so no lane idles on padding rows.  Earlier stripes are full.
    // same reference in both halves (the all-pairs drivers order pairs column-major to make this the rule)
    const bool same_ref = (rA == rB);
    uint4 *prof = prof_warp + lane;                   // [code 4][half 2][lane 32] uint4 = 4 packed scores
    const uint32_t sc_mis = K.h_sx & 0xffffu;                                  // 4*mismatch + 3
    const uint32_t sc_mat = (K.h_sm + (3u << CS16)) & 0xffffu;                 // 4*match + 3
    auto run_stripe = [&](auto rr_tag, auto prof_tag, const int s) {
        constexpr int RR = decltype(rr_tag)::value;
        constexpr bool PROF = decltype(prof_tag)::value;
        const int i0 = s * SROWS + lane * RR;
        if (s == nsA - 1) { last_laneA = ((nA - 1) - s * SROWS) / RR; last_rA = ((nA - 1) - s * SROWS) % RR; if (lane == 0) P.plastrr[pA] = RR; }
        if (s == nsB - 1) { last_laneB = ((nB - 1) - s * SROWS) / RR; last_rB = ((nB - 1) - s * SROWS) % RR; if (lane == 0) P.plastrr[pB] = RR; }
        P16State S;
#pragma unroll
        for (int r = 0; r < RR; ++r) {
            const uint32_t a = (i0 + r < nA) ? (uint32_t)qA[i0 + r] : 0u;
            const uint32_t b = (i0 + r < nB) ? (uint32_t)qB[i0 + r] : 0u;
            S.qc[r] = (a << CS16) | (b << (CS16 + 16));
            S.Hl[r] = 0u;
            S.Eh[r] = NEG16 | 0x00010001u;
        }
        if (PROF) {
            // query profile of this lane's rows: packed (A | B << 16) substitution scores per base code
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t v[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const uint32_t qa = (r < RR) ? ((S.qc[r < RR ? r : 0] >> CS16) & 7u) : 4u;
                    const uint32_t qb = (r < RR) ? ((S.qc[r < RR ? r : 0] >> (CS16 + 16)) & 7u) : 4u;
                    v[r] = ((qa == (uint32_t)c) ? sc_mat : sc_mis) | (((qb == (uint32_t)c) ? sc_mat : sc_mis) << 16);
                }
                prof[c * 64] = make_uint4(v[0], v[1], v[2], v[3]);
                prof[c * 64 + 32] = make_uint4(v[4], v[5], v[6], v[7]);
            }
            __syncwarp();
        }
        const bool has_prev = s > 0, has_next = s + 1 < ns;
        const bool liveA = s < nsA, liveB = s < nsB;
        const bool trackA = (s == nsA - 1) && lane == last_laneA, trackB = (s == nsB - 1) && lane == last_laneB;
        const bool track_stripe = (s == nsA - 1) || (s == nsB - 1);          // warp-uniform
        S.diag0 = 0u;
#pragma unroll
        for (int c = 0; c < PC; ++c) {
            S.Hbot[c] = 0u;
            S.Fbot[c] = NEG16 | 0x00020002u;
            S.rc_next[c] = 0u;
            S.b_next[c] = make_int2(0, (int)(NEG16 | 0x00020002u));
            if (lane == 0) {
                const uint32_t a = (c < mA) ? (uint32_t)rA[c] : 0u, b = (c < mB) ? (uint32_t)rB[c] : 0u;
                S.rc_next[c] = PROF ? a : ((a << CS16) | (b << (CS16 + 16)));
                if (has_prev && c < m) S.b_next[c] = __ldcg(&bnd[c]);
            }
        }
        uint2 *tpA = reinterpret_cast<uint2 *>(traceA) + (size_t)s * TA * 32 + lane;
        uint2 *tpB = reinterpret_cast<uint2 *>(traceB) + (size_t)s * TB * 32 + lane;
        uint32_t snap[RR];                           // H of the last column (A low half, B high half), set in the drain
#pragma unroll
        for (int r = 0; r < RR; ++r) snap[r] = 0x80008000u;

        // ---------------- general step (bounds checks + end-cell bookkeeping) ----------------
        auto general_step = [&](const int t, auto track_tag, auto row_tag) {
            constexpr bool TR = decltype(track_tag)::value;      // false: the step cannot reach a last column (fill steps before a steady range)
            constexpr bool TROW = decltype(row_tag)::value;      // false: not a last stripe, no last-row bookkeeping
            const int jb = PC * (t - lane);
            uint32_t rc[PC], hu[PC], fu[PC];
#pragma unroll
            for (int c = 0; c < PC; ++c) {
                rc[c] = S.rc_next[c];
                hu[c] = __shfl_up_sync(FULL, S.Hbot[c], 1);
                fu[c] = __shfl_up_sync(FULL, S.Fbot[c], 1);
                if (lane == 0) { hu[c] = (uint32_t)S.b_next[c].x; fu[c] = (uint32_t)S.b_next[c].y; }
            }
            {
                const int jn = jb + PC;
                if ((unsigned)jn < (unsigned)m) {
#pragma unroll
                    for (int c = 0; c < PC; ++c) {
                        const uint32_t a = (jn + c < mA) ? (uint32_t)rA[jn + c] : 0u;
                        const uint32_t b = (jn + c < mB) ? (uint32_t)rB[jn + c] : 0u;
                        S.rc_next[c] = PROF ? a : ((a << CS16) | (b << (CS16 + 16)));
                        if (has_prev && lane == 0 && jn + c < m) S.b_next[c] = __ldcg(&bnd[jn + c]);
                    }
                }
            }
            if ((unsigned)jb < (unsigned)m) {
                uint32_t wA[PC], wB[PC];
#pragma unroll
                for (int c = 0; c < PC; ++c) {
                    p16_column<RR, PROF>(S, K, rc[c], hu[c], fu[c], (c == 0) ? S.diag0 : hu[c - 1], S.Hbot[c], S.Fbot[c], wA[c], wB[c], prof);
                    const int j = jb + c;
                    ISOF_ASSERT(j < P.bnd_stride, CHK_BND);
                    if (has_next && lane == 31 && j < m) bnd[j] = make_int2((int)S.Hbot[c], (int)S.Fbot[c]);
                    // last column of A / B: every lane passes it exactly once per stripe; keep the column (one LOP3 per
                    // row) and fold it into the running maximum after the stripe's loop
                    if (TR && j == mA - 1) {
#pragma unroll
                        for (int r = 0; r < RR; ++r) snap[r] = (snap[r] & 0xffff0000u) | (S.Hl[r] & 0x0000ffffu);
                    }
                    if (TR && j == mB - 1) {
#pragma unroll
                        for (int r = 0; r < RR; ++r) snap[r] = (snap[r] & 0x0000ffffu) | (S.Hl[r] & 0xffff0000u);
                    }
                    if (TROW && trackA && j < mA) {         // last row of A: smallest j wins ties
                        const uint32_t hv = pick_row<RR>(S.Hl, last_rA);
                        const int v = lo16(hv);
                        if (v > browA) { browA = v; browA_j = j; }
                    }
                    if (TROW && trackB && j < mB) {
                        const uint32_t hv = pick_row<RR>(S.Hl, last_rB);
                        const int v = hi16(hv);
                        if (v > browB) { browB = v; browB_j = j; }
                    }
                }
                S.diag0 = hu[PC - 1];
                if (liveA && jb < mA) {
                    ISOF_ASSERT((uint32_t *)(tpA + (size_t)t * 32 + 1) <= P.trace + (P.ptrace[pA + 1] - P.chunk_base) - TRACE_SLACK, CHK_TRACE_STORE);
                    tpA[(size_t)t * 32] = make_uint2(wA[0], wA[1]);
                }
                if (liveB && jb < mB) {
                    ISOF_ASSERT((uint32_t *)(tpB + (size_t)t * 32 + 1) <= P.trace + (P.ptrace[pB + 1] - P.chunk_base) - TRACE_SLACK, CHK_TRACE_STORE);
                    tpB[(size_t)t * 32] = make_uint2(wB[0], wB[1]);
                }
            }
        };

        const bool can_steady = liveA && liveB && t_steady_end > 31;
        const int t1 = can_steady ? 31 : T, t2 = can_steady ? t_steady_end : T;
        // a steady range exists only when the shorter reference has >= 66 columns; the 31 fill steps before it
        // stay below column 62 and cannot reach a last column
        if (can_steady && !track_stripe) { for (int t = 0; t < t1; ++t) general_step(t, std::false_type{}, std::false_type{}); }
        else if (can_steady) { for (int t = 0; t < t1; ++t) general_step(t, std::false_type{}, std::true_type{}); }
        else { for (int t = 0; t < t1; ++t) general_step(t, std::true_type{}, std::true_type{}); }
        if (can_steady) {
            // ---------------- steady state: no bounds checks; end-cell bookkeeping only in the last stripes,
            // and there only the running maximum of the last row (the last column is reached in the drain)
            const uint8_t *pa_ = rA + PC * (31 - lane) + PC;       // look-ahead block of step t = 31
            const uint8_t *pb_ = rB + PC * (31 - lane) + PC;
            const int2 *pbn = bnd + PC * 31 + PC;                  // lane 0 only
            int2 *pbs = bnd + PC * (31 - 31);                      // lane 31 only: jb at t = 31
            uint2 *sA = tpA + (size_t)31 * 32, *sB = tpB + (size_t)31 * 32;
            auto steady_step = [&](const int t, const bool with_track) {
                uint32_t rc[PC], hu[PC], fu[PC];
#pragma unroll
                for (int c = 0; c < PC; ++c) {
                    rc[c] = S.rc_next[c];
                    hu[c] = __shfl_up_sync(FULL, S.Hbot[c], 1);
                    fu[c] = __shfl_up_sync(FULL, S.Fbot[c], 1);
                    if (lane == 0) { hu[c] = (uint32_t)S.b_next[c].x; fu[c] = (uint32_t)S.b_next[c].y; }
                }
#pragma unroll
                for (int c = 0; c < PC; ++c)
                {
                    ISOF_ASSERT(pa_ + c < rA + mA && pb_ + c < rB + mB && pa_ >= rA && pb_ >= rB, CHK_SEQ_LOAD);
                    S.rc_next[c] = PROF ? (uint32_t)pa_[c] : (((uint32_t)pa_[c] << CS16) | ((uint32_t)pb_[c] << (CS16 + 16)));
                }
                if (has_prev && lane == 0) {
#pragma unroll
                    for (int c = 0; c < PC; ++c) S.b_next[c] = __ldcg(pbn + c);
                }
                uint32_t wA[PC], wB[PC];
#pragma unroll
                for (int c = 0; c < PC; ++c) {
                    p16_column<RR, PROF>(S, K, rc[c], hu[c], fu[c], (c == 0) ? S.diag0 : hu[c - 1], S.Hbot[c], S.Fbot[c], wA[c], wB[c], prof);
                    if (with_track) {
                        const int j = PC * (t - lane) + c;
                        if (trackA) {
                            const uint32_t hv = pick_row<RR>(S.Hl, last_rA);
                            const int v = lo16(hv);
                            if (v > browA) { browA = v; browA_j = j; }
                        }
                        if (trackB) {
                            const uint32_t hv = pick_row<RR>(S.Hl, last_rB);
                            const int v = hi16(hv);
                            if (v > browB) { browB = v; browB_j = j; }
                        }
                    }
                }
                S.diag0 = hu[PC - 1];
                ISOF_ASSERT((uint32_t *)(sA + 1) <= P.trace + (P.ptrace[pA + 1] - P.chunk_base) - TRACE_SLACK, CHK_TRACE_STORE);
                ISOF_ASSERT((uint32_t *)(sB + 1) <= P.trace + (P.ptrace[pB + 1] - P.chunk_base) - TRACE_SLACK, CHK_TRACE_STORE);
                ISOF_ASSERT((uint32_t *)sA >= traceA && (uint32_t *)sB >= traceB, CHK_TRACE_STORE);
                *sA = make_uint2(wA[0], wA[1]);
                *sB = make_uint2(wB[0], wB[1]);
                if (has_next && lane == 31) {
#pragma unroll
                    for (int c = 0; c < PC; ++c) { ISOF_ASSERT(pbs + c < bnd + P.bnd_stride, CHK_BND); pbs[c] = make_int2((int)S.Hbot[c], (int)S.Fbot[c]); }
                }
                pa_ += PC; pb_ += PC; pbn += PC; pbs += PC; sA += 32; sB += 32;
            };
            if (!track_stripe) {
#pragma unroll 2
                for (int t = 31; t < t2; ++t) steady_step(t, false);
            } else {
#pragma unroll 2
                for (int t = 31; t < t2; ++t) steady_step(t, true);
            }
            if (track_stripe) { for (int t = t2; t < T; ++t) general_step(t, std::true_type{}, std::true_type{}); }
            else { for (int t = t2; t < T; ++t) general_step(t, std::true_type{}, std::false_type{}); }
        }
        // end cell, last column: smallest i wins ties (rows ascend within a lane and from stripe to stripe)
        if (liveA) {
#pragma unroll
            for (int r = 0; r < RR; ++r) {
                const int v = lo16(snap[r]);
                if (i0 + r < nA && v > bcolA) { bcolA = v; bcolA_i = i0 + r; }
            }
        }
        if (liveB) {
#pragma unroll
            for (int r = 0; r < RR; ++r) {
                const int v = hi16(snap[r]);
                if (i0 + r < nB && v > bcolB) { bcolB = v; bcolB_i = i0 + r; }
            }
        }
        __syncwarp();
    };
    for (int s = 0; s < ns; ++s) {
        const int need = max(min(nA - s * SROWS, SROWS), min(nB - s * SROWS, SROWS));
        if (same_ref) {
            if (need > 192) run_stripe(std::integral_constant<int, 8>{}, std::true_type{}, s);
            else if (need > 128) run_stripe(std::integral_constant<int, 6>{}, std::true_type{}, s);
            else if (need > 64) run_stripe(std::integral_constant<int, 4>{}, std::true_type{}, s);
            else run_stripe(std::integral_constant<int, 2>{}, std::true_type{}, s);
        } else {
            if (need > 192) run_stripe(std::integral_constant<int, 8>{}, std::false_type{}, s);
            else if (need > 128) run_stripe(std::integral_constant<int, 6>{}, std::false_type{}, s);
            else if (need > 64) run_stripe(std::integral_constant<int, 4>{}, std::false_type{}, s);
            else run_stripe(std::integral_constant<int, 2>{}, std::false_type{}, s);
        }
    }
    browA = __shfl_sync(FULL, browA, last_laneA); browA_j = __shfl_sync(FULL, browA_j, last_laneA);
    browB = __shfl_sync(FULL, browB, last_laneB); browB_j = __shfl_sync(FULL, browB_j, last_laneB);
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        int ov = __shfl_xor_sync(FULL, bcolA, d), oi = __shfl_xor_sync(FULL, bcolA_i, d);
        if (ov > bcolA || (ov == bcolA && oi < bcolA_i)) { bcolA = ov; bcolA_i = oi; }
        ov = __shfl_xor_sync(FULL, bcolB, d); oi = __shfl_xor_sync(FULL, bcolB_i, d);
        if (ov > bcolB || (ov == bcolB && oi < bcolB_i)) { bcolB = ov; bcolB_i = oi; }
    }
    if (lane == 0) {
        int sc = browA, eq = nA - 1, er = browA_j;
        if (bcolA > sc) { sc = bcolA; eq = bcolA_i; er = mA - 1; }
        else if (bcolA == sc && er == mA - 1 && bcolA_i < eq) eq = bcolA_i;
        P.score[pA] = sc >> 2; P.endq[pA] = eq; P.endr[pA] = er;
        sc = browB; eq = nB - 1; er = browB_j;
        if (bcolB > sc) { sc = bcolB; eq = bcolB_i; er = mB - 1; }
        else if (bcolB == sc && er == mB - 1 && bcolB_i < eq) eq = bcolB_i;
        P.score[pB] = sc >> 2; P.endq[pB] = eq; P.endr[pB] = er;
    }
}

__device__ __forceinline__ void dp_pair32(const AlignParams &P, const int64_t p, const int lane, int2 *__restrict__ bnd,
                                          uint4 *__restrict__ prof_warp)
{
    if (P.pwild[p] & 1) dp_pair<true>(P, p, lane, bnd, prof_warp);
    else dp_pair<false>(P, p, lane, bnd, prof_warp);
}

// One task = two consecutive pairs of the chunk.  Both eligible -> one packed pass; otherwise each
// pair runs through the 32-bit path.
__global__ void __launch_bounds__(DP_WARPS * 32, DP_CTAS_PER_SM)
sg_dp_kernel(const AlignParams P, const int64_t start, const int64_t end, unsigned long long *__restrict__ counter)
{
    __shared__ uint4 s_prof[DP_WARPS][4 * 2 * 32];      // per warp: query profile [code][row half][lane], 4 KB
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int2 *bnd = P.bnd + (size_t)(blockIdx.x * DP_WARPS + warp) * P.bnd_stride;
    for (;;) {
        unsigned long long tkt = 0;
        if (lane == 0) tkt = atomicAdd(counter, 1ull);
        tkt = __shfl_sync(0xffffffffu, tkt, 0);
        // tickets [0, n_pk): packed tasks over the eligible pairs; the rest: one 32-bit pair per ticket
        const int64_t n_pk = (int64_t)(*P.n_elig) / 2;
        if ((int64_t)tkt < n_pk) {
            const int64_t xA = start + 2 * (int64_t)tkt;
            dp_pair16(P, P.order[xA], P.order[xA + 1], lane, bnd, s_prof[warp]);
        } else {
            const int64_t x = start + 2 * n_pk + ((int64_t)tkt - n_pk);
            if (x >= end) break;
            dp_pair32(P, P.order[x], lane, bnd, s_prof[warp]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// traceback -> reversed runs at the tail of the trace region
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int upc(int c) { return (c >= 'a' && c <= 'z') ? c - 32 : c; }

__global__ void sg_traceback_kernel(const AlignParams P, const int64_t start, const int64_t end)
{
    const int64_t p = start + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= end) return;
    const int n = P.pn[p], m = P.pm[p];
    const uint8_t *s1 = P.cat + P.seq_off[P.pa[p]];
    const uint8_t *s2 = P.cat + P.seq_off[P.pb[p]];
    const uint32_t *trace = P.trace + (P.ptrace[p] - P.chunk_base);
    uint32_t *tail = P.trace + (P.ptrace[p + 1] - P.chunk_base);     // runs go to tail[-1], tail[-2], ...
    const int lay = P.playout[p];
    const int pc = lay & 15;                          // columns per step of the writer
    const bool split = (lay & 16) != 0;               // packed path: [src x8 | fe x8] 2-bit fields; 32-bit path: nibbles
    const int T = (m + pc - 1) / pc + 31;
    const int last_s = (n - 1) >> 8, rr_last = P.plastrr[p];   // rows per lane in the (possibly compacted) last stripe
    int i = P.endq[p], j = P.endr[p];
    int nr = 0;
    uint32_t cur_op = 0xffu, cur_len = 0;
    const uint32_t *last_read = trace - 1;            // highest trace word still to be read (debug assertion)
    (void)last_read;
#define PUSH(op, len)                                                                              \
    do {                                                                                           \
        if ((len) > 0) {                                                                           \
            if (cur_op == (uint32_t)(op)) cur_len += (uint32_t)(len);                              \
            else {                                                                                 \
                if (cur_len) { ++nr; ISOF_ASSERT(tail - nr > last_read, CHK_TAIL_OVERLAP); tail[-nr] = (cur_len << 4) | cur_op; } \
                cur_op = (op); cur_len = (len);                                                    \
            }                                                                                      \
        }                                                                                          \
    } while (0)
    if (i + 1 == n) PUSH(ISOF_CIGAR_D, m - 1 - j);
    else PUSH(ISOF_CIGAR_I, n - 1 - i);
    int where = 3;
    while (i >= 0 || j >= 0) {
        if (i < 0) { PUSH(ISOF_CIGAR_D, j + 1); break; }
        if (j < 0) { PUSH(ISOF_CIGAR_I, i + 1); break; }
        const int s = i >> 8;
        int l = (i >> 3) & 31, r = i & 7;
        if (s == last_s && rr_last != 8) { l = (i & 255) / rr_last; r = (i & 255) % rr_last; }
        const int jt = (pc == 1) ? j : (j >> 1), jc = (pc == 1) ? 0 : (j & 1);
        {   // the path mostly follows the diagonal: pull the line it will need 16 steps from now into L2
            const int ip = i - 16, jp = j - 16;
            if (ip >= 0 && jp >= 0) {
                const int sp_ = ip >> 8;
                const int lp = (sp_ == last_s && rr_last != 8) ? (ip & 255) / rr_last : (ip >> 3) & 31;
                const int jtp = (pc == 1) ? jp : (jp >> 1);
                const uint32_t *pf = &trace[(((size_t)sp_ * T + (jtp + lp)) * 32 + lp) * pc];
                asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
            }
        }
        const uint32_t *wp = &trace[(((size_t)s * T + (jt + l)) * 32 + l) * pc + jc];
        ISOF_ASSERT(wp >= trace && wp < tail - TRACE_SLACK && wp < tail - nr, CHK_TRACE_LOAD);
        last_read = wp;
        const uint32_t word = __ldcg(wp);
        const int nib = split ? ((((word >> (16 + 2 * r)) & 3) << 2) | ((word >> (2 * r)) & 3)) : ((word >> (4 * r)) & 15);
        if (where == 3) {
            const int src = nib >> 2;
            if (src == 3) {
                PUSH(upc(s1[i]) == upc(s2[j]) ? ISOF_CIGAR_EQ : ISOF_CIGAR_X, 1);
                --i; --j;
            } else where = src;
        } else if (where == 1) {          // E: consumes the reference -> 'D'
            PUSH(ISOF_CIGAR_D, 1);
            --j;
            where = (nib & 1) ? 1 : 3;
        } else {                          // F: consumes the query -> 'I'
            PUSH(ISOF_CIGAR_I, 1);
            --i;
            where = (nib & 2) ? 2 : 3;
        }
    }
    if (cur_len) { ++nr; ISOF_ASSERT(tail - nr >= trace, CHK_TAIL_OVERLAP); tail[-nr] = (cur_len << 4) | cur_op; }
#undef PUSH
    P.nruns[p] = nr;
}

// after the traceback of a small batch: CIGAR offsets (block scan of the run counts), coalesced copy of the runs,
// running total -- the job of nruns_to_i64 + DeviceScan + cigar_copy_kernel + advance_base_kernel
__global__ void __launch_bounds__(1024)
small_finish_kernel(const AlignParams P, const int n_pairs, long long *__restrict__ run_base, uint32_t *__restrict__ cigar,
                    long long cigar_cap, long long *__restrict__ cigar_off)
{
    typedef cub::BlockScan<long long, 1024> Scan;
    __shared__ typename Scan::TempStorage scan_tmp;
    __shared__ long long s_off[SMALL_P + 1];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    long long nr = (tid < n_pairs) ? (long long)P.nruns[tid] : 0, excl = 0, total = 0;
    Scan(scan_tmp).ExclusiveSum(nr, excl, total);
    s_off[tid] = excl;
    if (tid == 0) s_off[n_pairs] = total;
    __syncthreads();
    if (cigar_off && tid <= n_pairs) cigar_off[tid] = s_off[tid];
    if (tid == 0) *run_base = total;
    if (!cigar) return;
    for (int p = warp; p < n_pairs; p += 32) {
        const int n = P.nruns[p];
        const long long dst = s_off[p];
        const uint32_t *src = P.trace + (P.ptrace[p + 1] - P.chunk_base) - n;
        ISOF_ASSERT(src >= P.trace + (P.ptrace[p] - P.chunk_base), CHK_CIGAR_COPY);
        for (int x = lane; x < n; x += 32)
            if (dst + x < cigar_cap) cigar[dst + x] = src[x];
    }
}

// local exclusive offsets (within the chunk) come from cub; *run_base is the running total
__global__ void cigar_copy_kernel(const AlignParams P, const int64_t start, const int64_t end,
                                  const int64_t *__restrict__ local_off, const int64_t *__restrict__ run_base,
                                  uint32_t *__restrict__ cigar, int64_t cigar_cap, int64_t *__restrict__ cigar_off)
{
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t p = start + warp;
    if (p >= end) return;
    const int64_t dst = *run_base + local_off[p - start];
    const int nr = P.nruns[p];
    if (lane == 0 && cigar_off) cigar_off[p] = dst;
    if (!cigar) return;
    const uint32_t *src = P.trace + (P.ptrace[p + 1] - P.chunk_base) - nr;
    ISOF_ASSERT(src >= P.trace + (P.ptrace[p] - P.chunk_base), CHK_CIGAR_COPY);
    for (int x = lane; x < nr; x += 32)
        if (dst + x < cigar_cap) cigar[dst + x] = src[x];
}

__global__ void advance_base_kernel(int64_t *run_base, const int64_t *local_off, const int32_t *nruns, int64_t start,
                                    int64_t end, int64_t *cigar_off, int64_t P)
{
    if (end > start) *run_base += local_off[end - 1 - start] + nruns[end - 1];
    if (cigar_off && end == P) cigar_off[P] = *run_base;
}

__global__ void nruns_to_i64_kernel(const int32_t *nruns, int64_t start, int64_t count, int64_t *out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = nruns[start + i];
}

// ---------------------------------------------------------------------------------------------
// fused epilogue: integer features of the reference's CIGAR decisions (one thread per pair)
// ---------------------------------------------------------------------------------------------
__global__ void sg_features_kernel(const AlignParams P, const int64_t start, const int64_t end, int delta_len,
                                   int32_t *__restrict__ feat)
{
    const int64_t p = start + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= end) return;
    const int n = P.pn[p], m = P.pm[p];
    const uint8_t *s1 = P.cat + P.seq_off[P.pa[p]];
    const uint8_t *s2 = P.cat + P.seq_off[P.pb[p]];
    const int nr = P.nruns[p];
    const uint32_t *runs = P.trace + (P.ptrace[p + 1] - P.chunk_base) - nr;
    int32_t *f = feat + p * ISOF_N_FEATURES;
    // pass 1: parse_cigar_diversity features (SimplifyGraph.py:175-196)
    int alen = 0, nonmatch = 0, maxrun = 0;
    for (int x = 0; x < nr; ++x) {
        const uint32_t rr = runs[x];
        const int len = (int)(rr >> 4), op = (int)(rr & 15);
        alen += len;
        if (op != ISOF_CIGAR_EQ && op != ISOF_CIGAR_M) { nonmatch += len; maxrun = max(maxrun, len); }
    }
    // pass 2: find_first_significant_match forward (IsoformGeneration.py:370-377): window 30, > 0.7
    const int WIN = 30, NEED = 22;
    int start_match = -1, end_match = -1;
    {
        int qi = 0, ri = 0, col = 0;
        uint32_t bits = 0;
        for (int x = 0; x < nr && start_match < 0; ++x) {
            const uint32_t rr = runs[x];
            const int len = (int)(rr >> 4), op = (int)(rr & 15);
            for (int y = 0; y < len; ++y) {
                int eq;
                if (op == ISOF_CIGAR_EQ || op == ISOF_CIGAR_X || op == ISOF_CIGAR_M) { eq = s1[qi] == s2[ri]; ++qi; ++ri; }
                else if (op == ISOF_CIGAR_I) { eq = s1[qi] == '-'; ++qi; }
                else { eq = s2[ri] == '-'; ++ri; }
                bits = ((bits << 1) | (uint32_t)eq) & 0x3fffffffu;
                if (col >= WIN - 1 && __popc(bits) >= NEED) { start_match = col - (WIN - 1); break; }
                ++col;
            }
        }
    }
    {
        int qi = n - 1, ri = m - 1, col = 0;
        uint32_t bits = 0;
        for (int x = nr - 1; x >= 0 && end_match < 0; --x) {
            const uint32_t rr = runs[x];
            const int len = (int)(rr >> 4), op = (int)(rr & 15);
            for (int y = 0; y < len; ++y) {
                int eq;
                if (op == ISOF_CIGAR_EQ || op == ISOF_CIGAR_X || op == ISOF_CIGAR_M) { eq = s1[qi] == s2[ri]; --qi; --ri; }
                else if (op == ISOF_CIGAR_I) { eq = s1[qi] == '-'; --qi; }
                else { eq = s2[ri] == '-'; --ri; }
                bits = ((bits << 1) | (uint32_t)eq) & 0x3fffffffu;
                if (col >= WIN - 1 && __popc(bits) >= NEED) { end_match = col - (WIN - 1); break; }
                ++col;
            }
        }
    }
    // pass 3: parse_cigar_diversity_isoform_level_new integers (IsoformGeneration.py:251-289)
    int structural = 0, miss_runs = 0, bm = 0, bn = 0, am = 0, an = 0;
    if (start_match >= 0 && end_match >= 0) {
        const int first = start_match, last = alen - end_match;
        int pos = 0;
        for (int x = 0; x < nr; ++x) {
            const uint32_t rr = runs[x];
            const int len = (int)(rr >> 4), op = (int)(rr & 15);
            const int st = pos;
            pos += len;
            if (op != ISOF_CIGAR_EQ && op != ISOF_CIGAR_M) {
                if (st < first) { if (op != ISOF_CIGAR_D) bn += len; }
                else if (st >= last || st + len >= last) { if (op != ISOF_CIGAR_D) an += len; }
                else if (len > delta_len) structural = 1;
                else ++miss_runs;
            } else {
                if (st < first) bm += len;
                else if (st >= last) am += len;
            }
        }
    }
    f[0] = alen; f[1] = nr; f[2] = nonmatch; f[3] = maxrun; f[4] = start_match; f[5] = end_match;
    f[6] = structural; f[7] = miss_runs; f[8] = bm; f[9] = bn; f[10] = am; f[11] = an;
}

}  // namespace

unsigned isof_viol_align() { return isof_read_violations_tu(); }

// ---------------------------------------------------------------------------------------------
// host-side orchestration (all pointers are device pointers)
// ---------------------------------------------------------------------------------------------
int isof_sg_align_core(isof_ctx *ctx, const uint8_t *d_cat, const int64_t *d_off, int32_t n_seqs, int64_t n_bases,
                       const int32_t *d_pa, const int32_t *d_pb, int64_t P, int32_t match, int32_t mismatch,
                       int32_t gap_open, int32_t gap_ext, int32_t *d_score, int32_t *d_endq, int32_t *d_endr,
                       uint32_t *d_cigar, int64_t cigar_cap, int64_t *d_cigar_off, int32_t *d_features,
                       int32_t delta_len, int64_t *h_total_runs)
{
    cudaStream_t st = ctx->stream;
    if (h_total_runs) *h_total_runs = 0;
    // ---- eligibility of the packed 16-bit path (4*score + 2 tag bits must fit a signed half):
    //   positive side  4*match*min(n,m) + 16 <= 32767
    //   negative side  4*(2*open + ext*max(n,m) + |mismatch| + |match|) <= 31000   (sentinel is -32000)
    //   substitution   (1 << CS16) >= 4*(match - mismatch)
    int h_max_min_len = 0, h_max_max_len = 0;
    if (!ctx->disable_packed && match > 0 && mismatch <= match && 4 * (match - mismatch) <= (1 << CS16)) {
        h_max_min_len = (32767 - 16) / (4 * match);
        const int room = 31000 / 4 - 2 * gap_open - std::abs(mismatch) - std::abs(match);
        h_max_max_len = room <= 0 ? 0 : (gap_ext > 0 ? room / gap_ext : INT_MAX);
    }
    if (P < 0 || n_seqs < 0) return isof_fail(ctx, ISOF_ERR_ARG, "negative count");
    if (match - mismatch > 256 || match - mismatch < 0 || gap_open < 0 || gap_ext < 0 ||
        std::abs(match) > 1024 || std::abs(mismatch) > 1024 || gap_open > 4096 || gap_ext > 4096)
        return isof_fail(ctx, ISOF_ERR_RANGE, "scoring parameters outside the supported range");
    if (P == 0) {
        if (d_cigar_off) CUDA_TRY(ctx, cudaMemsetAsync(d_cigar_off, 0, sizeof(int64_t), st));
        return ISOF_OK;
    }
    // ---- scratch
    TRY(dev_reserve(ctx, ctx->b_codes, (size_t)n_bases + 16));
    TRY(dev_reserve(ctx, ctx->b_seqflags, (size_t)n_seqs + 16));
    TRY(dev_reserve(ctx, ctx->b_pn, sizeof(int32_t) * (size_t)P));
    TRY(dev_reserve(ctx, ctx->b_pm, sizeof(int32_t) * (size_t)P));
    TRY(dev_reserve(ctx, ctx->b_status, (size_t)P));
    TRY(dev_reserve(ctx, ctx->b_words, sizeof(int64_t) * (size_t)(P + 1)));      // trace words per pair
    TRY(dev_reserve(ctx, ctx->b_ptrace, sizeof(int64_t) * (size_t)(P + 1)));     // exclusive scan
    TRY(dev_reserve(ctx, ctx->b_nruns, sizeof(int32_t) * (size_t)P));
    TRY(dev_reserve(ctx, ctx->b_runoff, sizeof(int64_t) * (size_t)(P + 1)));
    TRY(dev_reserve(ctx, ctx->b_nruns64, sizeof(int64_t) * (size_t)(P + 1)));      // nruns as int64
    TRY(dev_reserve(ctx, ctx->b_misc, 64));
    uint8_t *codes = (uint8_t *)ctx->b_codes.p, *flags = (uint8_t *)ctx->b_seqflags.p;
    int32_t *pn = (int32_t *)ctx->b_pn.p, *pm = (int32_t *)ctx->b_pm.p, *nruns = (int32_t *)ctx->b_nruns.p;
    uint8_t *pwild = (uint8_t *)ctx->b_status.p;
    int64_t *words = (int64_t *)ctx->b_words.p, *ptrace = (int64_t *)ctx->b_ptrace.p;
    int64_t *runoff = (int64_t *)ctx->b_runoff.p, *nruns64 = (int64_t *)ctx->b_nruns64.p;
    unsigned long long *misc = (unsigned long long *)ctx->b_misc.p;   // [0] err [1] max m [2] cells [3] padded [4] run_base
    // small batches: one fused set-up kernel (codes, plan, trace offsets, task order) and one fused CIGAR
    // scan/copy kernel; the layout of b_counters below is the single-chunk one
    const bool small = P <= SMALL_P && n_bases <= SMALL_BASES;
    TRY(dev_reserve(ctx, ctx->b_order, sizeof(int32_t) * (size_t)P));
    if (small) {
        TRY(dev_reserve(ctx, ctx->b_counters, sizeof(int64_t) * 3 * 3));
        unsigned long long *cnt0 = (unsigned long long *)ctx->b_counters.p + 3;          // counters[0] of a 1-chunk layout
        LaunchTimer t(ctx, SLOT_OTHER);
        small_plan_kernel<<<1, 1024, 0, st>>>(d_cat, d_off, n_seqs, d_pa, d_pb, (int)P, codes, flags, pn, pm, pwild, ptrace,
                                              misc, h_max_min_len, h_max_max_len, (int32_t *)ctx->b_order.p, cnt0 + 3, cnt0);
    } else {
        CUDA_TRY(ctx, cudaMemsetAsync(misc, 0, 64, st));
        CUDA_TRY(ctx, cudaMemsetAsync(pwild, 0, (size_t)P, st));
        {
            LaunchTimer t(ctx, SLOT_OTHER);
            int blocks = std::min((n_seqs * 32 + 255) / 256, ctx->sm_count * 8);
            encode_kernel<<<std::max(blocks, 1), 256, 0, st>>>(d_cat, d_off, n_seqs, codes, flags);
        }
        {
            LaunchTimer t(ctx, SLOT_OTHER);
            plan_kernel<<<(unsigned)((P + 1 + 255) / 256), 256, 0, st>>>(d_off, n_seqs, flags, d_pa, d_pb, P, pn, pm, pwild,
                                                                          words, misc, h_max_min_len, h_max_max_len);
        }
        CUDA_TRY(ctx, cudaGetLastError());
        size_t tmp_bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, words, ptrace, P + 1, st);
        TRY(dev_reserve(ctx, ctx->b_scan_tmp, tmp_bytes));
        {
            LaunchTimer t(ctx, SLOT_OTHER);
            CUDA_TRY(ctx, cub::DeviceScan::ExclusiveSum(ctx->b_scan_tmp.p, tmp_bytes, words, ptrace, P + 1, st));
        }
    }
    CUDA_TRY(ctx, cudaGetLastError());
    // ---- read back the plan summary
    unsigned long long *pin = (unsigned long long *)ctx->pinned;
    CUDA_TRY(ctx, cudaMemcpyAsync(pin, misc, 32, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(pin + 4, ptrace + P, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    if (pin[0] & 1) return isof_fail(ctx, ISOF_ERR_ARG, "pair index outside the sequence pool");
    if (pin[0] & 2) return isof_fail(ctx, ISOF_ERR_ARG, "empty sequence in a pair (parasail rejects it too)");
    const int max_m = (int)pin[1];
    const int64_t total_words = (int64_t)pin[4];
    ctx->prof.dp_cells += (int64_t)pin[2];
    ctx->prof.dp_cells_padded += (int64_t)pin[3];
    ctx->prof.trace_bytes += (int64_t)pin[3] / 2;                 // 4 bits per updated cell
    if ((int64_t)max_m * 16 * (std::abs(match) + std::abs(mismatch) + gap_ext) + 16LL * gap_open > (1LL << 28))
        return isof_fail(ctx, ISOF_ERR_RANGE, "sequence too long for the 32-bit tagged score range");

    // ---- trace budget / chunks.  Default: 40 % of the device memory that was free at first use; raised up to
    // 55 % when the pairs are so long (10 kb reads: 36 MB of trace each) that a smaller buffer could not hold
    // enough pairs to occupy every resident warp in each of the NLANES chunks in flight.
    if (ctx->trace_free_at_first_use == 0) {
        size_t fr = 0, tot = 0;
        CUDA_TRY(ctx, cudaMemGetInfo(&fr, &tot));
        ctx->trace_free_at_first_use = fr + ctx->b_trace.cap;
    }
    size_t budget_bytes = ctx->trace_budget;
    if (budget_bytes == 0) {
        const double fr = (double)ctx->trace_free_at_first_use;
        const double want = 4.0 * (double)total_words / (double)P * (double)(ctx->sm_count * DP_CTAS_PER_SM * DP_WARPS) * 1.25 *
                            (double)isof_ctx::NLANES;
        budget_bytes = (size_t)std::min(0.55 * fr, std::max(0.40 * fr, want));
        if (budget_bytes < (64u << 20)) budget_bytes = 64u << 20;
    }
    int64_t budget_words = (int64_t)(budget_bytes / 4);
    // pipelined mode: NLANES smaller buffers, chunk c runs on stream c % NLANES
    constexpr int NL = isof_ctx::NLANES;
    const bool want_overlap = ctx->overlap && total_words > budget_words / NL;
    if (want_overlap) budget_words /= NL;
    if (budget_words > total_words) budget_words = total_words;
    if (budget_words < 1) budget_words = 1;
    const bool single = total_words <= budget_words;          // everything fits one chunk: no device round trips
    const int n_chunks = single ? 1 : (int)(total_words / budget_words) + 1;
    TRY(dev_reserve(ctx, ctx->b_counters, sizeof(int64_t) * (size_t)(n_chunks + 2) * 3));
    int64_t *chunk_start_d = (int64_t *)ctx->b_counters.p;
    unsigned long long *counters = (unsigned long long *)(chunk_start_d + n_chunks + 2);      // task tickets per chunk
    unsigned long long *elig_counts = counters + n_chunks + 2;                                  // eligible pairs per chunk
    const bool fused_small = small && single;          // tickets, eligible count and task order came from small_plan_kernel
    if (!fused_small) CUDA_TRY(ctx, cudaMemsetAsync(counters, 0, sizeof(int64_t) * (size_t)(n_chunks + 2) * 2, st));
    std::vector<int64_t> chunk_start((size_t)n_chunks + 1), start_off((size_t)n_chunks + 1);
    if (single) {
        chunk_start[0] = 0; chunk_start[1] = P;
        start_off[0] = 0; start_off[1] = total_words;
    } else {
        {
            LaunchTimer t(ctx, SLOT_OTHER);
            chunk_bounds_kernel<<<(n_chunks + 1 + 127) / 128, 128, 0, st>>>(ptrace, P, budget_words, n_chunks, chunk_start_d);
        }
        CUDA_TRY(ctx, cudaMemcpyAsync(chunk_start.data(), chunk_start_d, sizeof(int64_t) * (size_t)(n_chunks + 1),
                                      cudaMemcpyDeviceToHost, st));
        CUDA_TRY(ctx, cudaStreamSynchronize(st));
        // the largest chunk decides the scratch size: offsets of chunk starts are needed on the host
        for (int c = 0; c <= n_chunks; ++c)
            CUDA_TRY(ctx, cudaMemcpyAsync(&start_off[c], ptrace + chunk_start[c], 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(ctx, cudaStreamSynchronize(st));
    }
    int64_t max_chunk_words = 0;
    for (int c = 0; c < n_chunks; ++c) max_chunk_words = std::max(max_chunk_words, start_off[c + 1] - start_off[c]);
    const bool pipelined = want_overlap && n_chunks > 1;
    for (int w = 0; w < (pipelined ? NL : 1); ++w) {
        DevBuf *tb = w ? &ctx->b_trace_x[w - 1] : &ctx->b_trace;
        if ((size_t)max_chunk_words * 4 > tb->cap) {
            int rc = dev_reserve(ctx, *tb, (size_t)max_chunk_words * 4);
            if (rc != ISOF_OK)
                return isof_fail(ctx, ISOF_ERR_NOMEM, "trace scratch of %lld bytes does not fit (largest pair/chunk)",
                                 (long long)max_chunk_words * 4);
        }
    }
    const int occ_blocks = ctx->sm_count * DP_CTAS_PER_SM;   // resident CTAs (persistent: one launch fills the GPU once)
    const int64_t bnd_stride = ((int64_t)max_m + 128) & ~63LL;
    TRY(dev_reserve(ctx, ctx->b_bnd, sizeof(int2) * (size_t)bnd_stride * (size_t)occ_blocks * DP_WARPS));
    if (pipelined) {
        for (int w = 0; w < NL - 1; ++w) {
            TRY(dev_reserve(ctx, ctx->b_bnd_x[w], sizeof(int2) * (size_t)bnd_stride * (size_t)occ_blocks * DP_WARPS));
            TRY(dev_reserve(ctx, ctx->b_runoff_x[w], sizeof(int64_t) * (size_t)(P + 1)));
            TRY(dev_reserve(ctx, ctx->b_nruns64_x[w], sizeof(int64_t) * (size_t)(P + 1)));
        }
    }

    AlignParams A;
    A.cat = d_cat; A.codes = codes; A.seq_off = d_off; A.pa = d_pa; A.pb = d_pb; A.pn = pn; A.pm = pm; A.pwild = pwild;
    A.ptrace = ptrace; A.trace = (uint32_t *)ctx->b_trace.p; A.bnd = (int2 *)ctx->b_bnd.p; A.bnd_stride = bnd_stride;
    A.score = d_score; A.endq = d_endq; A.endr = d_endr; A.nruns = nruns;
    const int S = 1 << TAG_BITS;
    A.c_eo = -gap_open * S + 4;
    A.c_ee = -gap_ext * S + 5;
    A.c_fo = -gap_open * S + 8;
    A.c_fe = -gap_ext * S + 10;
    A.c_sm = match * S + 12 - (3 << CODE_SHIFT);
    A.c_sx = mismatch * S + 12;
    auto pk = [](int v) { return ((uint32_t)(v & 0xffff)) | ((uint32_t)(v & 0xffff) << 16); };
    A.h_eo = pk(-4 * gap_open);
    A.h_ee = pk(-4 * gap_ext);
    A.h_fo = pk(-4 * gap_open);
    A.h_fe = pk(-4 * gap_ext);
    A.h_sm = pk(4 * match + 3 - (3 << CS16));
    A.h_sx = pk(4 * mismatch + 3);
    TRY(dev_reserve(ctx, ctx->b_layout, (size_t)P * 2));
    A.playout = (uint8_t *)ctx->b_layout.p;
    A.plastrr = A.playout + P;
    A.h_max_min_len = h_max_min_len;
    A.h_max_max_len = h_max_max_len;
    A.chunk_words = 0;
    int64_t *run_base = (int64_t *)(misc + 4);
    CUDA_TRY(ctx, cudaMemsetAsync(run_base, 0, 8, st));

    // ---- reserve every per-lane scratch for the largest chunk up front: a reallocation inside the loop would
    // free memory that an earlier chunk of the same lane may still be using and stall the pipeline
    {
        int64_t max_cnt = 1;
        for (int c = 0; c < n_chunks; ++c) max_cnt = std::max(max_cnt, chunk_start[c + 1] - chunk_start[c]);
        size_t t_sort = 0, t_scan = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, t_sort, (uint32_t *)nullptr, (uint32_t *)nullptr, (int32_t *)nullptr,
                                        (int32_t *)nullptr, (int)max_cnt, 0, 32, st);
        cub::DeviceScan::ExclusiveSum(nullptr, t_scan, (int64_t *)nullptr, (int64_t *)nullptr, max_cnt, st);
        for (int w = 0; w < (pipelined ? NL : 1); ++w) {
            TRY(dev_reserve(ctx, w ? ctx->b_taskkey_x[w - 1] : ctx->b_taskkey,
                            (sizeof(uint32_t) * 2 + sizeof(int32_t)) * (size_t)max_cnt + 64));
            TRY(dev_reserve(ctx, w ? ctx->b_scan_tmp_x[w - 1] : ctx->b_scan_tmp, std::max(t_sort, t_scan) + 256));
        }
    }
    // ---- chunk pipeline.  Serial mode: everything on `st`.  Pipelined mode: chunk c runs on stream c % NLANES with
    // its own trace / boundary / scan scratch; the only cross-chunk dependency is the running CIGAR offset
    // (cigar_copy of chunk c needs advance_base of chunk c-1), carried by one event per chunk.
    cudaStream_t lanes[NL];
    lanes[0] = st;
    for (int w = 1; w < NL; ++w) lanes[w] = pipelined ? ctx->aux_stream[w - 1] : st;
    std::vector<cudaEvent_t> done((size_t)n_chunks, nullptr);
    cudaEvent_t fork = nullptr;
    if (pipelined) {
        CUDA_TRY(ctx, cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
        CUDA_TRY(ctx, cudaEventRecord(fork, st));
        for (int w = 1; w < NL; ++w) CUDA_TRY(ctx, cudaStreamWaitEvent(lanes[w], fork, 0));
    }
    int prev = -1, live = 0;
    for (int c = 0; c < n_chunks; ++c) {
        const int64_t ps = chunk_start[c], pe = chunk_start[c + 1];
        if (pe <= ps) continue;
        const int64_t cnt = pe - ps;
        const int w = pipelined ? (live % NL) : 0;
        ++live;
        cudaStream_t cs = lanes[w];
        A.chunk_base = start_off[c];
        A.trace = (uint32_t *)(w ? ctx->b_trace_x[w - 1].p : ctx->b_trace.p);
        A.bnd = (int2 *)(w ? ctx->b_bnd_x[w - 1].p : ctx->b_bnd.p);
        int64_t *c_runoff = w ? (int64_t *)ctx->b_runoff_x[w - 1].p : runoff;
        int64_t *c_nruns64 = w ? (int64_t *)ctx->b_nruns64_x[w - 1].p : nruns64;
        DevBuf &c_scan = w ? ctx->b_scan_tmp_x[w - 1] : ctx->b_scan_tmp;
        if (fused_small) {
            A.order = (int32_t *)ctx->b_order.p;
            A.n_elig = elig_counts + c;
        } else {   // task order of the chunk: stable radix sort of the pair ids by (eligibility, reference)
            DevBuf &kb = w ? ctx->b_taskkey_x[w - 1] : ctx->b_taskkey;
            TRY(dev_reserve(ctx, kb, (sizeof(uint32_t) * 2 + sizeof(int32_t)) * (size_t)cnt + 64));
            uint32_t *key_in = (uint32_t *)kb.p, *key_out = key_in + cnt;
            int32_t *val_in = (int32_t *)(key_out + cnt);
            int32_t *order = (int32_t *)ctx->b_order.p;
            LaunchTimer t(ctx, SLOT_OTHER, cs);
            task_key_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, cs>>>(pwild, d_pb, ps, cnt, key_in, val_in, elig_counts + c);
            size_t tso = 0;
            cub::DeviceRadixSort::SortPairs(nullptr, tso, key_in, key_out, val_in, order + ps, (int)cnt, 0, 32, cs);
            TRY(dev_reserve(ctx, c_scan, tso));
            CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(c_scan.p, tso, key_in, key_out, val_in, order + ps, (int)cnt, 0, 32, cs));
            A.order = order;
            A.n_elig = elig_counts + c;
        }
        {
            LaunchTimer t(ctx, SLOT_DP, cs);
            const int64_t tasks = cnt;               // upper bound on tickets (all pairs single); CTAs are persistent anyway
            if (ctx->profiling)
                packed_cells_kernel<<<(unsigned)((tasks + 255) / 256), 256, 0, cs>>>(pwild, pn, pm, d_pb, A.order, A.n_elig, ps, pe, misc + 5);
            int blocks = (int)std::min<int64_t>((tasks + DP_WARPS - 1) / DP_WARPS, occ_blocks);
            sg_dp_kernel<<<blocks, DP_WARPS * 32, 0, cs>>>(A, ps, pe, counters + c);
        }
        {
            LaunchTimer t(ctx, SLOT_TB, cs);
            sg_traceback_kernel<<<(unsigned)((cnt + 63) / 64), 64, 0, cs>>>(A, ps, pe);
        }
        CUDA_TRY(ctx, cudaGetLastError());
        if (d_features) {
            LaunchTimer t(ctx, SLOT_TB, cs);
            sg_features_kernel<<<(unsigned)((cnt + 63) / 64), 64, 0, cs>>>(A, ps, pe, delta_len, d_features);
        }
        if (fused_small) {
            LaunchTimer t(ctx, SLOT_TB, cs);
            small_finish_kernel<<<1, 1024, 0, cs>>>(A, (int)P, (long long *)run_base, d_cigar, (long long)cigar_cap,
                                                    (long long *)d_cigar_off);
        } else {
            {
                LaunchTimer t(ctx, SLOT_TB, cs);
                nruns_to_i64_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, cs>>>(nruns, ps, cnt, c_nruns64);
            }
            size_t tb2 = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, tb2, c_nruns64, c_runoff, cnt, cs);
            TRY(dev_reserve(ctx, c_scan, tb2));
            {
                LaunchTimer t(ctx, SLOT_TB, cs);
                CUDA_TRY(ctx, cub::DeviceScan::ExclusiveSum(c_scan.p, tb2, c_nruns64, c_runoff, cnt, cs));
            }
            if (pipelined && prev >= 0) CUDA_TRY(ctx, cudaStreamWaitEvent(cs, done[prev], 0));
            {
                LaunchTimer t(ctx, SLOT_TB, cs);
                cigar_copy_kernel<<<(unsigned)((cnt * 32 + 255) / 256), 256, 0, cs>>>(A, ps, pe, c_runoff, run_base, d_cigar,
                                                                                     cigar_cap, d_cigar_off);
            }
            {
                LaunchTimer t(ctx, SLOT_TB, cs);
                advance_base_kernel<<<1, 1, 0, cs>>>(run_base, c_runoff, nruns, ps, pe, d_cigar_off, P);
            }
        }
        CUDA_TRY(ctx, cudaGetLastError());
        if (pipelined) {
            CUDA_TRY(ctx, cudaEventCreateWithFlags(&done[c], cudaEventDisableTiming));
            CUDA_TRY(ctx, cudaEventRecord(done[c], cs));
        }
        prev = c;
    }
    if (pipelined) {
        if (prev >= 0) CUDA_TRY(ctx, cudaStreamWaitEvent(st, done[prev], 0));      // join: the chain orders all chunks
        // the aux streams may still hold (ordered-before) earlier chunks: join them explicitly
        for (int w = 1; w < NL; ++w) {
            cudaEvent_t tail = nullptr;
            CUDA_TRY(ctx, cudaEventCreateWithFlags(&tail, cudaEventDisableTiming));
            CUDA_TRY(ctx, cudaEventRecord(tail, lanes[w]));
            CUDA_TRY(ctx, cudaStreamWaitEvent(st, tail, 0));
            cudaEventDestroy(tail);
        }
        cudaEventDestroy(fork);
        for (cudaEvent_t e : done) if (e) cudaEventDestroy(e);
    }
    int64_t *pin64 = (int64_t *)ctx->pinned;
    CUDA_TRY(ctx, cudaMemcpyAsync(pin64, run_base, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    if (h_total_runs) *h_total_runs = pin64[0];
    if (ctx->profiling) {
        CUDA_TRY(ctx, cudaMemcpyAsync(pin64 + 1, misc + 5, 16, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(ctx, cudaStreamSynchronize(st));
        ctx->prof.dp_cells_packed += pin64[1];
        ctx->prof.dp_cells_profile += pin64[2];
    }
    if (d_cigar && pin64[0] > cigar_cap)
        return isof_fail(ctx, ISOF_ERR_CAPACITY, "cigar_runs holds %lld runs, %lld needed", (long long)cigar_cap,
                         (long long)pin64[0]);
    return ISOF_OK;
}
