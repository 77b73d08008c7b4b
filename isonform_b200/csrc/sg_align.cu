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
#include "sg_dp.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/block/block_scan.cuh>
#include <climits>
#include <type_traits>

namespace {


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

// banded trace for a pair of the re-based class: keep the diagonals between the main one and the one the free end gaps
// make room for, +- band_w0; worth it only when it is well below the whole matrix.  Returns the stored steps per stripe
// (0 = not banded), sets the lowest kept diagonal and the trace words.
__device__ __forceinline__ int band_plan(const int n, const int m, const int ns, const int band_w0, int &dmin_out, long long &w)
{
    const int dmin = min(0, m - n) - band_w0, dmax = max(0, m - n) + band_w0;
    const int t_full = (m + 1) / 2 + 31, t_band = ((dmax - dmin + 320) >> 1) + 2;
    dmin_out = 0;
    if (band_w0 <= 0 || t_band + 64 >= t_full) return 0;
    dmin_out = dmin;
    w = (long long)ns * t_band * 64 + TRACE_SLACK;
    return t_band;
}

// per pair: lengths, wildcard flag, trace words; misc[0]=error flags, misc[1]=max m, misc[2..3]=cells
__global__ void plan_kernel(const int64_t *__restrict__ off, int n_seqs, const uint8_t *__restrict__ flags,
                            const int32_t *__restrict__ pa, const int32_t *__restrict__ pb, int64_t P,
                            int32_t *__restrict__ pn, int32_t *__restrict__ pm, uint8_t *__restrict__ pwild,
                            int64_t *__restrict__ words, unsigned long long *__restrict__ misc, int h_max_min_len,
                            int h_max_max_len, int short_max, int rb_ok, int band_w0, int32_t *__restrict__ pband)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long cells = 0, padded = 0;
    int mm = 0, mn = 0, not_short = 0, any_wild = 0;
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
                // bit0: a wildcard is present (32-bit WILD path); bit1: eligible for the packed 16-bit path; bit2: beyond
                // the plain 16-bit range but fine for the re-based 16-bit lanes (long reads)
                const int wild = flags[a] | flags[b];
                const int elig = !wild && min(n, m) <= h_max_min_len && max(n, m) <= h_max_max_len;
                const int elig_rb = !wild && !elig && rb_ok;
                // rows actually updated: full stripes + the compacted last stripe (2/4/6/8 rows per lane)
                const int need = n - (ns - 1) * SROWS;
                const int rr = need > 192 ? 8 : need > 128 ? 6 : need > 64 ? 4 : 2;
                padded = (unsigned long long)((ns - 1) * SROWS + ((elig | elig_rb) ? rr * 32 : SROWS)) * m;
                pwild[p] = (uint8_t)(wild | (elig << 1) | (elig_rb << 2));
                if (pband) {
                    int bd = 0, tw = 0;
                    long long wb = w;
                    if (elig_rb) tw = band_plan(n, m, ns, band_w0, bd, wb);
                    w = wb;
                    pband[2 * p] = bd; pband[2 * p + 1] = tw;
                }
                mm = m; mn = n;
                // short-pair regime (sg_short.cu): both lengths small and inside the 16-bit range (wildcards allowed)
                not_short = !(max(n, m) <= short_max && min(n, m) <= h_max_min_len && max(n, m) <= h_max_max_len);
                any_wild = wild;
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
        mn = max(mn, __shfl_down_sync(0xffffffffu, mn, d));
        not_short += __shfl_down_sync(0xffffffffu, not_short, d);
        any_wild |= __shfl_down_sync(0xffffffffu, any_wild, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (cells) { atomicAdd(&misc[2], cells); atomicAdd(&misc[3], padded); }
        if (mm) atomicMax(&misc[1], (unsigned long long)mm);
        if (mn) atomicMax(&misc[8], (unsigned long long)mn);
        if (not_short) atomicAdd(&misc[9], (unsigned long long)not_short);
        if (any_wild) atomicOr(&misc[10], 1ull);
    }
    if (p < P && pband && pn[p] == 0) { pband[2 * p] = 0; pband[2 * p + 1] = 0; }
    if (p == P) words[P] = 0;
}

// pairs whose traceback left the band (nruns == -1): their ids, densely
__global__ void collect_failed_kernel(const int32_t *__restrict__ nruns, int64_t P, int32_t *__restrict__ ids,
                                      unsigned long long *__restrict__ count)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < P && nruns[p] < 0) ids[atomicAdd(count, 1ull)] = (int32_t)p;
}

__global__ void gather_pairs_kernel(const int32_t *__restrict__ ids, int64_t n, const int32_t *__restrict__ pa,
                                    const int32_t *__restrict__ pb, int32_t *__restrict__ pa2, int32_t *__restrict__ pb2)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) { pa2[k] = pa[ids[k]]; pb2[k] = pb[ids[k]]; }
}

__global__ void scatter_results_kernel(const int32_t *__restrict__ ids, int64_t n, const int32_t *__restrict__ score2,
                                       const int32_t *__restrict__ endq2, const int32_t *__restrict__ endr2,
                                       const int32_t *__restrict__ feat2, int32_t *__restrict__ score, int32_t *__restrict__ endq,
                                       int32_t *__restrict__ endr, int32_t *__restrict__ feat)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int64_t p = ids[k];
    score[p] = score2[k]; endq[p] = endq2[k]; endr[p] = endr2[k];
    if (feat) for (int q = 0; q < ISOF_N_FEATURES; ++q) feat[p * ISOF_N_FEATURES + q] = feat2[k * ISOF_N_FEATURES + q];
}

// banded second pass of a call that returns the runs: pair p's runs move from their place in the first pass's stream
// (`old_stream` + old_off[p]) -- or, for a re-aligned pair, from the second pass's stream -- to the rebuilt offsets
__global__ void mark_failed_kernel(const int32_t *__restrict__ ids, int64_t nf, int32_t *__restrict__ kidx)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nf) kidx[ids[k]] = (int32_t)k;
}

__global__ void splice_runs_kernel(const uint32_t *__restrict__ old_stream, const int64_t *__restrict__ old_off,
                                   const uint32_t *__restrict__ cig2, const int64_t *__restrict__ off2,
                                   const int32_t *__restrict__ kidx, const int64_t *__restrict__ new_off, int64_t P,
                                   uint32_t *__restrict__ out)
{
    const int64_t p = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= P) return;
    const int lane = threadIdx.x & 31;
    const int k = kidx[p];
    const uint32_t *src = k >= 0 ? cig2 + off2[k] : old_stream + old_off[p];
    uint32_t *dst = out + new_off[p];
    const int64_t len = new_off[p + 1] - new_off[p];
    for (int64_t x = lane; x < len; x += 32) dst[x] = src[x];
}

// sort key of a pair inside its chunk: packed-eligible pairs first, then by reference index, so that the two
// alignments of a task share their reference whenever the batch allows it (query-profile path)
__global__ void task_key_kernel(const uint8_t *__restrict__ pwild, const int32_t *__restrict__ pb, int64_t start,
                                int64_t cnt, uint32_t *__restrict__ key, int32_t *__restrict__ val,
                                unsigned long long *__restrict__ n_elig, unsigned long long *__restrict__ n_rb)
{
    const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int el = 0, rb = 0;
    if (x < cnt) {
        const int64_t p = start + x;
        el = (pwild[p] & 2) ? 1 : 0;
        rb = (pwild[p] & 4) ? 1 : 0;
        // class 0: plain 16-bit, class 1: re-based 16-bit, class 2: 32-bit; inside a class by reference index
        key[x] = (el ? 0u : rb ? 0x40000000u : 0x80000000u) | ((uint32_t)pb[p] & 0x3fffffffu);
        val[x] = (int32_t)p;
    }
    const unsigned m = __ballot_sync(0xffffffffu, el), m1 = __ballot_sync(0xffffffffu, rb);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_elig, (unsigned long long)__popc(m));
    if ((threadIdx.x & 31) == 0 && m1) atomicAdd(n_rb, (unsigned long long)__popc(m1));
}

// statistics only: cells of the pairs that take the packed path in chunk [start,end)
__global__ void packed_cells_kernel(const uint8_t *__restrict__ pwild, const int32_t *__restrict__ pn,
                                    const int32_t *__restrict__ pm, const int32_t *__restrict__ pb,
                                    const int32_t *__restrict__ order, const unsigned long long *__restrict__ n_elig,
                                    const unsigned long long *__restrict__ n_rb, int64_t start, int64_t end,
                                    unsigned long long *__restrict__ out)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n0 = (int64_t)(*n_elig), n1 = (int64_t)(*n_rb);
    // packed tasks of the plain class, then of the re-based class (which starts at n0 in the sorted order)
    // (the odd pair of a class runs as a task of its own, both halves on the same pair)
    const int64_t t0 = (n0 + 1) / 2;
    const bool plain = t < t0;
    const int64_t u = t - t0;
    const int64_t xA = plain ? start + 2 * t : start + n0 + 2 * u;
    const int64_t xB = (plain ? 2 * t + 1 < n0 : 2 * u + 1 < n1) ? xA + 1 : xA;
    unsigned long long c = 0, cp = 0;               // packed path / of those, the query-profile variant
    const int64_t pA = xA < end ? order[xA] : 0, pB = xB < end ? order[xB] : 0;
    if (xB < end && t < t0 + (n1 + 1) / 2) {
        c = (unsigned long long)pn[pA] * pm[pA] + (xB != xA ? (unsigned long long)pn[pB] * pm[pB] : 0ull);
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
                  unsigned long long *__restrict__ n_elig, unsigned long long *__restrict__ ticket, int short_max, int rb_ok,
                  unsigned long long *__restrict__ n_rb, unsigned long long *__restrict__ rb_ticket, int band_w0,
                  int32_t *__restrict__ pband)
{
    typedef cub::BlockScan<long long, 1024> Scan;
    __shared__ typename Scan::TempStorage scan_tmp;
    __shared__ uint32_t s_key[SMALL_P];
    __shared__ unsigned long long s_stat[7];             // err, max m, cells, padded, max n, not short, any wildcard
    __shared__ int s_nelig, s_nrb;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid < 7) s_stat[tid] = 0ull;
    if (tid == 0) { s_nelig = 0; s_nrb = 0; }
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
        int bdm = 0, btw = 0;
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
                const int elig_rb = !wild && !elig && rb_ok;
                const int need = n - (ns - 1) * SROWS;
                const int rr = need > 192 ? 8 : need > 128 ? 6 : need > 64 ? 4 : 2;
                atomicAdd(&s_stat[2], (unsigned long long)n * m);
                atomicAdd(&s_stat[3], (unsigned long long)((ns - 1) * SROWS + ((elig | elig_rb) ? rr * 32 : SROWS)) * m);
                atomicMax(&s_stat[1], (unsigned long long)m);
                atomicMax(&s_stat[4], (unsigned long long)n);
                if (!(max(n, m) <= short_max && min(n, m) <= h_max_min_len && max(n, m) <= h_max_max_len)) atomicAdd(&s_stat[5], 1ull);
                if (wild) atomicOr(&s_stat[6], 1ull);
                pw = (uint8_t)(wild | (elig << 1) | (elig_rb << 2));
                if (elig) atomicAdd(&s_nelig, 1);
                if (elig_rb) atomicAdd(&s_nrb, 1);
                key = (elig ? 0u : elig_rb ? 0x40000000u : 0x80000000u) | ((uint32_t)b & 0x3fffffffu);
                if (pband && elig_rb) btw = band_plan(n, m, ns, band_w0, bdm, w);
            }
        }
        pn[tid] = n; pm[tid] = m; pwild[tid] = pw;
        if (pband) { pband[2 * tid] = bdm; pband[2 * tid + 1] = btw; }
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
        misc[8] = s_stat[4]; misc[9] = s_stat[5]; misc[10] = s_stat[6];
        *n_elig = (unsigned long long)s_nelig;
        *ticket = 0ull;
        *n_rb = (unsigned long long)s_nrb;
        *rb_ticket = 0ull;
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


// banded second pass: run counts of the re-aligned pairs go back to their places, then the run offsets are rebuilt
__global__ void scatter_nruns_kernel(const int32_t *__restrict__ ids, int64_t nf, const int32_t *__restrict__ nruns2,
                                     int32_t *__restrict__ saved)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nf) saved[ids[k]] = nruns2[k];
}

__global__ void __launch_bounds__(1024) rebuild_offsets_kernel(const int32_t *__restrict__ nr, int64_t P, long long *__restrict__ cigar_off)
{
    typedef cub::BlockScan<long long, 1024> Scan;
    __shared__ typename Scan::TempStorage tmp;
    long long carry = 0;
    for (int64_t base = 0; base < P; base += 1024) {
        const int64_t p = base + threadIdx.x;
        long long v = p < P ? (long long)max(nr[p], 0) : 0, excl, total;
        Scan(tmp).ExclusiveSum(v, excl, total);
        if (p < P) cigar_off[p] = carry + excl;
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) cigar_off[P] = carry;
}

// ---------------------------------------------------------------------------------------------
// traceback -> reversed runs at the tail of the trace region
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int upc(int c) { return (c >= 'a' && c <= 'z') ? c - 32 : c; }

// SMEM: the trace lives in shared memory (fused small-call kernel): plain loads, no L2 prefetch
template <bool SMEM>
__device__ __forceinline__ void traceback_pair(const AlignParams &P, const int64_t p)
{
    const int n = P.pn[p], m = P.pm[p];
    const uint8_t *s1 = P.cat + P.seq_off[P.pa[p]];
    const uint8_t *s2 = P.cat + P.seq_off[P.pb[p]];
    const uint32_t *trace = P.trace + (P.ptrace[p] - P.chunk_base);
    uint32_t *tail = run_tail(P, p);                                  // runs go to tail[-1], tail[-2], ...
    const int lay = P.playout[p];
    const int pc = lay & 15;                          // columns per step of the writer
    const bool split = (lay & 16) != 0;               // packed path: [src x8 | fe x8] 2-bit fields; 32-bit path: nibbles
    const int T = (m + pc - 1) / pc + 31;
    // banded trace (AlignParams::pband): only the steps tlo(s) .. tlo(s)+tw-1 of every stripe were stored
    const int band_dmin = (!SMEM && P.pband) ? P.pband[2 * p] : 0, tw = (!SMEM && P.pband) ? P.pband[2 * p + 1] : 0;
    const int pitch = tw ? tw : T;
    const int last_s = (n - 1) >> 8, rr_last = P.plastrr[p];   // rows per lane in the (possibly compacted) last stripe
    // x / rr_last for x < 256 and rr_last in {2, 4, 6, 8} without an integer division in the walk: multiply-shift
    // ((x * 171) >> 10 == x / 6 for x < 512)
    const int rr_mul = rr_last == 6 ? 171 : 1, rr_sh = rr_last == 6 ? 10 : rr_last == 2 ? 1 : rr_last == 4 ? 2 : 3;
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
    // tie-break table (run-time part): which source code is state E, what the gap flags mean, the CIGAR letters
    const int src_e = P.tb_src_e;
    const int e_mask = split ? src_e : 1, f_mask = split ? 3 - src_e : 2;       // packed path: a state's flag bit is its class tag
    const int flag_opened = P.tb_open_tie;
    const uint32_t op_e = P.tb_swap_id ? ISOF_CIGAR_I : ISOF_CIGAR_D;           // E walks the reference
    const uint32_t op_f = P.tb_swap_id ? ISOF_CIGAR_D : ISOF_CIGAR_I;           // F walks the query
    if (i + 1 == n) PUSH(op_e, m - 1 - j);
    else PUSH(op_f, n - 1 - i);
    int where = 3;
    while (i >= 0 || j >= 0) {
        if (i < 0) { PUSH(op_e, j + 1); break; }
        if (j < 0) { PUSH(op_f, i + 1); break; }
        const int s = i >> 8;
        int l = (i >> 3) & 31, r = i & 7;
        if (s == last_s && rr_last != 8) { l = ((i & 255) * rr_mul) >> rr_sh; r = (i & 255) - l * rr_last; }
        const int jt = (pc == 1) ? j : (j >> 1), jc = (pc == 1) ? 0 : (j & 1);
        if (!SMEM) {   // the path mostly follows the diagonal: pull the line it will need 16 steps from now into L2
            const int ip = i - 16, jp = j - 16;
            if (ip >= 0 && jp >= 0) {
                const int sp_ = ip >> 8;
                const int lp = (sp_ == last_s && rr_last != 8) ? ((ip & 255) * rr_mul) >> rr_sh : (ip >> 3) & 31;
                const int jtp = (pc == 1) ? jp : (jp >> 1);
                const int tp_ = jtp + lp - (tw ? band_tlo(sp_, band_dmin) : 0);
                if ((unsigned)tp_ < (unsigned)pitch) {
                    const uint32_t *pf = &trace[(((size_t)sp_ * pitch + tp_) * 32 + lp) * pc];
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
                }
            }
        }
        const int tt = jt + l - (tw ? band_tlo(s, band_dmin) : 0);
        if (tw && (unsigned)tt >= (unsigned)tw) {     // the path left the band: this pair is re-aligned with the full trace
            P.nruns[p] = -1;
            return;
        }
        const uint32_t *wp = &trace[(((size_t)s * pitch + tt) * 32 + l) * pc + jc];
        ISOF_ASSERT(wp >= trace && wp < tail - TRACE_SLACK && wp < tail - nr, CHK_TRACE_LOAD);
        last_read = wp;
        const uint32_t word = SMEM ? *wp : __ldcg(wp);
        const int nib = split ? ((((word >> (16 + 2 * r)) & 3) << 2) | ((word >> (2 * r)) & 3)) : ((word >> (4 * r)) & 15);
        if (where == 3) {
            const int src = nib >> 2;
            if (src == 3) {
                PUSH(upc(s1[i]) == upc(s2[j]) ? ISOF_CIGAR_EQ : ISOF_CIGAR_X, 1);
                --i; --j;
            } else where = src;
        } else if (where == src_e) {      // E: consumes the reference -> 'D'
            PUSH(op_e, 1);
            --j;
            where = (((nib & e_mask) != 0) != (flag_opened != 0)) ? src_e : 3;
        } else {                          // F: consumes the query -> 'I'
            PUSH(op_f, 1);
            --i;
            where = (((nib & f_mask) != 0) != (flag_opened != 0)) ? 3 - src_e : 3;
        }
    }
    if (cur_len) { ++nr; ISOF_ASSERT(tail - nr >= trace, CHK_TAIL_OVERLAP); tail[-nr] = (cur_len << 4) | cur_op; }
#undef PUSH
    P.nruns[p] = nr;
}

// Warp-cooperative traceback of ONE pair whose trace lives in shared memory (fused small-call kernel).  The walk is a
// chain of dependent loads and address arithmetic -- ~300 cycles per step for a lone thread, 17 us for a 100 x 100 pair,
// as much as its DP.  But the trace holds the decision of EVERY cell, so the 32 lanes can look ahead: in state H lane k
// reads the cell k steps down the diagonal; the leading lanes whose source is "diagonal" are all on the path, and their
// = / X letters come from one ballot.  In a gap state lane k reads the cell k steps along the row (E) or column (F); the
// leading lanes whose flag says "extended" stay in the gap.  One iteration (~one lone-thread step) thus covers a whole
// diagonal stretch or a whole gap, up to 32 cells.  All lanes carry the same scalar state; same results as
// traceback_pair by construction (tests compare the fused call with the regular path and the oracle).
// SMEM = false: the trace lives in device memory (regular path, one warp per pair): loads bypass L1, the trace may be
// banded (AlignParams::pband) -- a cell outside the band ends a look-ahead, and the walk itself stepping on one reports
// nruns = -1 (second pass with the full trace).
template <bool SMEM>
__device__ __forceinline__ void traceback_pair_warp(const AlignParams &P, const int64_t p, const int lane)
{
    constexpr unsigned FULLW = 0xffffffffu;
    const int n = P.pn[p], m = P.pm[p];
    const uint8_t *s1 = P.cat + P.seq_off[P.pa[p]];
    const uint8_t *s2 = P.cat + P.seq_off[P.pb[p]];
    const uint32_t *trace = P.trace + (P.ptrace[p] - P.chunk_base);
    uint32_t *tail = run_tail(P, p);
    const int lay = P.playout[p];
    const int pc = lay & 15;
    const bool split = (lay & 16) != 0;
    const int T = (m + pc - 1) / pc + 31;
    const int band_dmin = (!SMEM && P.pband) ? P.pband[2 * p] : 0, tw = (!SMEM && P.pband) ? P.pband[2 * p + 1] : 0;
    const int pitch = tw ? tw : T;
    const int last_s = (n - 1) >> 8, rr_last = P.plastrr[p];
    const int rr_mul = rr_last == 6 ? 171 : 1, rr_sh = rr_last == 6 ? 10 : rr_last == 2 ? 1 : rr_last == 4 ? 2 : 3;
    int i = P.endq[p], j = P.endr[p];
    int nr = 0;
    uint32_t cur_op = 0xffu, cur_len = 0;
#define PUSHW(op, len)                                                                             \
    do {                                                                                           \
        if ((len) > 0) {                                                                           \
            if (cur_op == (uint32_t)(op)) cur_len += (uint32_t)(len);                              \
            else {                                                                                 \
                if (cur_len) { ++nr; ISOF_ASSERT(tail - nr >= trace, CHK_TAIL_OVERLAP); if (lane == 0) tail[-nr] = (cur_len << 4) | cur_op; } \
                cur_op = (op); cur_len = (len);                                                    \
            }                                                                                      \
        }                                                                                          \
    } while (0)
    // the 4-bit decision of cell (ii, jj): [source 2 bits | gap flags 2 bits]; -1 when a banded trace does not hold the cell
    auto nibble = [&](const int ii, const int jj) {
        const int s = ii >> 8;
        int l = (ii >> 3) & 31, r = ii & 7;
        if (s == last_s && rr_last != 8) { l = ((ii & 255) * rr_mul) >> rr_sh; r = (ii & 255) - l * rr_last; }
        const int jt = (pc == 1) ? jj : (jj >> 1), jc = (pc == 1) ? 0 : (jj & 1);
        const int tt = jt + l - (tw ? band_tlo(s, band_dmin) : 0);
        if (tw && (unsigned)tt >= (unsigned)tw) return -1;
        const uint32_t *wp = &trace[(((size_t)s * pitch + tt) * 32 + l) * pc + jc];
        ISOF_ASSERT(wp >= trace && wp < tail - TRACE_SLACK, CHK_TRACE_LOAD);
        const uint32_t word = SMEM ? *wp : __ldcg(wp);
        return split ? (int)((((word >> (16 + 2 * r)) & 3) << 2) | ((word >> (2 * r)) & 3)) : (int)((word >> (4 * r)) & 15);
    };
    const int src_e = P.tb_src_e;
    const int e_mask = split ? src_e : 1, f_mask = split ? 3 - src_e : 2;
    const int flag_opened = P.tb_open_tie;
    const uint32_t op_e = P.tb_swap_id ? ISOF_CIGAR_I : ISOF_CIGAR_D;
    const uint32_t op_f = P.tb_swap_id ? ISOF_CIGAR_D : ISOF_CIGAR_I;
    if (i + 1 == n) PUSHW(op_e, m - 1 - j);
    else PUSHW(op_f, n - 1 - i);
    int where = 3;
    while (i >= 0 || j >= 0) {
        if (i < 0) { PUSHW(op_e, j + 1); break; }
        if (j < 0) { PUSHW(op_f, i + 1); break; }
        if (where == 3) {
            const int ii = i - lane, jj = j - lane;
            const bool ok = ii >= 0 && jj >= 0;
            int nib = -1;
            bool eq = false;
            if (ok) { nib = nibble(ii, jj); eq = upc(s1[ii]) == upc(s2[jj]); }
            const unsigned diag = __ballot_sync(FULLW, nib >= 0 && (nib >> 2) == 3);
            const unsigned eqm = __ballot_sync(FULLW, eq);
            const int nd = (diag == FULLW) ? 32 : __ffs((int)~diag) - 1;           // leading diagonal cells
            if (nd == 0) {                                                         // same cell again, now in its gap state
                const int nib0 = __shfl_sync(FULLW, nib, 0);
                if (nib0 < 0) { if (lane == 0) P.nruns[p] = -1; return; }           // the path left the band
                where = nib0 >> 2;
                continue;
            }
            const unsigned mask = (nd == 32) ? FULLW : ((1u << nd) - 1u);
            const unsigned e = eqm & mask;
            int k = 0;
            while (k < nd) {                                                      // one pass per run of = / X
                const bool is_eq = (e >> k) & 1u;
                const unsigned other = (is_eq ? ~e : e) & mask & ~((1u << k) - 1u);
                const int k2 = other ? __ffs((int)other) - 1 : nd;
                PUSHW(is_eq ? ISOF_CIGAR_EQ : ISOF_CIGAR_X, k2 - k);
                k = k2;
            }
            i -= nd; j -= nd;
        } else {
            const bool in_e = (where == src_e);
            const int ii = in_e ? i : i - lane, jj = in_e ? j - lane : j;
            const bool ok = ii >= 0 && jj >= 0;
            bool stay = false;
            int nib = -1;
            if (ok) { nib = nibble(ii, jj); stay = ((nib & (in_e ? e_mask : f_mask)) != 0) != (flag_opened != 0); }
            if (__shfl_sync(FULLW, nib, 0) < 0) { if (lane == 0) P.nruns[p] = -1; return; }       // the path left the band
            const unsigned okm = __ballot_sync(FULLW, nib >= 0);
            const unsigned staym = __ballot_sync(FULLW, nib >= 0 && stay);
            const int ns = (staym == FULLW) ? 32 : __ffs((int)~staym) - 1;         // cells that keep the gap open
            // the first cell that closes the gap is consumed too (when it exists), and the walk returns to H
            const bool closes = ns < 32 && ((okm >> ns) & 1u);
            const int steps = ns + (closes ? 1 : 0);
            PUSHW(in_e ? op_e : op_f, steps);
            if (in_e) j -= steps; else i -= steps;
            if (closes) where = 3;
        }
    }
    if (cur_len) { ++nr; ISOF_ASSERT(tail - nr >= trace, CHK_TAIL_OVERLAP); if (lane == 0) tail[-nr] = (cur_len << 4) | cur_op; }
#undef PUSHW
    if (lane == 0) P.nruns[p] = nr;
}

// one warp per pair (regular path)
__global__ void __launch_bounds__(128) sg_traceback_warp_kernel(const AlignParams P, const int64_t start, const int64_t end)
{
    const int64_t p = start + (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (p >= end) return;
    traceback_pair_warp<false>(P, p, threadIdx.x & 31);
}

__global__ void sg_traceback_kernel(const AlignParams P, const int64_t start, const int64_t end)
{
    const int64_t p = start + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= end) return;
    traceback_pair<false>(P, p);
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
    long long nr = (tid < n_pairs) ? (long long)max(P.nruns[tid], 0) : 0, excl = 0, total = 0;
    Scan(scan_tmp).ExclusiveSum(nr, excl, total);
    s_off[tid] = excl;
    if (tid == 0) s_off[n_pairs] = total;
    __syncthreads();
    if (cigar_off && tid <= n_pairs) cigar_off[tid] = s_off[tid];
    if (tid == 0) *run_base = total;
    if (!cigar) return;
    for (int p = warp; p < n_pairs; p += 32) {
        const int n = max(P.nruns[p], 0);
        const long long dst = s_off[p];
        const uint32_t *src = run_tail(P, p) - n;
        ISOF_ASSERT(P.ptail || src >= P.trace + (P.ptrace[p] - P.chunk_base), CHK_CIGAR_COPY);
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
    const int nr = max(P.nruns[p], 0);
    if (lane == 0 && cigar_off) cigar_off[p] = dst;
    if (!cigar) return;
    const uint32_t *src = run_tail(P, p) - nr;
    ISOF_ASSERT(P.ptail || src >= P.trace + (P.ptrace[p] - P.chunk_base), CHK_CIGAR_COPY);
    for (int x = lane; x < nr; x += 32)
        if (dst + x < cigar_cap) cigar[dst + x] = src[x];
}

__global__ void advance_base_kernel(int64_t *run_base, const int64_t *local_off, const int32_t *nruns, int64_t start,
                                    int64_t end, int64_t *cigar_off, int64_t P)
{
    if (end > start) *run_base += local_off[end - 1 - start] + max(nruns[end - 1], 0);
    if (cigar_off && end == P) cigar_off[P] = *run_base;
}

__global__ void nruns_to_i64_kernel(const int32_t *nruns, int64_t start, int64_t count, int64_t *out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = max(nruns[start + i], 0);
}

// ---------------------------------------------------------------------------------------------
// fused epilogue: integer features of the reference's CIGAR decisions (one thread per pair)
// ---------------------------------------------------------------------------------------------
// One side of the reference's cigar_to_seq (consensus.py:35-47) as a character stream: the aligned string of one
// sequence under the run list, produced left to right (FWD) or right to left.  `own_gap` is the op that consumes only
// the OTHER sequence (this side shows '-').  The reference slices with Python semantics -- a slice past the end of
// the string is silently shortened -- so under a swapped I/D letter convention (tie-break table) the two aligned
// strings may differ in length and drift against each other; zip() then stops at the shorter one.  With the default
// letters both streams advance in lock step and the kernel uses the direct form below instead.
template <bool FWD>
struct AlnStream {
    const uint8_t *seq;
    const uint32_t *runs;
    int n, nr, other_only_op, x, y, idx;
    __device__ AlnStream(const uint8_t *seq_, int n_, const uint32_t *runs_, int nr_, int other_only_op_)
        : seq(seq_), runs(runs_), n(n_), nr(nr_), other_only_op(other_only_op_), x(FWD ? 0 : nr_ - 1), y(0), idx(0)
    {
        if (!FWD) {
            int tot = 0;
            for (int q = 0; q < nr; ++q) if ((int)(runs[q] & 15) != other_only_op) tot += (int)(runs[q] >> 4);
            idx = tot - 1;
        }
    }
    __device__ int next()          // next character of the aligned string, -1 at its end
    {
        while (FWD ? x < nr : x >= 0) {
            const uint32_t rr = runs[x];
            const int len = (int)(rr >> 4), op = (int)(rr & 15);
            if (y >= len) { x += FWD ? 1 : -1; y = 0; continue; }
            ++y;
            if (op == other_only_op) return '-';
            const int at = idx;
            idx += FWD ? 1 : -1;
            if (at < n) return seq[at];          // beyond the end: Python's slice yields nothing
        }
        return -1;
    }
};

__device__ __forceinline__ void features_pair(const AlignParams &P, const int64_t p, const int delta_len, int32_t *__restrict__ f)
{
    const int n = P.pn[p], m = P.pm[p];
    const uint8_t *s1 = P.cat + P.seq_off[P.pa[p]];
    const uint8_t *s2 = P.cat + P.seq_off[P.pb[p]];
    const int nr = P.nruns[p];
    if (nr < 0) { f[0] = -1; return; }                // banded trace, path left the band: filled in by the second pass
    const uint32_t *runs = run_tail(P, p) - nr;
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
    if (P.tb_swap_id) {
        // swapped letters: the reference's aligned strings are no longer column-synchronous (see AlnStream)
        {
            AlnStream<true> a1(s1, n, runs, nr, ISOF_CIGAR_D), a2(s2, m, runs, nr, ISOF_CIGAR_I);
            uint32_t bits = 0;
            for (int col = 0;; ++col) {
                const int c1 = a1.next(), c2 = a2.next();
                if (c1 < 0 || c2 < 0) break;
                bits = ((bits << 1) | (uint32_t)(c1 == c2)) & 0x3fffffffu;
                if (col >= WIN - 1 && __popc(bits) >= NEED) { start_match = col - (WIN - 1); break; }
            }
        }
        {
            AlnStream<false> a1(s1, n, runs, nr, ISOF_CIGAR_D), a2(s2, m, runs, nr, ISOF_CIGAR_I);
            uint32_t bits = 0;
            for (int col = 0;; ++col) {
                const int c1 = a1.next(), c2 = a2.next();
                if (c1 < 0 || c2 < 0) break;
                bits = ((bits << 1) | (uint32_t)(c1 == c2)) & 0x3fffffffu;
                if (col >= WIN - 1 && __popc(bits) >= NEED) { end_match = col - (WIN - 1); break; }
            }
        }
    } else {
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

__global__ void sg_features_kernel(const AlignParams P, const int64_t start, const int64_t end, int delta_len,
                                   int32_t *__restrict__ feat)
{
    const int64_t p = start + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= end) return;
    features_pair(P, p, delta_len, feat + p * ISOF_N_FEATURES);
}

// ---------------------------------------------------------------------------------------------
// Tiny calls (the per-pair seams: a memo miss of consensus.parasail_alignment, one align_to_merge): ONE launch does
// everything -- base codes, DP (32-bit wavefront path, one warp per pair), traceback, features -- with the trace in
// SHARED memory (the traceback's dependent loads then cost ~30 cycles instead of an L2 round trip each) and inputs /
// outputs in mapped pinned host memory (no memcpy calls: the kernel reads the sequences and writes the results over
// PCIe itself).  The regular path spends 0.25 ms on such a call in five launches, eight small copies and two stream
// synchronisations around a DP of 8 us.
// ---------------------------------------------------------------------------------------------
struct FusedArgs {
    int4 desc[64];                    // kernel parameter (no PCIe read): per pair {n, m, byte offset of s1, of s2 | run slot, bit 0 packed-eligible + bit 1 chase,
                                      //   1 + word offset of the pair's trace in `gtrace` (64 bit; 0 = trace in shared memory)}
    uint32_t *gtrace;                 // device memory: traces of the pairs too large for shared memory
    // chase mode (one block per stripe, see ChaseCtl): per (pair, stripe) completion flag and end-cell candidate
    unsigned *done; int2 *acc; int n_stripes;
    unsigned *ticket;                 // chase mode: blocks take their (pair, stripe) in the order they START
    const uint8_t *in;                // mapped host memory: the sequences, each at a 16-byte aligned offset
    int32_t *score, *endq, *endr, *nruns, *feat; uint32_t *runs;       // mapped host memory
    int2 *bnd; int64_t bnd_stride;
    int c_eo, c_ee, c_fo, c_fe, c_sm, c_sx;
    uint32_t h_eo, h_ee, h_fo, h_fe, h_sm, h_sx;
    int tb_src_e, tb_open_tie, tb_end_col_first, tb_swap_id, delta_len, tb_variant;
    long long *stamps;                // mapped host memory: SM clock at the phase boundaries of pair 0 (latency accounting)
};

__global__ void __launch_bounds__(32) sg_fused_small_kernel(const FusedArgs a)
{
    extern __shared__ uint4 fs_smem[];                          // [query profile 4 KB | trace words | cat | codes]
    __shared__ int64_t s_off[3], s_ptrace[2];
    __shared__ int32_t s_i32[8];                                 // pa pb pn pm score endq endr nruns
    __shared__ uint8_t s_u8[4];                                  // pwild layout lastrr
    __shared__ int2 s_ring[64];                                  // chase mode: settled batches of the boundary row
    const int lane = threadIdx.x;
    int p = blockIdx.y, g = blockIdx.x;
    if (a.n_stripes > 1) {
        // A stripe's block waits for the block of the stripe before it.  The hardware usually starts blocks in index order,
        // but nothing guarantees it, and a waiting block whose predecessor never becomes resident would hang the device:
        // (pair, stripe) is therefore assigned by a ticket taken at block start (as in decoupled look-back scans) -- the
        // block a waiter depends on has taken its ticket earlier, so it is running.
        unsigned t = 0;
        if (lane == 0) t = atomicAdd(a.ticket, 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        p = (int)(t / (unsigned)a.n_stripes); g = (int)(t % (unsigned)a.n_stripes);
    }
    const long long t_start = clock64();
    const int4 d0 = a.desc[2 * p], d1 = a.desc[2 * p + 1];       // kernel parameters: constant bank
    const int n = d0.x, m = d0.y;
    const int ns = (n + SROWS - 1) / SROWS;
    const bool chase = (d1.y & 2) != 0;                          // one block per stripe; otherwise block 0 does the pair
    if (g > 0 && (!chase || g >= ns)) return;
    const int64_t words = (int64_t)ns * (m + 64) * 32 + TRACE_SLACK;
    const int n16 = (n + 15) & ~15, m16 = (m + 15) & ~15;
    uint4 *prof = fs_smem;
    const long long goff = ((long long)(uint32_t)d1.z | ((long long)d1.w << 32)) - 1;       // < 0: shared memory
    uint32_t *smem_words = reinterpret_cast<uint32_t *>(fs_smem + 4 * 2 * 32);
    uint32_t *trace = goff < 0 ? smem_words : a.gtrace + goff;
    uint8_t *s_cat = reinterpret_cast<uint8_t *>(goff < 0 ? smem_words + words : smem_words);   // s1 at 0, s2 at n16 (16-byte aligned)
    uint8_t *s_codes = s_cat + n16 + m16;
    int wild = 0;
    // 16 bytes per lane and round: a 100 x 100 pair arrives in one PCIe round trip instead of seven
    const int c1 = n16 >> 4, c2 = m16 >> 4;
    // (four loads in flight per lane: every round is a PCIe round trip of ~3 us, an 8 KB pair would take 16 of them)
    for (int cb = lane; cb < c1 + c2; cb += 128) {
      uint4 vq[4];
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
          const int c = cb + 32 * q4;
          if (c < c1 + c2) vq[q4] = *reinterpret_cast<const uint4 *>(a.in + (c < c1 ? d0.z + c * 16 : d0.w + (c - c1) * 16));
      }
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const int c = cb + 32 * q4;
        if (c >= c1 + c2) break;
        const bool first = c < c1;
        const int byte0 = first ? c * 16 : (c - c1) * 16, len = first ? n : m;
        const uint4 v = vq[q4];
        const int so = first ? byte0 : n16 + byte0;
        *reinterpret_cast<uint4 *>(s_cat + so) = v;
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t cw[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t cc = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int code = base_code((uint8_t)(w[q] >> (8 * b)));
                cc |= (uint32_t)code << (8 * b);
                wild |= (code == 4) && (byte0 + 4 * q + b < len);
            }
            cw[q] = cc;
        }
        *reinterpret_cast<uint4 *>(s_codes + so) = make_uint4(cw[0], cw[1], cw[2], cw[3]);
      }
    }
    wild = __any_sync(0xffffffffu, wild);
    const long long t_loaded = clock64();
    if (lane == 0) {
        s_off[0] = 0; s_off[1] = n16; s_off[2] = n16 + m16;
        s_ptrace[0] = 0; s_ptrace[1] = words;
        s_i32[0] = 0; s_i32[1] = 1; s_i32[2] = n; s_i32[3] = m;
        s_u8[0] = (uint8_t)wild;
    }
    __syncwarp();
    AlignParams P;
    P.cat = s_cat; P.codes = s_codes; P.seq_off = s_off; P.pa = &s_i32[0]; P.pb = &s_i32[1]; P.pn = &s_i32[2]; P.pm = &s_i32[3];
    P.pwild = &s_u8[0]; P.ptrace = s_ptrace; P.chunk_base = 0; P.trace = trace;
    P.bnd = a.bnd + (size_t)p * a.n_stripes * a.bnd_stride; P.bnd_stride = a.bnd_stride;
    P.score = &s_i32[4]; P.endq = &s_i32[5]; P.endr = &s_i32[6]; P.nruns = &s_i32[7];
    P.c_eo = a.c_eo; P.c_ee = a.c_ee; P.c_fo = a.c_fo; P.c_fe = a.c_fe; P.c_sm = a.c_sm; P.c_sx = a.c_sx;
    P.h_eo = a.h_eo; P.h_ee = a.h_ee; P.h_fo = a.h_fo; P.h_fe = a.h_fe; P.h_sm = a.h_sm; P.h_sx = a.h_sx;
    P.playout = &s_u8[1]; P.plastrr = &s_u8[2]; P.chunk_words = 0; P.order = nullptr; P.n_elig = nullptr; P.n_rb = nullptr;
    P.tb_src_e = a.tb_src_e; P.tb_open_tie = a.tb_open_tie; P.tb_end_col_first = a.tb_end_col_first; P.tb_swap_id = a.tb_swap_id;
    P.ptail = nullptr;
    P.pband = nullptr;
    // A lone pair inside the 16-bit range runs the PACKED path with itself in both halves (identical words go to the same
    // trace addresses): that path compacts the last stripe to 2 / 4 / 6 rows per lane, so a 100-row query keeps 25 lanes
    // busy with a 4-row dependency chain per step instead of 13 lanes with an 8-row one -- half the latency per step.
    // (default tie-break table only: the other tables take the 32-bit path, whose constants are run-time values)
    if (chase) {
        ChaseCtl ch;
        const size_t slot = (size_t)p * a.n_stripes + g;
        ch.s = g; ch.ring = s_ring; ch.spins = 0; ch.spin_cycles = 0;
        ch.bnd_out = a.bnd + slot * a.bnd_stride; ch.bnd_in = ch.bnd_out - a.bnd_stride;      // bnd_in unused by stripe 0
        long long g0 = 0, g1 = 0;
        if (p == 0 && lane == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
        dp_pair16<0, false, true>(P, 0, 0, lane, nullptr, prof, &ch);
        if (p == 0 && lane == 0 && g < 16) {             // latency accounting: DP begin / end of every stripe of pair 0 (ns)
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
            a.stamps[8 + 2 * g] = g0; a.stamps[9 + 2 * g] = g1;
            a.stamps[40 + g] = (ch.spins << 40) | ch.spin_cycles;
        }
        // the last column's best cell travels down the chain of stripes (ties: the smaller row, i.e. the earlier stripe)
        int bcol = ch.bcol, bcol_i = ch.bcol_i;
        if (lane == 0 && g > 0) {
            while (ld_acquire_u32(a.done + slot - 1) == 0u) { }
            const int2 prev = __ldcg(a.acc + slot - 1);
            if (prev.x >= bcol) { bcol = prev.x; bcol_i = prev.y; }
        }
        __syncwarp();                                    // every lane's trace stores are ordered before lane 0's release
        if (g + 1 < ns) {
            if (lane == 0) { a.acc[slot] = make_int2(bcol, bcol_i); __threadfence(); st_release_u32(a.done + slot, 1u); }
            return;
        }
        if (lane == 0) {                                 // last stripe: end cell, then the traceback below
            int sc, eq, er;
            pick_end_cell(P.tb_end_col_first, ch.brow, ch.brow_j, bcol, bcol_i, n, m, sc, eq, er);
            s_i32[4] = sc >> 2; s_i32[5] = eq; s_i32[6] = er;
        }
    }
    else if (!wild && (d1.y & 1) && a.tb_variant == 0) dp_pair16<0, false>(P, 0, 0, lane, P.bnd, prof);
    else if (wild) dp_pair<true>(P, 0, lane, P.bnd, prof);
    else dp_pair<false>(P, 0, lane, P.bnd, prof);
    __syncwarp();
    const long long t_dp = clock64();
    long long t_tb = 0;
    traceback_pair_warp<true>(P, 0, lane);
    __syncwarp();
    if (lane == 0) {
        t_tb = clock64();
        if (a.feat) features_pair(P, 0, a.delta_len, a.feat + (size_t)p * ISOF_N_FEATURES);
        a.score[p] = s_i32[4]; a.endq[p] = s_i32[5]; a.endr[p] = s_i32[6]; a.nruns[p] = s_i32[7];
    }
    __syncwarp();
    const int nr = s_i32[7];
    const uint32_t *src = trace + words - nr;
    uint32_t *dst = a.runs + d1.x;
    for (int x = lane; x < nr; x += 32) dst[x] = src[x];
    if (p == 0 && lane == 0) {                          // (chase mode: the last stripe's block reports)
        a.stamps[0] = t_loaded - t_start; a.stamps[1] = t_dp - t_loaded; a.stamps[2] = t_tb - t_dp;
        a.stamps[3] = clock64() - t_tb;
    }
}

}  // namespace

// Host buffers in, host buffers out, one launch.  Returns 1 when the call is not eligible (caller takes the regular path).
int isof_sg_fused_small(isof_ctx *ctx, const uint8_t *seq_cat, const int64_t *seq_off, int32_t n_seqs, const int32_t *pair_a,
                        const int32_t *pair_b, int64_t P, int32_t match, int32_t mismatch, int32_t gap_open, int32_t gap_ext,
                        int32_t delta_len, int32_t *score, int32_t *end_q, int32_t *end_r, uint32_t *cigar_runs,
                        int64_t cigar_cap, int64_t *cigar_off, int32_t *features)
{
    constexpr int64_t MAX_P = 32, MAX_SMEM = 200 * 1024, IO_BYTES = 512 * 1024;
    if (P < 1 || P > MAX_P || ctx->profiling) return 1;
    if (match - mismatch > 256 || match - mismatch < 0 || gap_open < 0 || gap_ext < 0 || std::abs(match) > 1024 ||
        std::abs(mismatch) > 1024 || gap_open > 4096 || gap_ext > 4096)
        return 1;                                          // the regular path reports the range error
    int h_max_min_len = 0, h_max_max_len = 0;              // packed 16-bit eligibility, as in isof_sg_align_core
    if (!ctx->disable_packed && match > 0 && mismatch <= match && 4 * (match - mismatch) <= (1 << CS16)) {
        h_max_min_len = (32767 - 16) / (4 * match);
        const int room = 31000 / 4 - 2 * gap_open - std::abs(mismatch) - std::abs(match);
        h_max_max_len = room <= 0 ? 0 : (gap_ext > 0 ? room / gap_ext : INT_MAX);
    }
    // ---- plan on the host: lengths, validity, shared memory, run slots, staging offsets
    int64_t smem_max = 0, run_words = 0, in_bytes = 0, gwords = 0;
    int max_m = 0, n_stripes = 1;
    const bool default_table = !ctx->tb_h_e_before_f && !ctx->tb_gap_open_on_tie;
    auto plain_acgt = [](const uint8_t *x, int64_t len) {
        for (int64_t q = 0; q < len; ++q) {
            const uint8_t c = x[q] & 0xdf;                 // upper case
            if (c != 'A' && c != 'C' && c != 'G' && c != 'T') return false;
        }
        return true;
    };
    int64_t roff[MAX_P + 1];
    int32_t desc[MAX_P][8];
    bool chase_p[MAX_P];
    for (int64_t p = 0; p < P; ++p) {
        const int a = pair_a[p], b = pair_b[p];
        if (a < 0 || a >= n_seqs || b < 0 || b >= n_seqs) return 1;       // regular path: ISOF_ERR_ARG with its message
        const int64_t n = seq_off[a + 1] - seq_off[a], m = seq_off[b + 1] - seq_off[b];
        if (n <= 0 || m <= 0 || n > 4096 || m > 4096) return 1;
        const int64_t ns = (n + SROWS - 1) / SROWS;
        // several stripes: one block per stripe, chasing its predecessor (packed path, default table, no wildcard)
        chase_p[p] = ns >= 2 && std::min(n, m) <= h_max_min_len && std::max(n, m) <= h_max_max_len && default_table &&
                     plain_acgt(seq_cat + seq_off[a], n) && plain_acgt(seq_cat + seq_off[b], m);
        if (chase_p[p]) n_stripes = std::max(n_stripes, (int)ns);
    }
    for (int64_t p = 0; p < P; ++p) {
        const int a = pair_a[p], b = pair_b[p];
        const int64_t n = seq_off[a + 1] - seq_off[a], m = seq_off[b + 1] - seq_off[b];
        const int64_t ns = (n + SROWS - 1) / SROWS;
        const int64_t words = ns * (m + 64) * 32 + TRACE_SLACK;
        const int64_t n16 = (n + 15) & ~15LL, m16 = (m + 15) & ~15LL;
        int64_t smem = 4096 + words * 4 + 2 * (n16 + m16) + 64, goff1 = 0;
        const bool packed_ok = std::min(n, m) <= h_max_min_len && std::max(n, m) <= h_max_max_len;
        const bool chase = chase_p[p];
        // trace in device memory (the lanes' look-ahead hides its latency): too large for shared memory, or a chase call --
        // the stripes' blocks sit on different SMs, and all blocks of the launch share one shared-memory size
        if (smem > MAX_SMEM || n_stripes > 1) {
            smem = 4096 + 2 * (n16 + m16) + 64;
            goff1 = gwords + 1;
            gwords += (words + 31) & ~31LL;
        }
        smem_max = std::max(smem_max, smem);
        roff[p] = run_words;
        desc[p][0] = (int32_t)n; desc[p][1] = (int32_t)m; desc[p][2] = (int32_t)in_bytes; desc[p][3] = (int32_t)(in_bytes + n16);
        desc[p][4] = (int32_t)run_words;
        desc[p][5] = (packed_ok ? 1 : 0) | (chase ? 2 : 0);
        desc[p][6] = (int32_t)(uint32_t)(goff1 & 0xffffffffLL); desc[p][7] = (int32_t)(goff1 >> 32);
        in_bytes += n16 + m16;
        run_words += n + m + 2;
        max_m = std::max(max_m, (int)m);
    }
    roff[P] = run_words;
    const int64_t out_bytes = 4 * (4 + ISOF_N_FEATURES) * P + 4 * run_words + 64 + 512;      // + the phase clocks at the tail
    if (in_bytes > IO_BYTES || out_bytes > IO_BYTES) return 1;
    if (!ctx->fused_io) {
        CUDA_TRY(ctx, cudaHostAlloc(&ctx->fused_io, 2 * IO_BYTES, cudaHostAllocMapped));
        CUDA_TRY(ctx, cudaHostGetDevicePointer(&ctx->fused_io_dev, ctx->fused_io, 0));
        CUDA_TRY(ctx, cudaFuncSetAttribute(sg_fused_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MAX_SMEM));
    }
    // ---- stage the inputs (mapped pinned memory: the kernel reads them over PCIe, no memcpy call)
    uint8_t *in = (uint8_t *)ctx->fused_io, *out = in + IO_BYTES;
    uint8_t *d_in = (uint8_t *)ctx->fused_io_dev, *d_out = d_in + IO_BYTES;
    for (int64_t p = 0; p < P; ++p) {
        const int64_t n = desc[p][0], m = desc[p][1];
        memcpy(in + desc[p][2], seq_cat + seq_off[pair_a[p]], (size_t)n);
        memcpy(in + desc[p][3], seq_cat + seq_off[pair_b[p]], (size_t)m);
    }
    int32_t *o_score = (int32_t *)out, *o_endq = o_score + P, *o_endr = o_endq + P, *o_nruns = o_endr + P;
    int32_t *o_feat = o_nruns + P;
    uint32_t *o_runs = (uint32_t *)(o_feat + ISOF_N_FEATURES * P);
    const int64_t bnd_stride = ((int64_t)max_m + 128) & ~63LL;       // touched by multi-stripe pairs only; grow-only reserve
    TRY(dev_reserve(ctx, ctx->b_bnd, sizeof(int2) * (size_t)bnd_stride * (size_t)P * (size_t)n_stripes));
    FusedArgs a;
    const ptrdiff_t dout = d_out - out;
    memcpy(a.desc, desc, 32 * (size_t)P);
    a.in = d_in;
    a.score = (int32_t *)((uint8_t *)o_score + dout); a.endq = (int32_t *)((uint8_t *)o_endq + dout);
    a.endr = (int32_t *)((uint8_t *)o_endr + dout); a.nruns = (int32_t *)((uint8_t *)o_nruns + dout);
    a.feat = features ? (int32_t *)((uint8_t *)o_feat + dout) : nullptr;
    a.runs = (uint32_t *)((uint8_t *)o_runs + dout);
    a.bnd = (int2 *)ctx->b_bnd.p; a.bnd_stride = bnd_stride;
    a.n_stripes = n_stripes; a.done = nullptr; a.acc = nullptr; a.ticket = nullptr;
    if (n_stripes > 1) {
        const size_t slots = (size_t)P * (size_t)n_stripes;
        TRY(dev_reserve(ctx, ctx->b_chase, 16 * slots + 16));
        CUDA_TRY(ctx, cudaMemsetAsync(ctx->b_chase.p, 0, 4 * slots + 16, ctx->stream));  // done flags + the ticket counter
        a.done = (unsigned *)ctx->b_chase.p; a.ticket = a.done + slots;
        a.acc = (int2 *)((uint8_t *)ctx->b_chase.p + 8 * slots + 16);
        // boundary lines start "empty" (CHASE_EMPTY in every word)
        CUDA_TRY(ctx, cudaMemsetAsync(ctx->b_bnd.p, 0x80, sizeof(int2) * (size_t)bnd_stride * slots, ctx->stream));
    }
    a.gtrace = nullptr;
    if (gwords) {
        TRY(dev_reserve(ctx, ctx->b_trace, sizeof(uint32_t) * (size_t)gwords));
        a.gtrace = (uint32_t *)ctx->b_trace.p;
    }
    a.stamps = (long long *)(d_out + IO_BYTES - 512);
    const bool e_first = ctx->tb_h_e_before_f != 0, open_tie = ctx->tb_gap_open_on_tie != 0;
    const int S = 1 << TAG_BITS, class_e = e_first ? 8 : 4, class_f = e_first ? 4 : 8;
    a.c_eo = -gap_open * S + class_e + (open_tie ? 1 : 0);
    a.c_ee = -gap_ext * S + class_e + (open_tie ? 0 : 1);
    a.c_fo = -gap_open * S + class_f + (open_tie ? 2 : 0);
    a.c_fe = -gap_ext * S + class_f + (open_tie ? 0 : 2);
    a.c_sm = match * S + 12 - (3 << CODE_SHIFT);
    a.c_sx = mismatch * S + 12;
    auto pk = [](int v) { return ((uint32_t)(v & 0xffff)) | ((uint32_t)(v & 0xffff) << 16); };
    a.h_eo = pk(-4 * gap_open); a.h_ee = pk(-4 * gap_ext); a.h_fo = a.h_eo; a.h_fe = a.h_ee;      // default table only (tb_variant 0)
    a.h_sm = pk(4 * match + 3 - (3 << CS16)); a.h_sx = pk(4 * mismatch + 3);
    a.tb_variant = (e_first ? 1 : 0) | (open_tie ? 2 : 0);
    a.tb_src_e = e_first ? 2 : 1; a.tb_open_tie = open_tie ? 1 : 0; a.tb_end_col_first = ctx->tb_end_col_first ? 1 : 0;
    a.tb_swap_id = ctx->tb_swap_id ? 1 : 0; a.delta_len = delta_len;
    ctx->prof.launches++; ctx->prof.dp_launches++;
    sg_fused_small_kernel<<<dim3((unsigned)n_stripes, (unsigned)P), 32, (size_t)smem_max, ctx->stream>>>(a);
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    // ---- results
    int64_t total = 0;
    for (int64_t p = 0; p < P; ++p) {
        score[p] = o_score[p];
        if (end_q) end_q[p] = o_endq[p];
        if (end_r) end_r[p] = o_endr[p];
        if (cigar_off) cigar_off[p] = total;
        const int nr = o_nruns[p];
        if (cigar_runs && total + nr <= cigar_cap) memcpy(cigar_runs + total, o_runs + roff[p], 4 * (size_t)nr);
        total += nr;
        ctx->prof.dp_cells += (int64_t)desc[p][0] * desc[p][1];
    }
    if (cigar_off) cigar_off[P] = total;
    if (features) memcpy(features, o_feat, 4 * (size_t)ISOF_N_FEATURES * (size_t)P);
    if (cigar_runs && cigar_cap > 0 && total > cigar_cap)
        return isof_fail(ctx, ISOF_ERR_CAPACITY, "cigar_runs holds %lld runs, %lld needed", (long long)cigar_cap, (long long)total);
    return ISOF_OK;
}

unsigned isof_viol_align()
{
    return isof_read_violations_tu() | isof_viol_dp_tb1() | isof_viol_dp_tb2() | isof_viol_dp_tb3() | isof_viol_short() |
           isof_viol_dp_rb();
}

void isof_launch_dp_tb0(int blocks, cudaStream_t st, const AlignParams &A, int64_t start, int64_t end, unsigned long long *counter)
{
    launch_sg_dp<0>(blocks, st, A, start, end, counter);
}

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
    // ---- re-based 16-bit lanes for pairs beyond that range (dp_pair16<., RB>): the packed values follow a running base, so
    // only the SPREAD of the scores inside the wavefront must fit: neighbouring cells differ by at most match + open (or
    // open + ext), the wavefront spans 128 rows either side of the reference lane, 62 columns of lane skew and 64 columns
    // of drift between two re-centrings (+32 slack)
    int rb_ok = 0;
    if (h_max_min_len > 0 && 4 * std::max(gap_open + gap_ext, match + gap_open) * (128 + 62 + 64 + 32) <= 24000) rb_ok = 1;
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
    TRY(dev_reserve(ctx, ctx->b_misc, 128));
    // banded trace for the re-based class: feature-only calls (the CIGAR stream of a call that returns runs is laid out
    // is then spliced together from the two passes, which needs the offsets)
    const int band_w0 = ((!d_cigar || d_cigar_off) && rb_ok && d_endq && d_endr) ? ctx->band_w0 : 0;
    int32_t *pband = nullptr;
    if (band_w0 > 0) {
        TRY(dev_reserve(ctx, ctx->b_band, sizeof(int32_t) * 2 * (size_t)P));
        pband = (int32_t *)ctx->b_band.p;
    }
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
        TRY(dev_reserve(ctx, ctx->b_counters, sizeof(int64_t) * 3 * 5));
        unsigned long long *cnt0 = (unsigned long long *)ctx->b_counters.p + 3;          // counters[0] of a 1-chunk layout
        LaunchTimer t(ctx, SLOT_OTHER);
        small_plan_kernel<<<1, 1024, 0, st>>>(d_cat, d_off, n_seqs, d_pa, d_pb, (int)P, codes, flags, pn, pm, pwild, ptrace,
                                              misc, h_max_min_len, h_max_max_len, (int32_t *)ctx->b_order.p, cnt0 + 3, cnt0,
                                              SHORT_MAX, rb_ok, cnt0 + 6, cnt0 + 9, band_w0, pband);
    } else {
        CUDA_TRY(ctx, cudaMemsetAsync(misc, 0, 128, st));
        CUDA_TRY(ctx, cudaMemsetAsync(pwild, 0, (size_t)P, st));
        {
            LaunchTimer t(ctx, SLOT_OTHER);
            int blocks = std::min((n_seqs * 32 + 255) / 256, ctx->sm_count * 8);
            encode_kernel<<<std::max(blocks, 1), 256, 0, st>>>(d_cat, d_off, n_seqs, codes, flags);
        }
        {
            LaunchTimer t(ctx, SLOT_OTHER);
            plan_kernel<<<(unsigned)((P + 1 + 255) / 256), 256, 0, st>>>(d_off, n_seqs, flags, d_pa, d_pb, P, pn, pm, pwild,
                                                                          words, misc, h_max_min_len, h_max_max_len, SHORT_MAX, rb_ok, band_w0,
                                                                          pband);
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
    CUDA_TRY(ctx, cudaMemcpyAsync(pin + 8, misc + 8, 24, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    if (pin[0] & 1) return isof_fail(ctx, ISOF_ERR_ARG, "pair index outside the sequence pool");
    if (pin[0] & 2) return isof_fail(ctx, ISOF_ERR_ARG, "empty sequence in a pair (parasail rejects it too)");
    const int max_m = (int)pin[1], max_n = (int)pin[8];
    // Short-pair regime: a thread per task has ~10x the per-task latency of the 32-lane wavefront (a 100 x 100 group takes
    // 0.3 ms whatever the pair count), so it only pays once the call holds enough groups to fill the machine; measured
    // crossover (profiles/r02_short_crossover.jsonl): P ~ 64 x SMs x (L / 60)^2.  short_mode 2 forces it (tests).
    const double len_ratio = (double)std::max((int)pin[1], (int)pin[8]) / 60.0;
    const bool enough = ctx->short_mode == 2 || (double)P >= 64.0 * ctx->sm_count * len_ratio * len_ratio;
    const bool all_short = pin[9] == 0 && !ctx->disable_packed && ctx->short_mode != 0 && h_max_min_len > 0 && enough;
    const bool any_wild = pin[10] != 0;
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
    AlignParams A;
    A.cat = d_cat; A.codes = codes; A.seq_off = d_off; A.pa = d_pa; A.pb = d_pb; A.pn = pn; A.pm = pm; A.pwild = pwild;
    A.ptrace = ptrace; A.trace = nullptr; A.bnd = nullptr; A.bnd_stride = 0;
    A.score = d_score; A.endq = d_endq; A.endr = d_endr; A.nruns = nruns;
    // ---- tagged constants under the tie-break table (DESIGN.md "parasail tie-break table").  The class tag of a gap
    // state orders the H tie (larger wins; the diagonal is the largest), the flag inside the class orders open vs
    // extend: the candidate that must win a tie carries the larger tag.
    const bool e_first = ctx->tb_h_e_before_f != 0, open_tie = ctx->tb_gap_open_on_tie != 0;
    const int S = 1 << TAG_BITS;
    {
        const int class_e = e_first ? 8 : 4, class_f = e_first ? 4 : 8;       // 32-bit path: flag of E in bit 0, of F in bit 1
        A.c_eo = -gap_open * S + class_e + (open_tie ? 1 : 0);
        A.c_ee = -gap_ext * S + class_e + (open_tie ? 0 : 1);
        A.c_fo = -gap_open * S + class_f + (open_tie ? 2 : 0);
        A.c_fe = -gap_ext * S + class_f + (open_tie ? 0 : 2);
    }
    A.c_sm = match * S + 12 - (3 << CODE_SHIFT);
    A.c_sx = mismatch * S + 12;
    auto pk = [](int v) { return ((uint32_t)(v & 0xffff)) | ((uint32_t)(v & 0xffff) << 16); };
    {
        // packed path: the stored E / F carry their class tag (TieBreakK<TB>), which doubles as the flag bit.  Default:
        // opening candidates have tag 00 and extension keeps the class tag.  gap_open_on_tie: the opening candidate
        // gets the class tag and the extension addend removes it.
        const int tag_e = e_first ? 2 : 1, tag_f = e_first ? 1 : 2;
        A.h_eo = pk(-4 * gap_open + (open_tie ? tag_e : 0));
        A.h_ee = pk(-4 * gap_ext - (open_tie ? tag_e : 0));
        A.h_fo = pk(-4 * gap_open + (open_tie ? tag_f : 0));
        A.h_fe = pk(-4 * gap_ext - (open_tie ? tag_f : 0));
    }
    A.h_sm = pk(4 * match + 3 - (3 << CS16));
    A.h_sx = pk(4 * mismatch + 3);
    A.tb_src_e = e_first ? 2 : 1;
    A.tb_open_tie = open_tie ? 1 : 0;
    A.tb_end_col_first = ctx->tb_end_col_first ? 1 : 0;
    A.tb_swap_id = ctx->tb_swap_id ? 1 : 0;
    const int tb_variant = (e_first ? 1 : 0) | (open_tie ? 2 : 0);
    A.ptail = nullptr;
    A.pband = pband;
    A.chunk_base = 0;
    int64_t *run_base = (int64_t *)(misc + 4);
    CUDA_TRY(ctx, cudaMemsetAsync(run_base, 0, 8, st));
    // ---- short-pair regime: every pair of the call is short (bubble popping, SimplifyGraph.py:630-664) -> one thread
    // per task, no wavefront (sg_short.cu); falls through to the long path when the trace would not fit one buffer
    if (all_short) {
        const size_t fr = ctx->trace_free_at_first_use;
        const size_t cap_bytes = ctx->trace_budget ? ctx->trace_budget : (size_t)(0.40 * (double)fr);
        const int rc = isof_sg_short_path(ctx, A, P, max_n, max_m, any_wild, tb_variant, (int64_t)(cap_bytes / 4));
        if (rc == ISOF_OK) {
            if (d_features) {
                LaunchTimer t(ctx, SLOT_TB);
                sg_features_kernel<<<(unsigned)((P + 63) / 64), 64, 0, st>>>(A, 0, P, delta_len, d_features);
            }
            if (small) {
                LaunchTimer t(ctx, SLOT_TB);
                small_finish_kernel<<<1, 1024, 0, st>>>(A, (int)P, (long long *)run_base, d_cigar, (long long)cigar_cap,
                                                        (long long *)d_cigar_off);
            } else {
                {
                    LaunchTimer t(ctx, SLOT_TB);
                    nruns_to_i64_kernel<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(nruns, 0, P, nruns64);
                }
                size_t tb2 = 0;
                cub::DeviceScan::ExclusiveSum(nullptr, tb2, nruns64, runoff, P, st);
                TRY(dev_reserve(ctx, ctx->b_scan_tmp, tb2));
                {
                    LaunchTimer t(ctx, SLOT_TB);
                    CUDA_TRY(ctx, cub::DeviceScan::ExclusiveSum(ctx->b_scan_tmp.p, tb2, nruns64, runoff, P, st));
                }
                {
                    LaunchTimer t(ctx, SLOT_TB);
                    cigar_copy_kernel<<<(unsigned)((P * 32 + 255) / 256), 256, 0, st>>>(A, 0, P, runoff, run_base, d_cigar, cigar_cap,
                                                                                       d_cigar_off);
                }
                {
                    LaunchTimer t(ctx, SLOT_TB);
                    advance_base_kernel<<<1, 1, 0, st>>>(run_base, runoff, nruns, 0, P, d_cigar_off, P);
                }
            }
            CUDA_TRY(ctx, cudaGetLastError());
            int64_t *pin64s = (int64_t *)ctx->pinned;
            CUDA_TRY(ctx, cudaMemcpyAsync(pin64s, run_base, 8, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(ctx, cudaStreamSynchronize(st));
            if (h_total_runs) *h_total_runs = pin64s[0];
            ctx->prof.dp_cells_packed += (int64_t)pin[2];
            if (d_cigar && pin64s[0] > cigar_cap)
                return isof_fail(ctx, ISOF_ERR_CAPACITY, "cigar_runs holds %lld runs, %lld needed", (long long)cigar_cap,
                                 (long long)pin64s[0]);
            return ISOF_OK;
        }
        if (rc != ISOF_ERR_NOMEM) return rc;
    }
    size_t budget_bytes = ctx->trace_budget;
    bool memory_bound = false;            // the trace of the pairs that fill the machine does not fit: deeper pipeline
    if (budget_bytes == 0) {
        const double fr = (double)ctx->trace_free_at_first_use;
        const double want = 4.0 * (double)total_words / (double)P * (double)(ctx->sm_count * DP_CTAS_PER_SM * DP_WARPS) * 1.25 *
                            (double)isof_ctx::NLANES;
        // long pairs (10 kb reads: 36 MB of trace each) are occupancy-bound by trace memory: 2368 resident warps x 2
        // alignments of the 16-bit lanes would need 175 GB; let such batches take up to 80 % of the free memory
        budget_bytes = (size_t)std::min(0.80 * fr, std::max(0.40 * fr, want));
        memory_bound = want > 0.80 * fr;
        if (budget_bytes < (64u << 20)) budget_bytes = 64u << 20;
    }
    int64_t budget_words = (int64_t)(budget_bytes / 4);
    // pipelined mode: NL smaller buffers, chunk c runs on stream c % NL.  Memory-bound batches use twice as many, half as
    // large: chunks then finish at staggered times, so buffers turn over continuously instead of three at once, and the
    // latency-bound traceback of one chunk (few, very long walks) overlaps the DP of five others
    const int NL = memory_bound ? isof_ctx::MAX_LANES : isof_ctx::NLANES;
    const bool want_overlap = ctx->overlap && total_words > budget_words / NL;
    if (want_overlap) budget_words /= NL;
    // a chunk holds the pairs that START inside its slice of the trace offsets, so it can exceed the slice by one pair
    const int64_t pair_max_words = (int64_t)((max_n + SROWS - 1) / SROWS) * (max_m + 64) * 32 + TRACE_SLACK;
    if (want_overlap && ctx->trace_budget == 0) {
        // lanes left by an earlier call (e.g. the first pass of a banded call, whose chunks were a sliver smaller): plan
        // the chunks to fit them rather than free and re-allocate tens of GB while the other lanes fill the memory
        size_t lane_cap = ctx->b_trace.cap;
        for (int w = 1; w < NL; ++w) lane_cap = std::min(lane_cap, ctx->b_trace_x[w - 1].cap);
        const int64_t fit = (int64_t)(lane_cap / 4) - pair_max_words;
        if (fit >= budget_words / 2 && fit < budget_words) budget_words = fit;
    }
    bool single = false, pipelined = false, fused_small = false;
    int n_chunks = 1;
    int64_t *chunk_start_d = nullptr;
    unsigned long long *counters = nullptr, *elig_counts = nullptr, *rb_counts = nullptr, *rb_tickets = nullptr;
    // a pair of the re-based class exists only if some pair exceeds the plain 16-bit range
    const bool may_rb = rb_ok && (std::max(max_n, max_m) > h_max_max_len || std::min(max_n, max_m) > h_max_min_len);
    std::vector<int64_t> chunk_start, start_off;
    for (int attempt = 0;; ++attempt) {
        if (budget_words > total_words) budget_words = total_words;
        if (budget_words < 1) budget_words = 1;
        single = total_words <= budget_words;              // everything fits one chunk: no device round trips
        n_chunks = single ? 1 : (int)(total_words / budget_words) + 1;
        TRY(dev_reserve(ctx, ctx->b_counters, sizeof(int64_t) * (size_t)(n_chunks + 2) * 5));
        chunk_start_d = (int64_t *)ctx->b_counters.p;
        counters = (unsigned long long *)(chunk_start_d + n_chunks + 2);       // task tickets per chunk
        elig_counts = counters + n_chunks + 2;                                  // eligible pairs per chunk
        rb_counts = elig_counts + n_chunks + 2;                                 // re-based class per chunk
        rb_tickets = rb_counts + n_chunks + 2;                                  // its task tickets
        fused_small = small && single;                     // tickets, class counts and task order came from small_plan_kernel
        if (!fused_small) CUDA_TRY(ctx, cudaMemsetAsync(counters, 0, sizeof(int64_t) * (size_t)(n_chunks + 2) * 4, st));
        chunk_start.assign((size_t)n_chunks + 1, 0);
        start_off.assign((size_t)n_chunks + 1, 0);
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
        pipelined = want_overlap && n_chunks > 1;
        bool fits = true;
        for (int w = 0; w < (pipelined ? NL : 1) && fits; ++w) {
            DevBuf *tb = w ? &ctx->b_trace_x[w - 1] : &ctx->b_trace;
            if ((size_t)max_chunk_words * 4 > tb->cap) fits = dev_reserve(ctx, *tb, (size_t)max_chunk_words * 4) == ISOF_OK;
        }
        if (fits) break;
        // the device is fuller than at first use (the caller's own tensors, a fragmented heap): plan smaller chunks
        if (attempt == 3 || budget_words <= pair_max_words)
            return isof_fail(ctx, ISOF_ERR_NOMEM, "trace scratch of %lld bytes does not fit (largest pair/chunk)",
                             (long long)max_chunk_words * 4);
        budget_words = std::max(budget_words / 2, pair_max_words);
        ctx->prof.trace_replans++;
        // hand every lane back first: a lane that did get its (too large) share would otherwise keep it
        for (int w = 0; w < isof_ctx::MAX_LANES; ++w) {
            DevBuf *tb = w ? &ctx->b_trace_x[w - 1] : &ctx->b_trace;
            if (tb->p) { CUDA_TRY(ctx, cudaFree(tb->p)); tb->p = nullptr; tb->cap = 0; }
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

    A.trace = (uint32_t *)ctx->b_trace.p; A.bnd = (int2 *)ctx->b_bnd.p; A.bnd_stride = bnd_stride;
    TRY(dev_reserve(ctx, ctx->b_layout, (size_t)P * 2));
    A.playout = (uint8_t *)ctx->b_layout.p;
    A.plastrr = A.playout + P;
    A.h_max_min_len = h_max_min_len;
    A.h_max_max_len = h_max_max_len;
    A.chunk_words = 0;

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
    cudaStream_t lanes[isof_ctx::MAX_LANES];
    lanes[0] = st;
    for (int w = 1; w < NL; ++w) lanes[w] = pipelined ? ctx->aux_stream[w - 1] : st;
    std::vector<cudaEvent_t> done((size_t)n_chunks, nullptr);
    cudaEvent_t fork = nullptr;
    // Every exit from here on -- the CUDA_TRY / TRY early returns included -- goes through this guard: an error return
    // first waits for all lanes, so that no kernel of the failed call still runs on an auxiliary stream when the next
    // call reuses the per-lane scratch, and the events are destroyed on every path.
    struct PipelineGuard {
        cudaStream_t *lanes; int n_lanes; std::vector<cudaEvent_t> &done; cudaEvent_t &fork; bool joined = false;
        ~PipelineGuard()
        {
            if (!joined) for (int w = 0; w < n_lanes; ++w) cudaStreamSynchronize(lanes[w]);
            if (fork) cudaEventDestroy(fork);
            for (cudaEvent_t e : done) if (e) cudaEventDestroy(e);
        }
    } guard{lanes, pipelined ? NL : 1, done, fork};
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
            A.n_rb = rb_counts + c;
        } else {   // task order of the chunk: stable radix sort of the pair ids by (eligibility, reference)
            DevBuf &kb = w ? ctx->b_taskkey_x[w - 1] : ctx->b_taskkey;
            TRY(dev_reserve(ctx, kb, (sizeof(uint32_t) * 2 + sizeof(int32_t)) * (size_t)cnt + 64));
            uint32_t *key_in = (uint32_t *)kb.p, *key_out = key_in + cnt;
            int32_t *val_in = (int32_t *)(key_out + cnt);
            int32_t *order = (int32_t *)ctx->b_order.p;
            LaunchTimer t(ctx, SLOT_OTHER, cs);
            task_key_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, cs>>>(pwild, d_pb, ps, cnt, key_in, val_in, elig_counts + c, rb_counts + c);
            size_t tso = 0;
            cub::DeviceRadixSort::SortPairs(nullptr, tso, key_in, key_out, val_in, order + ps, (int)cnt, 0, 32, cs);
            TRY(dev_reserve(ctx, c_scan, tso));
            CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(c_scan.p, tso, key_in, key_out, val_in, order + ps, (int)cnt, 0, 32, cs));
            A.order = order;
            A.n_elig = elig_counts + c;
            A.n_rb = rb_counts + c;
        }
        {
            LaunchTimer t(ctx, SLOT_DP, cs);
            const int64_t tasks = cnt;               // upper bound on tickets (all pairs single); CTAs are persistent anyway
            if (ctx->profiling)
                packed_cells_kernel<<<(unsigned)((tasks + 255) / 256), 256, 0, cs>>>(pwild, pn, pm, d_pb, A.order, A.n_elig, A.n_rb, ps, pe, misc + 5);
            int blocks = (int)std::min<int64_t>((tasks + DP_WARPS - 1) / DP_WARPS, occ_blocks);
            switch (tb_variant) {
                case 0: isof_launch_dp_tb0(blocks, cs, A, ps, pe, counters + c); break;
                case 1: isof_launch_dp_tb1(blocks, cs, A, ps, pe, counters + c); break;
                case 2: isof_launch_dp_tb2(blocks, cs, A, ps, pe, counters + c); break;
                default: isof_launch_dp_tb3(blocks, cs, A, ps, pe, counters + c); break;
            }
        }
        if (may_rb) {       // the chunk's re-based 16-bit tasks (long reads); exits at once when the class is empty
            LaunchTimer t(ctx, SLOT_DP, cs);
            const int blocks = (int)std::min<int64_t>((cnt / 2 + DP_WARPS) / DP_WARPS, occ_blocks);
            isof_launch_dp_rb(tb_variant, blocks, cs, A, ps, rb_tickets + c);
        }
        {
            LaunchTimer t(ctx, SLOT_TB, cs);
            // one warp per pair is the faster walk (look-ahead), one thread per pair the lighter one: in a pipelined step a
            // chunk of 18 k pairs as 18 k warps takes the slots the next chunk's DP is waiting for (C2: -3 % on the step)
            const bool warp_tb = ctx->tb_warp == 1 || (ctx->tb_warp == 2 && cnt <= 4096);
            if (warp_tb) sg_traceback_warp_kernel<<<(unsigned)((cnt + 3) / 4), 128, 0, cs>>>(A, ps, pe);
            else sg_traceback_kernel<<<(unsigned)((cnt + 63) / 64), 64, 0, cs>>>(A, ps, pe);
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
    }
    guard.joined = true;
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
    if (pband) {
        // ---- second pass of the banded trace: pairs whose optimal path left the band are aligned again with the full
        // trace (same kernels, banding off) and their results scattered into place -- results stay exact
        unsigned long long *fail_cnt = misc + 13;
        CUDA_TRY(ctx, cudaMemsetAsync(fail_cnt, 0, 8, st));
        TRY(dev_reserve(ctx, ctx->b_fail_ids, sizeof(int32_t) * (size_t)P));
        int32_t *ids = (int32_t *)ctx->b_fail_ids.p;
        {
            LaunchTimer t(ctx, SLOT_OTHER);
            collect_failed_kernel<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(nruns, P, ids, fail_cnt);
        }
        CUDA_TRY(ctx, cudaMemcpyAsync(pin64, fail_cnt, 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(ctx, cudaStreamSynchronize(st));
        const int64_t nf = pin64[0];
        ctx->prof.band_pairs_retried += nf;
        if (nf > 0) {
            TRY(dev_reserve(ctx, ctx->b_fail, sizeof(int32_t) * (5 + ISOF_N_FEATURES) * (size_t)nf + sizeof(int64_t) * (size_t)(nf + 2)));
            int64_t *off2 = (int64_t *)ctx->b_fail.p;                 // run offsets of the second pass (nf + 1)
            int32_t *pa2 = (int32_t *)(off2 + nf + 2), *pb2 = pa2 + nf, *score2 = pb2 + nf, *endq2 = score2 + nf, *endr2 = endq2 + nf;
            int32_t *feat2 = endr2 + nf;
            // a call that returns the runs: the first pass's stream and offsets are set aside, the second pass writes a
            // stream of its own, and both are spliced into the caller's buffer at the rebuilt offsets
            uint32_t *old_stream = nullptr, *cig2 = nullptr;
            int64_t *old_off = nullptr;
            int32_t *kidx = nullptr;
            int64_t cap2 = 0, first_total = 0;
            if (d_cigar) {
                CUDA_TRY(ctx, cudaMemcpyAsync(pin64 + 4, d_cigar_off + P, 8, cudaMemcpyDeviceToHost, st));
                CUDA_TRY(ctx, cudaStreamSynchronize(st));
                first_total = pin64[4];
                cap2 = nf * ((int64_t)max_n + max_m + 2);
                TRY(dev_reserve(ctx, ctx->b_splice, sizeof(uint32_t) * (size_t)(first_total + cap2) + sizeof(int64_t) * (size_t)(P + 1) +
                                                        sizeof(int32_t) * (size_t)P + 64));
                old_off = (int64_t *)ctx->b_splice.p;
                old_stream = (uint32_t *)(old_off + P + 1);
                cig2 = old_stream + first_total;
                kidx = (int32_t *)(cig2 + cap2);
                CUDA_TRY(ctx, cudaMemcpyAsync(old_off, d_cigar_off, sizeof(int64_t) * (size_t)(P + 1), cudaMemcpyDeviceToDevice, st));
                CUDA_TRY(ctx, cudaMemcpyAsync(old_stream, d_cigar, sizeof(uint32_t) * (size_t)first_total, cudaMemcpyDeviceToDevice, st));
                CUDA_TRY(ctx, cudaMemsetAsync(kidx, 0xff, sizeof(int32_t) * (size_t)P, st));
            }
            {
                LaunchTimer t(ctx, SLOT_OTHER);
                gather_pairs_kernel<<<(unsigned)((nf + 255) / 256), 256, 0, st>>>(ids, nf, d_pa, d_pb, pa2, pb2);
            }
            // the run counts of the first pass survive in the (now idle) band table; the second pass reuses `nruns`
            int32_t *saved_nr = pband;
            CUDA_TRY(ctx, cudaMemcpyAsync(saved_nr, nruns, sizeof(int32_t) * (size_t)P, cudaMemcpyDeviceToDevice, st));
            const int saved = ctx->band_w0;
            ctx->band_w0 = 0;
            const int rc2 = isof_sg_align_core(ctx, d_cat, d_off, n_seqs, n_bases, pa2, pb2, nf, match, mismatch, gap_open, gap_ext,
                                               score2, endq2, endr2, cig2, cap2, d_cigar ? off2 : nullptr,
                                               d_features ? feat2 : nullptr, delta_len, nullptr);
            ctx->band_w0 = saved;
            if (rc2 != ISOF_OK) return rc2;
            {
                LaunchTimer t(ctx, SLOT_OTHER);
                scatter_results_kernel<<<(unsigned)((nf + 255) / 256), 256, 0, st>>>(ids, nf, score2, endq2, endr2,
                                                                                   d_features ? feat2 : nullptr, d_score, d_endq,
                                                                                   d_endr, d_features);
                if (d_cigar_off) {
                    scatter_nruns_kernel<<<(unsigned)((nf + 255) / 256), 256, 0, st>>>(ids, nf, nruns, saved_nr);
                    rebuild_offsets_kernel<<<1, 1024, 0, st>>>(saved_nr, P, (long long *)d_cigar_off);
                }
            }
            CUDA_TRY(ctx, cudaGetLastError());
            CUDA_TRY(ctx, cudaStreamSynchronize(st));
            if ((h_total_runs || d_cigar) && d_cigar_off) {
                CUDA_TRY(ctx, cudaMemcpyAsync(pin64, d_cigar_off + P, 8, cudaMemcpyDeviceToHost, st));
                CUDA_TRY(ctx, cudaStreamSynchronize(st));
                if (h_total_runs) *h_total_runs = pin64[0];
            }
            if (d_cigar) {
                if (pin64[0] > cigar_cap)
                    return isof_fail(ctx, ISOF_ERR_CAPACITY, "cigar_runs holds %lld runs, %lld needed", (long long)cigar_cap,
                                     (long long)pin64[0]);
                LaunchTimer t(ctx, SLOT_TB);
                mark_failed_kernel<<<(unsigned)((nf + 255) / 256), 256, 0, st>>>(ids, nf, kidx);
                splice_runs_kernel<<<(unsigned)((P + 3) / 4), 128, 0, st>>>(old_stream, old_off, cig2, off2, kidx, d_cigar_off, P, d_cigar);
                CUDA_TRY(ctx, cudaGetLastError());
                CUDA_TRY(ctx, cudaStreamSynchronize(st));
            }
        }
    }
    return ISOF_OK;
}
