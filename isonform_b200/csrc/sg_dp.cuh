// sg_dp.cuh -- the DP kernel of kernel (b) (semi-global affine alignment, trace to HBM), templated on the
// tie-break table of the restated library (DESIGN.md "parasail tie-break table"):
//     TB bit 0  h_e_before_f     H source on ties: 0 DIAG > F > E (default), 1 DIAG > E > F
//     TB bit 1  gap_open_on_tie  gap state on open == extend: 0 extend (default), 1 open
// The default table (TB = 0) is instantiated in sg_align.cu; the other three in sg_dp_tb{1,2,3}.cu so that
// the four variants compile in parallel.  The end-cell order and the I/D letter convention of the table are
// run-time parameters (AlignParams::tb_*): they do not touch the cell update.
#pragma once
#include "common.cuh"
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

}  // namespace

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
    const unsigned long long *n_rb;        // per chunk: number of pairs of the re-based 16-bit class (device scalar)
    // tie-break table, run-time part (the compile-time part is the TB template parameter of the DP kernel)
    int tb_src_e;                          // trace source code of state E (walks the ref axis): 1, or 2 under h_e_before_f
    int tb_open_tie;                       // 1: the gap flags of the trace mean "opened" (gap_open_on_tie), 0: "extended"
    int tb_end_col_first;                  // end cell: last column first
    int tb_swap_id;                        // CIGAR letters of the gap states swapped
    const int64_t *ptail;                  // short-pair path: per pair, word offset one past its run slot (else nullptr)
    // banded trace (re-based class, feature-only calls): per pair {dmin, Tw}; Tw = 0: the whole matrix is stored.  Only
    // the cells with dmin <= j - i <= dmin + width are kept: per stripe s the steps tlo(s) .. tlo(s)+Tw-1 with
    // tlo(s) = max(0, (256 s + dmin) >> 1); the traceback reports a pair whose path leaves the band (nruns = -1) and
    // the host re-aligns those pairs with the full trace.
    const int32_t *pband;
};

// first stored step of stripe s under a band starting at diagonal dmin
__device__ __forceinline__ int band_tlo(const int s, const int dmin) { return max(0, (s * SROWS + dmin) >> 1); }

constexpr int SHORT_MAX = 384;             // both lengths <= SHORT_MAX: candidate for the short-pair regime (sg_short.cu)

// where the traceback leaves a pair's CIGAR runs (written backwards from this address)
__device__ __forceinline__ uint32_t *run_tail(const AlignParams &P, const int64_t p)
{
    return P.ptail ? P.trace + P.ptail[p] : P.trace + (P.ptrace[p + 1] - P.chunk_base);
}

namespace {

// compile-time part of the tie-break table
template <int TBV>
struct TieBreakK {
    static constexpr bool E_FIRST = (TBV & 1) != 0, OPEN_TIE = (TBV & 2) != 0;
    // class tags of the stored E / F in the packed path (the diagonal is 11): the larger one wins an H tie
    static constexpr uint32_t TAG_E = E_FIRST ? 0x00020002u : 0x00010001u;
    static constexpr uint32_t TAG_F = E_FIRST ? 0x00010001u : 0x00020002u;
};

// end cell from the running maxima of the last row / last column (values in tagged units; ties inside
// the row / column are already resolved towards the smallest j / i)
__device__ __forceinline__ void pick_end_cell(const int col_first, const int best_row, const int best_row_j,
                                              const int best_col, const int best_col_i, const int n, const int m,
                                              int &sc, int &eq, int &er)
{
    if (!col_first) {          // last row first; the last column only if strictly greater
        sc = best_row; eq = n - 1; er = best_row_j;
        if (best_col > sc) { sc = best_col; eq = best_col_i; er = m - 1; }
        else if (best_col == sc && er == m - 1 && best_col_i < eq) eq = best_col_i;
    } else {                   // last column first; the last row only if strictly greater
        sc = best_col; eq = best_col_i; er = m - 1;
        if (best_row > sc) { sc = best_row; eq = n - 1; er = best_row_j; }
    }
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
        int sc, eq, er;
        pick_end_cell(P.tb_end_col_first, best_row, best_row_j, best_col, best_col_i, n, m, sc, eq, er);
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
template <int RR, bool PROF, int TBV>
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
        const uint32_t Ehn = E | TieBreakK<TBV>::TAG_E;
        const uint32_t Fe = __vadd2(fup, K.h_fe);                   // tag 10 (extension)
        const uint32_t F = __viaddmax_s16x2(up, K.h_fo, Fe);        // open: tag 00; bit 1 = extended, ties extend
        const uint32_t Fh = F | TieBreakK<TBV>::TAG_F;
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

// RB (re-based 16-bit lanes, for pairs beyond the plain 16-bit range such as 10 kb reads): the packed halves hold
// value - base, where base (one 32-bit scalar per alignment, a multiple of 4 so that tags survive) follows the scores of
// the wavefront.  Neighbouring DP cells differ by at most match + open per row / column (a gap of one more character
// costs at most `open`, a diagonal step gains at most `match`), so the cells of a 256-row stripe and its 64-column lane
// skew stay within +-12.6 k tagged units of lane 15's first row -- the warp re-centres on it every 32 steps (16 in fill
// and drain), with saturating subtraction so that the -32000 sentinels stay sentinels.  Stripe boundary rows travel
// through the line buffer as absolute values modulo 2^16 (one VIADD per value on either side): modular arithmetic
// gives the right re-based value because the true one fits.  End-cell maxima are tracked as absolute 32-bit values.
// Every comparison still happens between values of the same base, so scores, ties and trace bits are those of the
// 32-bit path, bit for bit.
// CHASE (lone pairs of the fused small-call kernel): one warp per STRIPE, each in a block of its own.  The warp of stripe
// s reads the boundary row that the warp of stripe s - 1 is still writing -- every stripe has its own line buffer -- and
// follows it at the distance of the lane skew.  No flags and no fences on that path: the line buffers start out filled
// with a value no boundary cell can hold (CHASE_EMPTY: H = -32640 in both halves, below every reachable score), a
// boundary cell is ONE 8-byte store, and the consumer's lane 0 simply re-reads (ld.volatile, L2) until the cell is there.
// A 4000 x 4000 pair then takes ~T + 15 * 35 steps instead of 16 * T.  The warp returns its end-cell candidates instead of
// writing results.  (A first version published "columns written so far" with st.release every 8 steps: each release had
// to wait for the warp's outstanding trace stores, and the chain ran 6x slower than this one.)
constexpr uint32_t CHASE_EMPTY = 0x80808080u;          // cudaMemset(line buffers, 0x80)
struct ChaseCtl {
    int s;                             // the one stripe this warp computes
    int2 *bnd_in, *bnd_out;            // boundary rows: stripe s-1's bottom row (read), this stripe's (written)
    int2 *ring;                        // shared memory, 64 entries: the settled batches of the boundary row
    int bcol, bcol_i, brow, brow_j;    // out: best of the last column over this stripe's rows; best of the last row
    long long spins, spin_cycles;      // out (latency accounting): re-reads of a batch that was not there yet, cycles spent on them
};
__device__ __forceinline__ int2 ld_volatile_int2(const int2 *p)
{
    int2 v;
    // no "memory" clobber: the compiler stays free to schedule the step's other loads (profile, codes) across this one
    // a WEAK load that bypasses L1 (L2 is the point of coherence): fresh data on every execution, and -- unlike
    // ld.volatile / ld.relaxed.gpu -- no ordering against the warp's outstanding trace stores
    asm volatile("ld.global.cg.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned *p, const unsigned v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int TBV, bool RB, bool CHASE = false>
__device__ __forceinline__ void dp_pair16(const AlignParams &P, const int64_t pA, const int64_t pB, const int lane,
                                          int2 *__restrict__ bnd, uint4 *__restrict__ prof_warp, ChaseCtl *ch = nullptr)
{
    const unsigned FULL = 0xffffffffu;
    const int2 *__restrict__ bnd_r = CHASE ? ch->bnd_in : bnd;       // boundary row read by lane 0
    int2 *__restrict__ bnd_w = CHASE ? ch->bnd_out : bnd;            // boundary row written by lane 31
    static_assert(!(CHASE && RB), "chase mode exists for the plain 16-bit lanes only");
    auto bnd_put = [&](int2 *dst, const int2 v) {
        if (CHASE) __stcg(dst, v); else *dst = v;
    };
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
    // CHASE: the boundary row comes in batches of 32 columns, one per lane, fetched a batch (16 steps) ahead of its use so
    // that the L2 round trip -- longer than a step -- never sits on the step's critical path; lane 0 picks its columns out
    // of the batch with shuffles.  A batch is re-read until none of its cells holds the empty marker.
    int2 ch_cur = make_int2(0, 0), ch_next = make_int2(0, 0);
    int ch_col0 = 0;                                                 // first column of ch_cur
    auto chase_issue = [&](const int col0) {                         // this lane's cell of the batch starting at col0
        const int col = col0 + lane;
        return (col < m) ? ld_volatile_int2(bnd_r + col) : make_int2(0, 0);
    };
    auto chase_settle = [&](int2 &v, const int col0) {
        const int col = col0 + lane;
        const long long c0 = clock64();
        // (both words are checked: an 8-byte store lands as one piece in practice, but PTX does not promise it for vectors)
        while (__any_sync(FULL, col < m && ((uint32_t)v.x == CHASE_EMPTY || (uint32_t)v.y == CHASE_EMPTY))) {
            if (col < m && ((uint32_t)v.x == CHASE_EMPTY || (uint32_t)v.y == CHASE_EMPTY)) v = ld_volatile_int2(bnd_r + col);
            ++ch->spins;
        }
        ch->spin_cycles += clock64() - c0;
    };
    // boundary cell of column `col` (warp-uniform) for lane 0: a settled batch is parked in a 64-entry ring in shared
    // memory, from which lane 0 reads like the other modes read the line buffer; the batches rotate when col runs past
    // the ring is addressed in the shared window directly (a generic pointer costs an S2UR per access: the shared window
    // base of the CTA -- ~100 cycles on the critical path of a lone warp's step)
    const uint32_t ring_sa = CHASE ? (uint32_t)__cvta_generic_to_shared(ch->ring) : 0u;
    auto ring_put = [&](const int col, const int2 v) {
        asm volatile("st.shared.v2.s32 [%0], {%1, %2};" ::"r"(ring_sa + ((uint32_t)(col & 63) << 3)), "r"(v.x), "r"(v.y) : "memory");
    };
    auto ring_get = [&](const int col) {
        int2 v;
        asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(ring_sa + ((uint32_t)(col & 63) << 3)));
        return v;
    };
    auto chase_rotate = [&](const int col) {
        if (col - ch_col0 >= 32) {
            ch_col0 += 32;
            ch_cur = ch_next;
            chase_settle(ch_cur, ch_col0);
            ring_put(ch_col0 + lane, ch_cur);
            __syncwarp();
            ch_next = chase_issue(ch_col0 + 32);
        }
    };
    auto chase_take = [&](const int col) {
        chase_rotate(col);
        return ring_get(col);
    };
    P16Const K;
    // the F constants equal the E constants (same penalties, tags come from the stored values): one register each
    // (under gap_open_on_tie the opening / extension addends carry the state's own tag and differ)
    K.h_eo = P.h_eo; K.h_ee = P.h_ee; K.h_sm = P.h_sm; K.h_sx = P.h_sx;
    if (TieBreakK<TBV>::OPEN_TIE) { K.h_fo = P.h_fo; K.h_fe = P.h_fe; } else { K.h_fo = K.h_eo; K.h_fe = K.h_ee; }
    constexpr uint32_t TAG_E = TieBreakK<TBV>::TAG_E, TAG_F = TieBreakK<TBV>::TAG_F;
    int bcolA = INT_MIN, bcolA_i = 0, browA = INT_MIN, browA_j = 0;
    int bcolB = INT_MIN, bcolB_i = 0, browB = INT_MIN, browB_j = 0;
    if (lane == 0) { P.playout[pA] = PC | TRACE_FMT_SPLIT; P.playout[pB] = PC | TRACE_FMT_SPLIT; }
    // steady range of t: lane 31 has started (t >= 31) and lane 0's look-ahead block is inside both
    // references: PC*t + 2*PC <= mmin
    const int t_steady_end = (mmin - 2 * PC) / PC + 1;          // exclusive; may be <= 31 (then no steady range)
    // banded trace (RB only; see AlignParams::pband): window of stored steps per stripe
    int dminA = 0, twA = 0, dminB = 0, twB = 0;
    if (RB && P.pband) { dminA = P.pband[2 * pA]; twA = P.pband[2 * pA + 1]; dminB = P.pband[2 * pB]; twB = P.pband[2 * pB + 1]; }

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
        int baseA = 0, baseB = 0;                    // RB: true value = packed half + base; every stripe starts at 0
        uint32_t basepk = 0u;                        // (baseA & 0xffff) | (baseB << 16)
#pragma unroll
        for (int r = 0; r < RR; ++r) {
            const uint32_t a = (i0 + r < nA) ? (uint32_t)qA[i0 + r] : 0u;
            const uint32_t b = (i0 + r < nB) ? (uint32_t)qB[i0 + r] : 0u;
            S.qc[r] = (a << CS16) | (b << (CS16 + 16));
            S.Hl[r] = 0u;
            S.Eh[r] = NEG16 | TAG_E;
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
            S.Fbot[c] = NEG16 | TAG_F;
            S.rc_next[c] = 0u;
            S.b_next[c] = make_int2(0, (int)(NEG16 | TAG_F));
            if (lane == 0) {
                const uint32_t a = (c < mA) ? (uint32_t)rA[c] : 0u, b = (c < mB) ? (uint32_t)rB[c] : 0u;
                S.rc_next[c] = PROF ? a : ((a << CS16) | (b << (CS16 + 16)));
                if (!CHASE && has_prev && c < m) S.b_next[c] = __ldcg(&bnd_r[c]);
            }
        }
        if (CHASE && has_prev) {                     // first two batches of the boundary row; columns 0 .. PC-1 feed step 0
            ch_col0 = 0;
            ch_cur = chase_issue(0);
            ch_next = chase_issue(32);
            chase_settle(ch_cur, 0);
            __syncwarp();
            ring_put(lane, ch_cur);
            __syncwarp();
#pragma unroll
            for (int c = 0; c < PC; ++c) {
                const int2 v = ring_get(c);
                if (lane == 0 && c < m) S.b_next[c] = v;
            }
        }
        // banded pairs keep the steps tlo .. tlo + win - 1 of the stripe (stripe pitch = win); others all T steps
        const int tloA = (RB && twA) ? band_tlo(s, dminA) : 0, winA = (RB && twA) ? twA : TA;
        const int tloB = (RB && twB) ? band_tlo(s, dminB) : 0, winB = (RB && twB) ? twB : TB;
        uint2 *tpA = reinterpret_cast<uint2 *>(traceA) + ((ptrdiff_t)s * winA - tloA) * 32 + lane;       // tp[t * 32]: step t
        uint2 *tpB = reinterpret_cast<uint2 *>(traceB) + ((ptrdiff_t)s * winB - tloB) * 32 + lane;
        uint32_t snap[RR];                           // H of the last column (A low half, B high half), set in the drain
#pragma unroll
        for (int r = 0; r < RR; ++r) snap[r] = 0x80008000u;

        // RB: re-centre every packed value of the warp on lane 15's first row (see the comment at the function head)
        auto rebase = [&]() {
            const uint32_t rep = __shfl_sync(FULL, S.Hl[0], 15);
            const int dA = lo16(rep) & ~3, dB = hi16(rep) & ~3;
            const uint32_t dpk = ((uint32_t)dA & 0xffffu) | ((uint32_t)dB << 16);
            auto adj = [&](const uint32_t v) { return __vmaxs2(__vsubss2(v, dpk), NEG16); };
#pragma unroll
            for (int r = 0; r < RR; ++r) { S.Hl[r] = adj(S.Hl[r]); S.Eh[r] = adj(S.Eh[r]); }
#pragma unroll
            for (int c = 0; c < PC; ++c) {
                S.Hbot[c] = adj(S.Hbot[c]); S.Fbot[c] = adj(S.Fbot[c]);
                S.b_next[c].x = (int)adj((uint32_t)S.b_next[c].x); S.b_next[c].y = (int)adj((uint32_t)S.b_next[c].y);
            }
            S.diag0 = adj(S.diag0);
            baseA += dA; baseB += dB;
            basepk = __vadd2(basepk, dpk);
        };
        // boundary rows in the line buffer are absolute values modulo 2^16 (RB only)
        auto bnd_load = [&](const int2 *src) {
            int2 v = __ldcg(src);
            if (RB) { v.x = (int)__vsub2((uint32_t)v.x, basepk); v.y = (int)__vsub2((uint32_t)v.y, basepk); }
            return v;
        };
        auto bnd_value = [&](const uint32_t h, const uint32_t f) {
            return RB ? make_int2((int)__vadd2(h, basepk), (int)__vadd2(f, basepk)) : make_int2((int)h, (int)f);
        };

        // ---------------- general step (bounds checks + end-cell bookkeeping) ----------------
        auto general_step = [&](const int t, auto track_tag, auto row_tag) {
            constexpr bool TR = decltype(track_tag)::value;      // false: the step cannot reach a last column (fill steps before a steady range)
            constexpr bool TROW = decltype(row_tag)::value;      // false: not a last stripe, no last-row bookkeeping
            if (RB && (t & 15) == 0) rebase();
            const int jb = PC * (t - lane);
            uint32_t rc[PC], hu[PC], fu[PC];
#pragma unroll
            for (int c = 0; c < PC; ++c) {
                rc[c] = S.rc_next[c];
                hu[c] = __shfl_up_sync(FULL, S.Hbot[c], 1);
                fu[c] = __shfl_up_sync(FULL, S.Fbot[c], 1);
                if (lane == 0) {
                    hu[c] = (uint32_t)S.b_next[c].x; fu[c] = (uint32_t)S.b_next[c].y;
                }
            }
            if (CHASE && has_prev && PC * (t + 1) < m) {          // lane 0's boundary cells of the next step (warp-uniform)
#pragma unroll
                for (int c = 0; c < PC; ++c) {
                    const int col = PC * (t + 1) + c;
                    if (col < m) { const int2 v = chase_take(col); if (lane == 0) S.b_next[c] = v; }
                }
            }
            {
                const int jn = jb + PC;
                if ((unsigned)jn < (unsigned)m) {
#pragma unroll
                    for (int c = 0; c < PC; ++c) {
                        const uint32_t a = (jn + c < mA) ? (uint32_t)rA[jn + c] : 0u;
                        const uint32_t b = (jn + c < mB) ? (uint32_t)rB[jn + c] : 0u;
                        S.rc_next[c] = PROF ? a : ((a << CS16) | (b << (CS16 + 16)));
                        if (!CHASE && has_prev && lane == 0 && jn + c < m) S.b_next[c] = bnd_load(&bnd_r[jn + c]);
                    }
                }
            }
            if ((unsigned)jb < (unsigned)m) {
                uint32_t wA[PC], wB[PC];
#pragma unroll
                for (int c = 0; c < PC; ++c) {
                    p16_column<RR, PROF, TBV>(S, K, rc[c], hu[c], fu[c], (c == 0) ? S.diag0 : hu[c - 1], S.Hbot[c], S.Fbot[c], wA[c], wB[c], prof);
                    const int j = jb + c;
                    ISOF_ASSERT(j < P.bnd_stride, CHK_BND);
                    if (has_next && lane == 31 && j < m) bnd_put(&bnd_w[j], bnd_value(S.Hbot[c], S.Fbot[c]));
                    // last column of A / B: every lane passes it exactly once per stripe; keep the column (one LOP3 per
                    // row) and fold it into the running maximum after the stripe's loop (RB: the base may move before
                    // then, so the column is folded right here as absolute values)
                    if (TR && j == mA - 1) {
#pragma unroll
                        for (int r = 0; r < RR; ++r) {
                            if (RB) {
                                const int v = lo16(S.Hl[r]) + baseA;
                                if (liveA && i0 + r < nA && v > bcolA) { bcolA = v; bcolA_i = i0 + r; }
                            } else snap[r] = (snap[r] & 0xffff0000u) | (S.Hl[r] & 0x0000ffffu);
                        }
                    }
                    if (TR && j == mB - 1) {
#pragma unroll
                        for (int r = 0; r < RR; ++r) {
                            if (RB) {
                                const int v = hi16(S.Hl[r]) + baseB;
                                if (liveB && i0 + r < nB && v > bcolB) { bcolB = v; bcolB_i = i0 + r; }
                            } else snap[r] = (snap[r] & 0x0000ffffu) | (S.Hl[r] & 0xffff0000u);
                        }
                    }
                    if (TROW && trackA && j < mA) {         // last row of A: smallest j wins ties
                        const uint32_t hv = pick_row<RR>(S.Hl, last_rA);
                        const int v = lo16(hv) + (RB ? baseA : 0);
                        if (v > browA) { browA = v; browA_j = j; }
                    }
                    if (TROW && trackB && j < mB) {
                        const uint32_t hv = pick_row<RR>(S.Hl, last_rB);
                        const int v = hi16(hv) + (RB ? baseB : 0);
                        if (v > browB) { browB = v; browB_j = j; }
                    }
                }
                S.diag0 = hu[PC - 1];
                if (liveA && jb < mA && (!RB || (unsigned)(t - tloA) < (unsigned)winA)) {
                    ISOF_ASSERT((uint32_t *)(tpA + (ptrdiff_t)t * 32 + 1) <= P.trace + (P.ptrace[pA + 1] - P.chunk_base) - TRACE_SLACK, CHK_TRACE_STORE);
                    ISOF_ASSERT((uint32_t *)(tpA + (ptrdiff_t)t * 32) >= traceA, CHK_TRACE_STORE);
                    tpA[(ptrdiff_t)t * 32] = make_uint2(wA[0], wA[1]);
                }
                if (liveB && jb < mB && (!RB || (unsigned)(t - tloB) < (unsigned)winB)) {
                    ISOF_ASSERT((uint32_t *)(tpB + (ptrdiff_t)t * 32 + 1) <= P.trace + (P.ptrace[pB + 1] - P.chunk_base) - TRACE_SLACK, CHK_TRACE_STORE);
                    ISOF_ASSERT((uint32_t *)(tpB + (ptrdiff_t)t * 32) >= traceB, CHK_TRACE_STORE);
                    tpB[(ptrdiff_t)t * 32] = make_uint2(wB[0], wB[1]);
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
            const int2 *pbn = bnd_r + PC * 31 + PC;                // lane 0 only
            int2 *pbs = bnd_w + PC * (31 - 31);                    // lane 31 only: jb at t = 31
            uint2 *sA = tpA + (ptrdiff_t)31 * 32, *sB = tpB + (ptrdiff_t)31 * 32;
            auto steady_step = [&](const int t, const bool with_track) {
                uint32_t rc[PC], hu[PC], fu[PC];
#pragma unroll
                for (int c = 0; c < PC; ++c) {
                    rc[c] = S.rc_next[c];
                    hu[c] = __shfl_up_sync(FULL, S.Hbot[c], 1);
                    fu[c] = __shfl_up_sync(FULL, S.Fbot[c], 1);
                    if (lane == 0) {
                        hu[c] = (uint32_t)S.b_next[c].x; fu[c] = (uint32_t)S.b_next[c].y;
                    }
                }
#pragma unroll
                for (int c = 0; c < PC; ++c)
                {
                    ISOF_ASSERT(pa_ + c < rA + mA && pb_ + c < rB + mB && pa_ >= rA && pb_ >= rB, CHK_SEQ_LOAD);
                    S.rc_next[c] = PROF ? (uint32_t)pa_[c] : (((uint32_t)pa_[c] << CS16) | ((uint32_t)pb_[c] << (CS16 + 16)));
                }
                if (CHASE) {
                    if (has_prev) {
#pragma unroll
                        for (int c = 0; c < PC; ++c) { const int2 v = ring_get(PC * (t + 1) + c); if (lane == 0) S.b_next[c] = v; }
                    }
                } else if (has_prev && lane == 0) {
#pragma unroll
                    for (int c = 0; c < PC; ++c) S.b_next[c] = bnd_load(pbn + c);
                }
                uint32_t wA[PC], wB[PC];
#pragma unroll
                for (int c = 0; c < PC; ++c) {
                    p16_column<RR, PROF, TBV>(S, K, rc[c], hu[c], fu[c], (c == 0) ? S.diag0 : hu[c - 1], S.Hbot[c], S.Fbot[c], wA[c], wB[c], prof);
                    if (with_track) {
                        const int j = PC * (t - lane) + c;
                        if (trackA) {
                            const uint32_t hv = pick_row<RR>(S.Hl, last_rA);
                            const int v = lo16(hv) + (RB ? baseA : 0);
                            if (v > browA) { browA = v; browA_j = j; }
                        }
                        if (trackB) {
                            const uint32_t hv = pick_row<RR>(S.Hl, last_rB);
                            const int v = hi16(hv) + (RB ? baseB : 0);
                            if (v > browB) { browB = v; browB_j = j; }
                        }
                    }
                }
                S.diag0 = hu[PC - 1];
                if (!RB) {
                    ISOF_ASSERT((uint32_t *)(sA + 1) <= P.trace + (P.ptrace[pA + 1] - P.chunk_base) - TRACE_SLACK, CHK_TRACE_STORE);
                    ISOF_ASSERT((uint32_t *)(sB + 1) <= P.trace + (P.ptrace[pB + 1] - P.chunk_base) - TRACE_SLACK, CHK_TRACE_STORE);
                    ISOF_ASSERT((uint32_t *)sA >= traceA && (uint32_t *)sB >= traceB, CHK_TRACE_STORE);
                    *sA = make_uint2(wA[0], wA[1]);
                    *sB = make_uint2(wB[0], wB[1]);
                } else {              // banded trace: the step is kept only inside the stripe's window (warp-uniform tests)
                    if ((unsigned)(t - tloA) < (unsigned)winA) {
                        ISOF_ASSERT((uint32_t *)(sA + 1) <= P.trace + (P.ptrace[pA + 1] - P.chunk_base) - TRACE_SLACK && (uint32_t *)sA >= traceA, CHK_TRACE_STORE);
                        *sA = make_uint2(wA[0], wA[1]);
                    }
                    if ((unsigned)(t - tloB) < (unsigned)winB) {
                        ISOF_ASSERT((uint32_t *)(sB + 1) <= P.trace + (P.ptrace[pB + 1] - P.chunk_base) - TRACE_SLACK && (uint32_t *)sB >= traceB, CHK_TRACE_STORE);
                        *sB = make_uint2(wB[0], wB[1]);
                    }
                }
                if (has_next && lane == 31) {
#pragma unroll
                    for (int c = 0; c < PC; ++c) { ISOF_ASSERT(pbs + c < bnd_w + P.bnd_stride, CHK_BND); bnd_put(&pbs[c], bnd_value(S.Hbot[c], S.Fbot[c])); }
                }
                pa_ += PC; pb_ += PC; pbn += PC; pbs += PC; sA += 32; sB += 32;
            };
            if (CHASE && has_prev) {
                // blocks of steps that stay inside one batch of the boundary row: the rotation (settle the next batch, park
                // it in the ring, issue the one after) happens between blocks, the steps themselves only read the ring
                for (int t = 31; t < t2;) {
                    chase_rotate(PC * (t + 1));
                    const int te = min(t2, (ch_col0 + 32) / PC - 1);          // first step whose look-ahead leaves the batch
                    if (!track_stripe) {
#pragma unroll 2
                        for (; t < te; ++t) steady_step(t, false);
                    } else {
#pragma unroll 2
                        for (; t < te; ++t) steady_step(t, true);
                    }
                }
            } else if (RB) {
                for (int t = 31; t < t2;) {
                    rebase();
                    const int te = min(t + 32, t2);
                    if (!track_stripe) {
#pragma unroll 2
                        for (; t < te; ++t) steady_step(t, false);
                    } else {
#pragma unroll 2
                        for (; t < te; ++t) steady_step(t, true);
                    }
                }
            } else if (!track_stripe) {
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
        if (!RB && liveA) {
#pragma unroll
            for (int r = 0; r < RR; ++r) {
                const int v = lo16(snap[r]);
                if (i0 + r < nA && v > bcolA) { bcolA = v; bcolA_i = i0 + r; }
            }
        }
        if (!RB && liveB) {
#pragma unroll
            for (int r = 0; r < RR; ++r) {
                const int v = hi16(snap[r]);
                if (i0 + r < nB && v > bcolB) { bcolB = v; bcolB_i = i0 + r; }
            }
        }
        __syncwarp();
    };
    for (int s = CHASE ? ch->s : 0; s < (CHASE ? ch->s + 1 : ns); ++s) {
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
    if (CHASE) { ch->bcol = bcolA; ch->bcol_i = bcolA_i; ch->brow = browA; ch->brow_j = browA_j; return; }
    if (lane == 0) {
        int sc, eq, er;
        pick_end_cell(P.tb_end_col_first, browA, browA_j, bcolA, bcolA_i, nA, mA, sc, eq, er);
        P.score[pA] = sc >> 2; P.endq[pA] = eq; P.endr[pA] = er;
        pick_end_cell(P.tb_end_col_first, browB, browB_j, bcolB, bcolB_i, nB, mB, sc, eq, er);
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
template <int TBV>
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
        // tickets [0, n_pk): packed tasks over the eligible pairs (the odd pair of the class runs with both halves on
        // itself: a lone 2.7 kb pair on the 32-bit path is a 7 ms tail of one warp); the rest: one 32-bit pair per ticket
        const int64_t n0 = (int64_t)(*P.n_elig);
        const int64_t n_pk = (n0 + 1) / 2;
        if ((int64_t)tkt < n_pk) {
            const int64_t xA = start + 2 * (int64_t)tkt, xB = (2 * (int64_t)tkt + 1 < n0) ? xA + 1 : xA;
            dp_pair16<TBV, false>(P, P.order[xA], P.order[xB], lane, bnd, s_prof[warp]);
        } else {
            // single 32-bit pairs: everything that fits neither 16-bit class (sg_dp_rb_kernel takes the whole re-based
            // class, its odd pair included: a banded trace has only that layout)
            const int64_t n1 = (int64_t)(*P.n_rb);
            const int64_t s = (int64_t)tkt - n_pk;
            const int64_t x = start + n0 + n1 + s;
            if (x >= end) break;
            dp_pair32(P, P.order[x], lane, bnd, s_prof[warp]);
        }
    }
}

template <int TBV>
inline void launch_sg_dp(const int blocks, cudaStream_t st, const AlignParams &A, const int64_t start, const int64_t end,
                         unsigned long long *counter)
{
    sg_dp_kernel<TBV><<<blocks, DP_WARPS * 32, 0, st>>>(A, start, end, counter);
}

// The re-based class of a chunk: pairs order[start + n0 + 2k], order[start + n0 + 2k + 1], k < (n1 + 1) / 2 (dp_pair16<., true>).
// A kernel of its own: the extra live scalars of the re-based path would otherwise cost the plain path registers.
template <int TBV>
__global__ void __launch_bounds__(DP_WARPS * 32, DP_CTAS_PER_SM)
sg_dp_rb_kernel(const AlignParams P, const int64_t start, unsigned long long *__restrict__ counter)
{
    __shared__ uint4 s_prof[DP_WARPS][4 * 2 * 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int2 *bnd = P.bnd + (size_t)(blockIdx.x * DP_WARPS + warp) * P.bnd_stride;
    for (;;) {
        unsigned long long tkt = 0;
        if (lane == 0) tkt = atomicAdd(counter, 1ull);
        tkt = __shfl_sync(0xffffffffu, tkt, 0);
        const int64_t n0 = (int64_t)(*P.n_elig), n1 = (int64_t)(*P.n_rb);
        if ((int64_t)tkt >= (n1 + 1) / 2) break;
        // the odd pair of the class runs with both halves on itself (same stores twice, same values)
        const int64_t xA = start + n0 + 2 * (int64_t)tkt, xB = (2 * (int64_t)tkt + 1 < n1) ? xA + 1 : xA;
        dp_pair16<TBV, true>(P, P.order[xA], P.order[xB], lane, bnd, s_prof[warp]);
    }
}

}  // namespace

// one launcher per compile-time tie-break variant (sg_align.cu: 0, sg_dp_tb{1,2,3}.cu), and the debug-assertion
// masks of the variant translation units
void isof_launch_dp_tb0(int blocks, cudaStream_t st, const AlignParams &A, int64_t start, int64_t end, unsigned long long *counter);
void isof_launch_dp_tb1(int blocks, cudaStream_t st, const AlignParams &A, int64_t start, int64_t end, unsigned long long *counter);
void isof_launch_dp_tb2(int blocks, cudaStream_t st, const AlignParams &A, int64_t start, int64_t end, unsigned long long *counter);
void isof_launch_dp_tb3(int blocks, cudaStream_t st, const AlignParams &A, int64_t start, int64_t end, unsigned long long *counter);
unsigned isof_viol_dp_tb1();
unsigned isof_viol_dp_tb2();
unsigned isof_viol_dp_tb3();
unsigned isof_viol_short();
unsigned isof_viol_dp_rb();
void isof_launch_dp_rb(int tb_variant, int blocks, cudaStream_t st, const AlignParams &A, int64_t start, unsigned long long *counter);
int isof_sg_short_path(isof_ctx *ctx, AlignParams &A, int64_t P, int max_n, int max_m, bool any_wild, int tb_variant,
                       int64_t budget_words);
