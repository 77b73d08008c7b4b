// sg_short.cu -- kernel (b), short-pair regime: one THREAD per task (two alignments in the 16-bit halves of every
// register), no wavefront.
//
// The long-pair kernel (sg_dp.cuh) spreads one task over the 32 lanes of a warp as an anti-diagonal wavefront; its
// 31-step fill and drain are noise at 2 kb but dominate at the ~100 x 100 pairs of bubble popping
// (SimplifyGraph.py:630-664, p50 99 x 99): 0.64 TCUPS at 100 bp against 3.3 TCUPS at 2 kb.  Here the parallelism is
// BETWEEN pairs instead: every lane owns its own task and sweeps its DP matrix in blocks of 8 rows (the same packed
// cell update as dp_pair16: tagged 4*score + 2-bit provenance, DPX max, trace bits gathered by IMADs), carrying the
// bottom row of a block to the next one through a per-lane line buffer.  No shuffles, no fill/drain, no idle lanes:
// the 64 alignments of a warp ("group") are sorted by shape, so all lanes run the same trip counts.
//   trace      uint4[(rb * M/2 + j/2) * 32 + lane] = {A(j even), B(j even), A(j odd), B(j odd)}   rb = row block, j =
//              column, M = longest reference of the group (a multiple of 4) -> every two column steps of a warp are one
//              coalesced 512 B store; 0.5 B per cell, no padding steps (the long-pair layout spends (m/2 + 31) * 256 B
//              per stripe whatever the row count).  Two columns of an alignment share a 32 B sector, which halves
//              the sector traffic of the traceback (it is bandwidth-bound on 32 B sectors at these sizes)
//   boundary   bnd[(j * 32 + lane)] = (H, F) of the block's bottom row, per warp, L1/L2 resident
//   CIGAR      traceback: one thread per pair (same walk as sg_traceback_kernel), runs written backwards into the
//              pair's slot of the group's run area; features / CIGAR copy kernels are shared with the long path.
// A call takes this path when EVERY pair is short (both lengths <= SHORT_MAX) and inside the 16-bit range; wildcard
// characters are handled by the WILD variant (substitution score 0).  Tie-break table: same template switch as the
// long kernel (TieBreakK).
#include "sg_dp.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace {

constexpr int SHORT_GROUP = 64;              // alignments per warp: 32 lanes x 2 halves
constexpr uint32_t WBIT = (4u << CS16) | (4u << (CS16 + 16));      // code 4 = outside ACGT

struct ShortParams {
    const uint8_t *codes;
    const int64_t *seq_off;
    const int32_t *pa, *pb, *pn, *pm;
    const int32_t *order;          // pair ids sorted by (row blocks, reference length)
    int64_t P;
    const int64_t *goff;           // n_groups+1 word offsets of the group regions
    const int32_t *gM, *gNRB;      // per group: longest reference, most row blocks
    uint32_t *trace;
    uint2 *bnd;                    // per resident warp: bnd_stride uint2
    int64_t bnd_stride;
    int32_t *score, *endq, *endr;
    int64_t *pbase;                // per pair: word offset of its (rb 0, column 0) trace word
    int32_t *pM;                   // per pair: M of its group (row-block stride = M * 64 words)
    int64_t *ptail;                // per pair: word offset one past its run slot
    uint32_t h_eo, h_ee, h_fo, h_fe, h_sm, h_sx;
    int tb_end_col_first;
    int n_groups;
};

// key = (row blocks, reference length): groups of 64 consecutive pairs of the sorted order share their loop bounds
__global__ void short_key_kernel(const int32_t *__restrict__ pn, const int32_t *__restrict__ pm, int64_t P,
                                 uint32_t *__restrict__ key, int32_t *__restrict__ val)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    key[p] = ((uint32_t)((pn[p] + R - 1) / R) << 10) | (uint32_t)pm[p];
    val[p] = (int32_t)p;
}

__global__ void short_group_kernel(const int32_t *__restrict__ pn, const int32_t *__restrict__ pm,
                                   const int32_t *__restrict__ order, int64_t P, int n_groups, int32_t *__restrict__ gM,
                                   int32_t *__restrict__ gNRB, int64_t *__restrict__ gwords)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g > n_groups) return;
    if (g == n_groups) { gwords[g] = 0; return; }
    int M = 1, nrb = 1;
    for (int s = 0; s < SHORT_GROUP; ++s) {
        const int64_t x = (int64_t)g * SHORT_GROUP + s;
        if (x >= P) break;
        const int p = order[x];
        M = max(M, pm[p]);
        nrb = max(nrb, (pn[p] + R - 1) / R);
    }
    M = (M + 3) & ~3;                      // the column loop runs in blocks of four (one 32-bit load of reference codes)
    gM[g] = M; gNRB[g] = nrb;
    // trace part + run area (one slot of n_max + m_max + 2 words per alignment)
    gwords[g] = (int64_t)nrb * M * SHORT_GROUP + (int64_t)SHORT_GROUP * (nrb * R + M + 2);
}

template <bool WILD, int TBV>
__global__ void __launch_bounds__(32, 16)
sg_dp_short_kernel(const ShortParams S, unsigned long long *__restrict__ counter)
{
    constexpr uint32_t TAG_E = TieBreakK<TBV>::TAG_E, TAG_F = TieBreakK<TBV>::TAG_F;
    const int lane = threadIdx.x;
    uint2 *__restrict__ bnd = S.bnd + (size_t)blockIdx.x * S.bnd_stride + lane;
    const uint32_t C3 = (3u << CS16) | (3u << (CS16 + 16));
    const uint32_t h_eo = S.h_eo, h_ee = S.h_ee, h_sm = S.h_sm, h_sx = S.h_sx;
    const uint32_t h_fo = TieBreakK<TBV>::OPEN_TIE ? S.h_fo : h_eo, h_fe = TieBreakK<TBV>::OPEN_TIE ? S.h_fe : h_ee;
    for (;;) {
        unsigned long long tkt = 0;
        if (lane == 0) tkt = atomicAdd(counter, 1ull);
        tkt = __shfl_sync(0xffffffffu, tkt, 0);
        if (tkt >= (unsigned long long)S.n_groups) break;
        const int g = (int)tkt;
        const int M = S.gM[g], NRB = S.gNRB[g];
        const int64_t xA = (int64_t)g * SHORT_GROUP + lane, xB = xA + 32;
        const bool hasA = xA < S.P, hasB = xB < S.P;
        const int pA = hasA ? S.order[xA] : 0, pB = hasB ? S.order[xB] : 0;
        const int nA = hasA ? S.pn[pA] : 0, mA = hasA ? S.pm[pA] : 0, nB = hasB ? S.pn[pB] : 0, mB = hasB ? S.pm[pB] : 0;
        const uint8_t *__restrict__ qA = S.codes + S.seq_off[S.pa[pA]];
        const uint8_t *__restrict__ rA = S.codes + S.seq_off[S.pb[pA]];
        const uint8_t *__restrict__ qB = S.codes + S.seq_off[S.pa[pB]];
        const uint8_t *__restrict__ rB = S.codes + S.seq_off[S.pb[pB]];
        const int64_t gbase = S.goff[g];
        uint4 *__restrict__ tr = reinterpret_cast<uint4 *>(S.trace + gbase) + lane;
        const int64_t run_area = gbase + (int64_t)NRB * M * SHORT_GROUP;
        const int slot_words = NRB * R + M + 2;
        if (hasA) { S.pbase[pA] = gbase + 4 * lane; S.pM[pA] = M; S.ptail[pA] = run_area + (int64_t)(lane + 1) * slot_words; }
        if (hasB) { S.pbase[pB] = gbase + 4 * lane + 1; S.pM[pB] = M; S.ptail[pB] = run_area + (int64_t)(lane + 33) * slot_words; }
        int bcolA = INT_MIN, bcolA_i = 0, browA = INT_MIN, browA_j = 0;
        int bcolB = INT_MIN, bcolB_i = 0, browB = INT_MIN, browB_j = 0;
        const int lastA_rb = (nA - 1) >> 3, lastA_r = (nA - 1) & 7, lastB_rb = (nB - 1) >> 3, lastB_r = (nB - 1) & 7;
        for (int rb = 0; rb < NRB; ++rb) {
            const int i0 = rb * R;
            uint32_t qc[R], Hl[R], Eh[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const uint32_t a = (i0 + r < nA) ? (uint32_t)qA[i0 + r] : 0u;
                const uint32_t b = (i0 + r < nB) ? (uint32_t)qB[i0 + r] : 0u;
                qc[r] = (a << CS16) | (b << (CS16 + 16));
                Hl[r] = 0u;                       // H(i,-1) = 0: free leading gap
                Eh[r] = NEG16 | TAG_E;
            }
            uint32_t diag0 = 0u;                  // H(i0-1,-1) = 0
            const bool has_prev = rb > 0, has_next = rb + 1 < NRB;
            const bool trackA = hasA && rb == lastA_rb, trackB = hasB && rb == lastB_rb;
            // reference codes: one aligned 32-bit load per four columns and alignment (byte loads would put 32 distinct
            // sectors per instruction on the LSU: the first version of this kernel was LSU-bound at 1.2 TCUPS); the
            // sequences start at arbitrary byte offsets of the pool, so two consecutive words are funnel-shifted
            const uint32_t shA = (uint32_t)((uintptr_t)rA & 3) * 8, shB = (uint32_t)((uintptr_t)rB & 3) * 8;
            const uint32_t *__restrict__ wA = reinterpret_cast<const uint32_t *>((uintptr_t)rA & ~(uintptr_t)3);
            const uint32_t *__restrict__ wB = reinterpret_cast<const uint32_t *>((uintptr_t)rB & ~(uintptr_t)3);
            uint32_t a0 = (mA > 0) ? __ldg(wA) : 0u, b0 = (mB > 0) ? __ldg(wB) : 0u;
            uint2 b_next = make_uint2(0u, NEG16 | TAG_F);
            if (has_prev) b_next = bnd[0];
            uint4 *__restrict__ tp = tr + (size_t)rb * (M / 2) * 32;
            uint32_t wA_even = 0u, wB_even = 0u;
            for (int j4 = 0; j4 < M; j4 += 4) {
                // word k+1 starts at byte 4(k+1) - misalignment of the sequence: load it only while inside the sequence
                const int nextA = j4 + 4 - (int)(shA >> 3), nextB = j4 + 4 - (int)(shB >> 3);
                const uint32_t a1 = (nextA < mA) ? __ldg(wA + (j4 >> 2) + 1) : 0u;
                const uint32_t b1 = (nextB < mB) ? __ldg(wB + (j4 >> 2) + 1) : 0u;
                const uint32_t bytesA = __funnelshift_r(a0, a1, shA), bytesB = __funnelshift_r(b0, b1, shB);
                a0 = a1; b0 = b1;
                // the word after next: into L1 now, so that the load above hits it one block later (no register held)
                if (nextA + 4 < mA) asm volatile("prefetch.global.L1 [%0];" ::"l"(wA + (j4 >> 2) + 2));
                if (nextB + 4 < mB) asm volatile("prefetch.global.L1 [%0];" ::"l"(wB + (j4 >> 2) + 2));
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const int j = j4 + c;
                const uint32_t rc = __byte_perm(bytesA, bytesB, (uint32_t)((c << 4) | ((4 + c) << 12))) & 0xff00ff00u;
                uint32_t up = b_next.x, fup = b_next.y;
                const uint32_t up0 = up;
                if (has_prev && j + 1 < M) b_next = bnd[(size_t)(j + 1) * 32];
                // ptxas sinks the load above to the end of the column (128 registers: it shortens the live range), which
                // exposed an L2 round trip per column (ncu: a third of all stall samples on the move that consumes it);
                // a prefetch three columns ahead costs no register and turns that load into an L1 hit
                if (has_prev && j + 4 < M) asm volatile("prefetch.global.L1 [%0];" ::"l"(bnd + (size_t)(j + 4) * 32));
                uint32_t diag = diag0;
                uint32_t acc_ef = 0u, acc_h = 0u;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const uint32_t left = Hl[r];
                    const uint32_t Ee = __vadd2(Eh[r], h_ee);
                    const uint32_t E = __viaddmax_s16x2(left, h_eo, Ee);
                    const uint32_t Ehn = E | TAG_E;
                    const uint32_t Fe = __vadd2(fup, h_fe);
                    const uint32_t F = __viaddmax_s16x2(up, h_fo, Fe);
                    const uint32_t Fh = F | TAG_F;
                    const uint32_t y = qc[r] ^ rc ^ C3;
                    uint32_t sc = __viaddmax_s16x2(y, h_sm, h_sx);
                    if (WILD) {          // either character outside ACGT: score 0 (diagonal tag 3), per half
                        const uint32_t w = ((qc[r] | rc) & WBIT) >> (CS16 + 2);          // 0/1 per half
                        const uint32_t wm = w * 0xffffu;
                        sc = (sc & ~wm) | (0x00030003u & wm);
                    }
                    const uint32_t D = __viaddmax_s16x2(diag, sc, Ehn);
                    const uint32_t H = __vmaxs2(D, Fh);
                    const uint32_t Hc = H & 0xfffcfffcu;
                    const uint32_t ef = (E | F) & 0x00030003u;
                    acc_ef = mad_u32(ef, 1u << (2 * r), acc_ef);
                    acc_h = mad_u32(H, 1u << (2 * r), acc_h);
                    acc_h = mad_u32(Hc, 0u - (1u << (2 * r)), acc_h);
                    diag = left;
                    Hl[r] = Hc;
                    Eh[r] = Ehn;
                    up = Hc;
                    fup = Fh;
                }
                diag0 = up0;
                if (has_next) bnd[(size_t)j * 32] = make_uint2(up, fup);
                if ((c & 1) == 0) {
                    wA_even = __byte_perm(acc_ef, acc_h, 0x5410); wB_even = __byte_perm(acc_ef, acc_h, 0x7632);
                } else {
                    ISOF_ASSERT((uint32_t *)(tp + (size_t)(j >> 1) * 32 + 1) <= S.trace + run_area, CHK_TRACE_STORE);
                    tp[(size_t)(j >> 1) * 32] = make_uint4(wA_even, wB_even, __byte_perm(acc_ef, acc_h, 0x5410),
                                                           __byte_perm(acc_ef, acc_h, 0x7632));
                }
                // ---- end cell bookkeeping (restated library: running maxima of the last row / last column)
                if (trackA && j < mA) {                    // last row of A: smallest j wins ties
                    const int v = lo16(pick_row<8>(Hl, lastA_r));
                    if (v > browA) { browA = v; browA_j = j; }
                }
                if (trackB && j < mB) {
                    const int v = hi16(pick_row<8>(Hl, lastB_r));
                    if (v > browB) { browB = v; browB_j = j; }
                }
                if (j == mA - 1 && hasA) {                 // last column of A: smallest i wins ties (rows ascend)
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const int v = lo16(Hl[r]);
                        if (i0 + r < nA && v > bcolA) { bcolA = v; bcolA_i = i0 + r; }
                    }
                }
                if (j == mB - 1 && hasB) {
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const int v = hi16(Hl[r]);
                        if (i0 + r < nB && v > bcolB) { bcolB = v; bcolB_i = i0 + r; }
                    }
                }
              }
            }
        }
        int sc, eq, er;
        if (hasA) {
            pick_end_cell(S.tb_end_col_first, browA, browA_j, bcolA, bcolA_i, nA, mA, sc, eq, er);
            S.score[pA] = sc >> 2; S.endq[pA] = eq; S.endr[pA] = er;
        }
        if (hasB) {
            pick_end_cell(S.tb_end_col_first, browB, browB_j, bcolB, bcolB_i, nB, mB, sc, eq, er);
            S.score[pB] = sc >> 2; S.endq[pB] = eq; S.endr[pB] = er;
        }
        __syncwarp();
    }
}

// traceback over the short layout: same walk as sg_traceback_kernel, other addresses.  Threads take the pairs in the
// sorted order of the DP groups: the 32 threads of a warp then walk the 32 alignments of one half of one group, whose
// trace words for the same cell sit 16 bytes apart -- pairs of similar shape stay near each other on their way up the
// diagonal, so a warp's loads fall into a few neighbouring lines instead of 32 unrelated sectors
__global__ void sg_traceback_short_kernel(const AlignParams P, const int64_t *__restrict__ pbase,
                                          const int32_t *__restrict__ pM, const int32_t *__restrict__ order,
                                          const int64_t n_pairs)
{
    const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n_pairs) return;
    const int64_t p = order[x];
    const int n = P.pn[p], m = P.pm[p];
    const uint8_t *s1 = P.cat + P.seq_off[P.pa[p]];
    const uint8_t *s2 = P.cat + P.seq_off[P.pb[p]];
    const uint32_t *trace = P.trace + pbase[p];
    const size_t pair_stride = 128, rb_stride = (size_t)(pM[p] / 2) * 128;      // words per column pair / per row block
    uint32_t *tail = P.trace + P.ptail[p];
    int i = P.endq[p], j = P.endr[p];
    int nr = 0;
    uint32_t cur_op = 0xffu, cur_len = 0;
#define PUSH(op, len)                                                                              \
    do {                                                                                           \
        if ((len) > 0) {                                                                           \
            if (cur_op == (uint32_t)(op)) cur_len += (uint32_t)(len);                              \
            else {                                                                                 \
                if (cur_len) { ++nr; tail[-nr] = (cur_len << 4) | cur_op; }                        \
                cur_op = (op); cur_len = (len);                                                    \
            }                                                                                      \
        }                                                                                          \
    } while (0)
    const int src_e = P.tb_src_e, e_mask = src_e, f_mask = 3 - src_e;       // split format: a state's flag bit is its class tag
    const int flag_opened = P.tb_open_tie;
    const uint32_t op_e = P.tb_swap_id ? ISOF_CIGAR_I : ISOF_CIGAR_D;
    const uint32_t op_f = P.tb_swap_id ? ISOF_CIGAR_D : ISOF_CIGAR_I;
    if (i + 1 == n) PUSH(op_e, m - 1 - j);
    else PUSH(op_f, n - 1 - i);
    int where = 3;
    while (i >= 0 || j >= 0) {
        if (i < 0) { PUSH(op_e, j + 1); break; }
        if (j < 0) { PUSH(op_f, i + 1); break; }
        const int r = i & 7;
        const uint32_t *wp = trace + (size_t)(i >> 3) * rb_stride + (size_t)(j >> 1) * pair_stride + (size_t)(j & 1) * 2;
        ISOF_ASSERT(wp >= P.trace && wp < tail - (n + m + 2), CHK_TRACE_LOAD);
        const uint32_t word = __ldcg(wp);
        const int nib = (((word >> (16 + 2 * r)) & 3) << 2) | ((word >> (2 * r)) & 3);
        if (where == 3) {
            const int src = nib >> 2;
            if (src == 3) {
                const int c1 = s1[i], c2 = s2[j];
                const int u1 = (c1 >= 'a' && c1 <= 'z') ? c1 - 32 : c1, u2 = (c2 >= 'a' && c2 <= 'z') ? c2 - 32 : c2;
                PUSH(u1 == u2 ? ISOF_CIGAR_EQ : ISOF_CIGAR_X, 1);
                --i; --j;
            } else where = src;
        } else if (where == src_e) {
            PUSH(op_e, 1);
            --j;
            where = (((nib & e_mask) != 0) != (flag_opened != 0)) ? src_e : 3;
        } else {
            PUSH(op_f, 1);
            --i;
            where = (((nib & f_mask) != 0) != (flag_opened != 0)) ? 3 - src_e : 3;
        }
    }
    if (cur_len) { ++nr; tail[-nr] = (cur_len << 4) | cur_op; }
#undef PUSH
    P.nruns[p] = nr;
}

template <bool WILD>
void launch_short(int tbv, int blocks, cudaStream_t st, const ShortParams &S, unsigned long long *counter)
{
    switch (tbv) {
        case 0: sg_dp_short_kernel<WILD, 0><<<blocks, 32, 0, st>>>(S, counter); break;
        case 1: sg_dp_short_kernel<WILD, 1><<<blocks, 32, 0, st>>>(S, counter); break;
        case 2: sg_dp_short_kernel<WILD, 2><<<blocks, 32, 0, st>>>(S, counter); break;
        default: sg_dp_short_kernel<WILD, 3><<<blocks, 32, 0, st>>>(S, counter); break;
    }
}

}  // namespace

unsigned isof_viol_short() { return isof_read_violations_tu(); }

// DP + traceback of an all-short call.  On return A.ptail (device array) holds every pair's run tail and A.trace the
// scratch they live in; the caller runs the shared features / CIGAR kernels over [0, P).
// Returns ISOF_ERR_NOMEM (without touching anything) when the trace does not fit `budget_words`: the caller then takes
// the long path, which chunks.
int isof_sg_short_path(isof_ctx *ctx, AlignParams &A, int64_t P, int max_n, int max_m, bool any_wild, int tb_variant,
                       int64_t budget_words)
{
    cudaStream_t st = ctx->stream;
    const int n_groups = (int)((P + SHORT_GROUP - 1) / SHORT_GROUP);
    // worst-case size first: no launch if it cannot fit
    {
        const int64_t nrb = (max_n + R - 1) / R;
        const int64_t m4 = (max_m + 3) & ~3;
        const int64_t worst = (int64_t)n_groups * (nrb * m4 * SHORT_GROUP + (int64_t)SHORT_GROUP * (nrb * R + m4 + 2));
        if (worst > budget_words) return ISOF_ERR_NOMEM;
    }
    TRY(dev_reserve(ctx, ctx->b_taskkey, (sizeof(uint32_t) * 2 + sizeof(int32_t)) * (size_t)P + 64));
    TRY(dev_reserve(ctx, ctx->b_order, sizeof(int32_t) * (size_t)P));
    TRY(dev_reserve(ctx, ctx->b_short_g, (sizeof(int32_t) * 2 + sizeof(int64_t) * 2) * (size_t)(n_groups + 1) + 64));
    TRY(dev_reserve(ctx, ctx->b_short_p, (sizeof(int64_t) * 2 + sizeof(int32_t)) * (size_t)P + 64));
    uint32_t *key_in = (uint32_t *)ctx->b_taskkey.p, *key_out = key_in + P;
    int32_t *val_in = (int32_t *)(key_out + P);
    int32_t *order = (int32_t *)ctx->b_order.p;
    int64_t *gwords = (int64_t *)ctx->b_short_g.p, *goff = gwords + (n_groups + 1);
    int32_t *gM = (int32_t *)(goff + (n_groups + 1)), *gNRB = gM + (n_groups + 1);
    int64_t *pbase = (int64_t *)ctx->b_short_p.p, *ptail = pbase + P;
    int32_t *pM = (int32_t *)(ptail + P);
    size_t t_sort = 0, t_scan = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, t_sort, key_in, key_out, val_in, order, (int)P, 0, 20, st);
    cub::DeviceScan::ExclusiveSum(nullptr, t_scan, gwords, goff, n_groups + 1, st);
    TRY(dev_reserve(ctx, ctx->b_scan_tmp, std::max(t_sort, t_scan) + 256));
    {
        LaunchTimer t(ctx, SLOT_OTHER);
        short_key_kernel<<<(unsigned)((P + 255) / 256), 256, 0, st>>>(A.pn, A.pm, P, key_in, val_in);
    }
    {
        LaunchTimer t(ctx, SLOT_OTHER);
        CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(ctx->b_scan_tmp.p, t_sort, key_in, key_out, val_in, order, (int)P, 0, 20, st));
    }
    {
        LaunchTimer t(ctx, SLOT_OTHER);
        short_group_kernel<<<(n_groups + 1 + 127) / 128, 128, 0, st>>>(A.pn, A.pm, order, P, n_groups, gM, gNRB, gwords);
    }
    {
        LaunchTimer t(ctx, SLOT_OTHER);
        CUDA_TRY(ctx, cub::DeviceScan::ExclusiveSum(ctx->b_scan_tmp.p, t_scan, gwords, goff, n_groups + 1, st));
    }
    CUDA_TRY(ctx, cudaGetLastError());
    int64_t *pin = (int64_t *)ctx->pinned;
    CUDA_TRY(ctx, cudaMemcpyAsync(pin, goff + n_groups, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    const int64_t total_words = pin[0];
    if (total_words > budget_words) return ISOF_ERR_NOMEM;
    if ((size_t)total_words * 4 + 64 > ctx->b_trace.cap) {
        int rc = dev_reserve(ctx, ctx->b_trace, (size_t)total_words * 4 + 64);
        if (rc != ISOF_OK) return ISOF_ERR_NOMEM;
    }
    const int blocks = std::min(n_groups, ctx->sm_count * 16);
    const int64_t bnd_stride = ((int64_t)((max_m + 3) & ~3) * 32 + 63) & ~63LL;
    TRY(dev_reserve(ctx, ctx->b_bnd, sizeof(uint2) * (size_t)bnd_stride * (size_t)blocks));
    TRY(dev_reserve(ctx, ctx->b_counters, 64));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->b_counters.p, 0, 64, st));
    ShortParams S;
    S.codes = A.codes; S.seq_off = A.seq_off; S.pa = A.pa; S.pb = A.pb; S.pn = A.pn; S.pm = A.pm; S.order = order; S.P = P;
    S.goff = goff; S.gM = gM; S.gNRB = gNRB; S.trace = (uint32_t *)ctx->b_trace.p; S.bnd = (uint2 *)ctx->b_bnd.p;
    S.bnd_stride = bnd_stride; S.score = A.score; S.endq = A.endq; S.endr = A.endr; S.pbase = pbase; S.pM = pM; S.ptail = ptail;
    S.h_eo = A.h_eo; S.h_ee = A.h_ee; S.h_fo = A.h_fo; S.h_fe = A.h_fe; S.h_sm = A.h_sm; S.h_sx = A.h_sx;
    S.tb_end_col_first = A.tb_end_col_first; S.n_groups = n_groups;
    {
        LaunchTimer t(ctx, SLOT_DP);
        if (any_wild) launch_short<true>(tb_variant, blocks, st, S, (unsigned long long *)ctx->b_counters.p);
        else launch_short<false>(tb_variant, blocks, st, S, (unsigned long long *)ctx->b_counters.p);
    }
    A.trace = (uint32_t *)ctx->b_trace.p;
    A.chunk_base = 0;
    A.ptail = ptail;
    A.pband = nullptr;
    {
        LaunchTimer t(ctx, SLOT_TB);
        sg_traceback_short_kernel<<<(unsigned)((P + 63) / 64), 64, 0, st>>>(A, pbase, pM, order, P);
    }
    CUDA_TRY(ctx, cudaGetLastError());
    return ISOF_OK;
}
