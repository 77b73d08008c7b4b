// common.cuh -- context, error handling and small device helpers shared by the kernels.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <algorithm>
#include <vector>

#include "../../include/isonform_b200.h"

#define ISOF_VERSION 100

struct DevBuf {            // grow-only device scratch
    void *p = nullptr;
    size_t cap = 0;
};

struct isof_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    char err[512] = {0};
    size_t trace_budget = 0;          // bytes set by isof_set_trace_budget; 0 = automatic (40-55 % of free memory)
    size_t trace_free_at_first_use = 0;
    bool profiling = false;           // bracket every launch with CUDA events (isof_profile_*)
    bool disable_packed = false;      // force the 32-bit DP path (tests / A-B measurements)
    uint64_t span_hash_mask = 0x7fffffffffffffffULL;      // test knob: fewer hash bits force collisions in kernel (c)
    int band_w0 = 1024;               // half-width of the banded trace of long pairs beyond the free-end-gap diagonals; 0 = off
    int tb_warp = 2;                  // traceback of the regular path: 1 = one warp per pair (look-ahead), 0 = one thread per pair,
                                      // 2 = by chunk size (warp per pair up to 4096 pairs)
    int ed_tb_warp = 2;               // kernel (b') traceback: 0 thread per pair, 1 warp per pair beyond 128 rows, 2 by pair count
    int short_mode = 1;               // short-pair regime: 0 never, 1 when the call is large enough to pay (default), 2 always
    // tie-break table of the restated aligner (isof_set_tiebreak); all zero = the library as restated
    int tb_h_e_before_f = 0, tb_gap_open_on_tie = 0, tb_end_col_first = 0, tb_swap_id = 0;
    bool overlap = true;              // pipeline trace chunks over NLANES streams, each with its own scratch
    static constexpr int NLANES = 3;  // pipeline depth: chunk c runs on stream c % NLANES
    static constexpr int MAX_LANES = 6;   // depth for batches whose occupancy is bound by trace memory (10 kb reads)
    cudaStream_t aux_stream[MAX_LANES - 1] = {};
    isof_profile prof;
    // pending (start, stop, slot) event pairs when profiling
    struct Timed { cudaEvent_t a, b; int slot; };
    std::vector<Timed> timed;
    // scratch
    DevBuf b_codes, b_seqflags, b_pn, b_pm, b_words, b_nruns64, b_ptrace, b_scan_tmp, b_bnd, b_trace, b_nruns,
        b_counters, b_misc, b_runoff, b_status, b_layout, b_order, b_taskkey, b_short_g, b_short_p, b_band, b_fail_ids, b_fail, b_chase, b_splice;
    DevBuf b_taskkey_x[MAX_LANES - 1];
    DevBuf b_trace_x[MAX_LANES - 1], b_bnd_x[MAX_LANES - 1], b_runoff_x[MAX_LANES - 1], b_nruns64_x[MAX_LANES - 1],
        b_scan_tmp_x[MAX_LANES - 1];
    DevBuf b_h_bases, b_h_off, b_h_pa, b_h_pb, b_h_score, b_h_endq, b_h_endr, b_h_cig, b_h_cigoff, b_h_feat;
    DevBuf b_m_tmp, b_m_cnt, b_m_off;
    DevBuf b_s_mread, b_s_mkey, b_s_mpolya, b_s_mcnt, b_s_moff, b_s_key1, b_s_key2, b_s_keyt, b_s_keyo, b_s_idx0, b_s_idx1,
        b_s_idx2, b_s_forb, b_s_head, b_s_bid, b_s_sorttmp, b_s_rbatch, b_sh_batchoff;
    DevBuf b_ed[16];                  // kernel (b'), myers.cu: codes, symbols, layout, plan, sort, tasks, trace, runs
    DevBuf b_h_dist, b_ed_eqg;
    DevBuf b_sh_minpos, b_sh_minoff, b_sh_rread, b_sh_rp1, b_sh_rp2, b_sh_rcount, b_sh_rbucket, b_sh_sorted, b_sh_boff, b_sh_bdrop;
    std::vector<DevBuf *> all_bufs()
    {
        std::vector<DevBuf *> v = {&b_codes, &b_seqflags, &b_pn, &b_pm, &b_words, &b_nruns64, &b_ptrace, &b_scan_tmp, &b_bnd, &b_trace, &b_nruns,
                &b_counters, &b_misc, &b_runoff, &b_status, &b_layout, &b_order, &b_taskkey, &b_short_g, &b_short_p, &b_band, &b_fail_ids, &b_fail, &b_chase, &b_splice, &b_h_bases, &b_h_off, &b_h_pa, &b_h_pb, &b_h_score, &b_h_endq,
                &b_h_endr, &b_h_cig, &b_h_cigoff, &b_h_feat, &b_m_tmp, &b_m_cnt, &b_m_off, &b_s_mread, &b_s_mkey,
                &b_s_mpolya, &b_s_mcnt, &b_s_moff, &b_s_key1, &b_s_key2, &b_s_keyt, &b_s_keyo, &b_s_idx0, &b_s_idx1, &b_s_idx2,
                &b_s_forb, &b_s_head, &b_s_bid, &b_s_sorttmp, &b_s_rbatch, &b_sh_batchoff, &b_sh_minpos, &b_sh_minoff, &b_sh_rread, &b_sh_rp1, &b_sh_rp2,
                &b_sh_rcount, &b_sh_rbucket, &b_sh_sorted, &b_sh_boff, &b_sh_bdrop};
        for (int w = 0; w < MAX_LANES - 1; ++w)
            for (DevBuf *b : {&b_taskkey_x[w], &b_trace_x[w], &b_bnd_x[w], &b_runoff_x[w], &b_nruns64_x[w], &b_scan_tmp_x[w]})
                v.push_back(b);
        for (DevBuf &b : b_ed) v.push_back(&b);
        v.push_back(&b_h_dist);
        v.push_back(&b_ed_eqg);
        return v;
    }
    void *fused_io = nullptr;         // mapped pinned memory of the fused tiny-call path (inputs | outputs)
    void *fused_io_dev = nullptr;     // its device address
    bool disable_fused = false;       // tiny calls through the regular path (tests / A-B measurements)
    void *pinned = nullptr;           // small pinned staging for scalar read-backs
    size_t pinned_cap = 0;
};

enum { SLOT_DP = 0, SLOT_TB = 1, SLOT_MIN = 2, SLOT_OTHER = 3 };

// Device-side bounds assertions (compute-sanitizer is not available on the target pool).  The debug
// build (-DISOF_CHECK, libisonform_b200_dbg.so) records violated checks as bits of a device word that
// isof_debug_violations() reads back; nothing traps, so a bad index cannot wedge the GPU.  Compiled out
// of the release library.
enum {
    CHK_TRACE_STORE = 0, CHK_BND = 1, CHK_SEQ_LOAD = 2, CHK_TRACE_LOAD = 3, CHK_TAIL_OVERLAP = 4, CHK_MIN_SMEM = 5,
    CHK_MIN_OUT = 6, CHK_SPAN_REC = 7, CHK_CIGAR_COPY = 8
};
#ifdef __CUDACC__
static __device__ unsigned int g_isof_violations;     // one copy per translation unit (no -rdc); see isof_viol_*()
static inline unsigned isof_read_violations_tu()
{
    unsigned v = 0, z = 0;
    cudaMemcpyFromSymbol(&v, g_isof_violations, sizeof(v));
    cudaMemcpyToSymbol(g_isof_violations, &z, sizeof(z));
    return v;
}
#ifdef ISOF_CHECK
#define ISOF_ASSERT(cond, code)                                                                    \
    do {                                                                                           \
        if (!(cond)) atomicOr(&g_isof_violations, 1u << (code));                                   \
    } while (0)
#else
#define ISOF_ASSERT(cond, code) ((void)0)
#endif
#endif

static inline int isof_fail(isof_ctx *ctx, int code, const char *fmt, ...)
{
    if (ctx) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(ctx->err, sizeof(ctx->err), fmt, ap);
        va_end(ap);
    }
    return code;
}

#define CUDA_TRY(ctx, call)                                                                        \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return isof_fail((ctx), ISOF_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__,      \
                             __LINE__, cudaGetErrorString(e__));                                   \
    } while (0)

static inline int dev_reserve(isof_ctx *ctx, DevBuf &b, size_t bytes)
{
    if (bytes <= b.cap) return ISOF_OK;
    if (b.p) { CUDA_TRY(ctx, cudaFree(b.p)); b.p = nullptr; b.cap = 0; }
    size_t want = bytes + std::min<size_t>(bytes / 8, (size_t)64 << 20) + 256;      // growth slack, capped: trace buffers are tens of GB
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        want = bytes;
        e = cudaMalloc(&b.p, want);
    }
    if (e != cudaSuccess) {
        b.p = nullptr;
        return isof_fail(ctx, ISOF_ERR_CUDA, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    }
    b.cap = want;
    return ISOF_OK;
}

#define TRY(x)                                                                                     \
    do {                                                                                           \
        int rc__ = (x);                                                                            \
        if (rc__ != ISOF_OK) return rc__;                                                          \
    } while (0)

// Launch bookkeeping: counts the launch and, when profiling, brackets it with events.
struct LaunchTimer {
    isof_ctx *ctx;
    int slot;
    cudaEvent_t a = nullptr, b = nullptr;
    cudaStream_t st;
    LaunchTimer(isof_ctx *c, int s, cudaStream_t stream = nullptr) : ctx(c), slot(s), st(stream ? stream : c->stream)
    {
        ctx->prof.launches++;
        if (slot == SLOT_DP) ctx->prof.dp_launches++;
        if (ctx->profiling) {
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            cudaEventRecord(a, st);
        }
    }
    ~LaunchTimer()
    {
        if (ctx->profiling) {
            cudaEventRecord(b, st);
            ctx->timed.push_back({a, b, slot});
        }
    }
};

unsigned isof_viol_align();
unsigned isof_viol_minimizers();
unsigned isof_viol_spans();
unsigned isof_viol_myers();

// internal entry points implemented in the .cu files
int isof_minimizers_core(isof_ctx *ctx, const uint8_t *d_bases, const int64_t *d_off, int32_t R, int64_t n_bases,
                         int32_t k, int32_t w_size, int32_t *d_out_pos, int64_t out_cap, int64_t *d_out_off,
                         int64_t *h_n_out);
int isof_spans_core(isof_ctx *ctx, const uint8_t *d_bases, const int64_t *d_read_off, int32_t R, const int32_t *d_min_pos,
                    const int64_t *d_min_off, int64_t M, int32_t k, int32_t x_low, int32_t x_high, int32_t delta_len,
                    int64_t rec_cap, int32_t *d_rec_read, int32_t *d_rec_p1, int32_t *d_rec_p2, int32_t *d_rec_count,
                    int32_t *d_rec_bucket, int32_t *d_sorted_rec, int32_t *d_bucket_off, uint8_t *d_bucket_dropped,
                    int64_t *h_n_rec, int64_t *h_n_buckets, const int32_t *d_batch_read_off = nullptr, int32_t n_batches = 1);
int isof_sg_fused_small(isof_ctx *ctx, const uint8_t *seq_cat, const int64_t *seq_off, int32_t n_seqs, const int32_t *pair_a,
                        const int32_t *pair_b, int64_t P, int32_t match, int32_t mismatch, int32_t gap_open, int32_t gap_ext,
                        int32_t delta_len, int32_t *score, int32_t *end_q, int32_t *end_r, uint32_t *cigar_runs,
                        int64_t cigar_cap, int64_t *cigar_off, int32_t *features);
int isof_sg_align_core(isof_ctx *ctx, const uint8_t *d_cat, const int64_t *d_off, int32_t n_seqs, int64_t n_bases,
                       const int32_t *d_pa, const int32_t *d_pb, int64_t P, int32_t match, int32_t mismatch,
                       int32_t gap_open, int32_t gap_ext, int32_t *d_score, int32_t *d_endq, int32_t *d_endr,
                       uint32_t *d_cigar, int64_t cigar_cap, int64_t *d_cigar_off, int32_t *d_features,
                       int32_t delta_len, int64_t *h_total_runs);
int isof_edit_distance_core(isof_ctx *ctx, const uint8_t *d_cat, const int64_t *d_off, int32_t n_seqs, int64_t n_bases,
                            const int32_t *d_pa, const int32_t *d_pb, int64_t P, int32_t k, int32_t *d_dist,
                            uint32_t *d_cigar, int64_t cigar_cap, int64_t *d_cigar_off, int64_t *h_total_runs);
