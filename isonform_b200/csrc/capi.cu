// capi.cu -- the extern "C" boundary declared in include/isonform_b200.h.
#include "common.cuh"
#include <algorithm>

namespace {

int flush_timers(isof_ctx *ctx)
{
    if (ctx->timed.empty()) return ISOF_OK;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto &t : ctx->timed) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, t.a, t.b);
        switch (t.slot) {
            case SLOT_DP: ctx->prof.dp_ms += ms; break;
            case SLOT_TB: ctx->prof.traceback_ms += ms; break;
            case SLOT_MIN: ctx->prof.minimizer_ms += ms; break;
            default: ctx->prof.other_ms += ms; break;
        }
        cudaEventDestroy(t.a);
        cudaEventDestroy(t.b);
    }
    ctx->timed.clear();
    return ISOF_OK;
}

int h2d(isof_ctx *ctx, DevBuf &b, const void *src, size_t bytes)
{
    TRY(dev_reserve(ctx, b, bytes + 16));
    if (bytes) CUDA_TRY(ctx, cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    ctx->prof.h2d_bytes += (int64_t)bytes;
    return ISOF_OK;
}

int d2h(isof_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    if (bytes && dst) CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (dst) ctx->prof.d2h_bytes += (int64_t)bytes;
    return ISOF_OK;
}

__global__ void __launch_bounds__(256) alu_peak_kernel(int *out, int iters, int c1, int c2)
{
    int a[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) a[u] = threadIdx.x * 17 + u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) a[u] = __viaddmax_s32(a[u], c1, c2 + u);
    }
    int s = 0;
#pragma unroll
    for (int u = 0; u < 16; ++u) s ^= a[u];
    if (s == 0x7fffffff) out[0] = s;          // keeps the chains alive, practically never taken
}

}  // namespace

extern "C" {

int isof_int32_peak(isof_ctx *ctx, double *lane_ops_per_s)
{
    if (!ctx || !lane_ops_per_s) return ISOF_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    TRY(dev_reserve(ctx, ctx->b_misc, 64));
    const int blocks = ctx->sm_count * 8, iters = 4096;
    cudaEvent_t a, b;
    CUDA_TRY(ctx, cudaEventCreate(&a));
    CUDA_TRY(ctx, cudaEventCreate(&b));
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        CUDA_TRY(ctx, cudaEventRecord(a, ctx->stream));
        alu_peak_kernel<<<blocks, 256, 0, ctx->stream>>>((int *)ctx->b_misc.p, iters, -3, 5);
        CUDA_TRY(ctx, cudaEventRecord(b, ctx->stream));
        CUDA_TRY(ctx, cudaEventSynchronize(b));
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        double ops = (double)blocks * 256.0 * iters * 16.0 / (ms * 1e-3);
        if (rep > 0 && ops > best) best = ops;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    ctx->prof.launches += 4;
    *lane_ops_per_s = best;
    return ISOF_OK;
}


int isof_version(void) { return ISOF_VERSION; }

int isof_init(int device, isof_ctx **out)
{
    if (!out) return ISOF_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
        cudaGetLastError();
        return ISOF_ERR_CUDA;                 // no CPU fallback: the caller must fail loudly
    }
    isof_ctx *ctx = new isof_ctx();
    ctx->device = device;
    memset(&ctx->prof, 0, sizeof(ctx->prof));
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return ISOF_ERR_CUDA; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return ISOF_ERR_CUDA; }
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return ISOF_ERR_CUDA; }
    ctx->stream = ctx->own_stream;
    for (auto &as : ctx->aux_stream)
        if (cudaStreamCreateWithFlags(&as, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return ISOF_ERR_CUDA; }
    ctx->pinned_cap = 4096;
    if (cudaMallocHost(&ctx->pinned, ctx->pinned_cap) != cudaSuccess) { delete ctx; return ISOF_ERR_CUDA; }
    *out = ctx;
    return ISOF_OK;
}

int isof_destroy(isof_ctx *ctx)
{
    if (!ctx) return ISOF_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    flush_timers(ctx);
    for (DevBuf *b : ctx->all_bufs()) if (b->p) cudaFree(b->p);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->fused_io) cudaFreeHost(ctx->fused_io);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    for (auto &as : ctx->aux_stream) if (as) cudaStreamDestroy(as);
    delete ctx;
    return ISOF_OK;
}

int isof_release_scratch(isof_ctx *ctx)
{
    if (!ctx) return ISOF_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto &as : ctx->aux_stream) if (as) CUDA_TRY(ctx, cudaStreamSynchronize(as));
    for (DevBuf *b : ctx->all_bufs()) {
        if (b->p) CUDA_TRY(ctx, cudaFree(b->p));
        b->p = nullptr;
        b->cap = 0;
    }
    ctx->trace_free_at_first_use = 0;          // the automatic trace budget is taken again from what is free at the next call
    return ISOF_OK;
}

const char *isof_last_error(const isof_ctx *ctx) { return ctx ? ctx->err : "no context"; }

int isof_set_stream(isof_ctx *ctx, void *cuda_stream)
{
    if (!ctx) return ISOF_ERR_ARG;
    cudaStreamSynchronize(ctx->stream);
    flush_timers(ctx);
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return ISOF_OK;
}

int isof_set_trace_budget(isof_ctx *ctx, size_t bytes)
{
    if (!ctx) return ISOF_ERR_ARG;
    ctx->trace_budget = bytes;
    return ISOF_OK;
}

int isof_set_packed(isof_ctx *ctx, int on)
{
    if (!ctx) return ISOF_ERR_ARG;
    ctx->disable_packed = (on == 0);
    return ISOF_OK;
}

int isof_debug_set_span_hash_mask(isof_ctx *ctx, uint64_t mask)
{
    if (!ctx) return ISOF_ERR_ARG;
    ctx->span_hash_mask = mask & 0x7fffffffffffffffULL;
    return ISOF_OK;
}

int isof_set_traceback_warp(isof_ctx *ctx, int mode)
{
    if (!ctx || mode < 0 || mode > 2) return ISOF_ERR_ARG;
    ctx->tb_warp = mode;
    ctx->ed_tb_warp = mode;            // kernel (b') has the same two walks (myers.cu)
    return ISOF_OK;
}

int isof_set_band(isof_ctx *ctx, int half_width)
{
    if (!ctx || half_width < 0) return ISOF_ERR_ARG;
    ctx->band_w0 = half_width;
    return ISOF_OK;
}

int isof_debug_fused_phases(isof_ctx *ctx, int64_t *cycles40)
{
    if (!ctx || !cycles40) return ISOF_ERR_ARG;
    if (!ctx->fused_io) { memset(cycles40, 0, 448); return ISOF_OK; }
    memcpy(cycles40, (const uint8_t *)ctx->fused_io + 2 * 512 * 1024 - 512, 448);
    return ISOF_OK;
}

int isof_set_fused_small(isof_ctx *ctx, int on)
{
    if (!ctx) return ISOF_ERR_ARG;
    ctx->disable_fused = (on == 0);
    return ISOF_OK;
}

int isof_set_short(isof_ctx *ctx, int on)
{
    if (!ctx) return ISOF_ERR_ARG;
    ctx->short_mode = on < 0 ? 0 : (on > 2 ? 2 : on);
    return ISOF_OK;
}

int isof_set_tiebreak(isof_ctx *ctx, int h_e_before_f, int gap_open_on_tie, int end_col_first, int swap_id)
{
    if (!ctx) return ISOF_ERR_ARG;
    ctx->tb_h_e_before_f = h_e_before_f != 0;
    ctx->tb_gap_open_on_tie = gap_open_on_tie != 0;
    ctx->tb_end_col_first = end_col_first != 0;
    ctx->tb_swap_id = swap_id != 0;
    return ISOF_OK;
}

namespace { __global__ void selftest_kernel() { ISOF_ASSERT(threadIdx.x > 1000u, 30); } }

int isof_debug_selftest(isof_ctx *ctx)
{
    if (!ctx) return ISOF_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    selftest_kernel<<<1, 32, 0, ctx->stream>>>();
    CUDA_TRY(ctx, cudaGetLastError());
    return ISOF_OK;
}

int isof_debug_violations(isof_ctx *ctx, unsigned int *mask)
{
    if (!ctx || !mask) return ISOF_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaDeviceSynchronize());
    *mask = isof_viol_align() | isof_viol_minimizers() | isof_viol_spans() | isof_viol_myers() | isof_read_violations_tu();
#ifdef ISOF_CHECK
    *mask |= 0x80000000u;             // this library was built with the assertions compiled in
#endif
    return ISOF_OK;
}

int isof_set_overlap(isof_ctx *ctx, int on)
{
    if (!ctx) return ISOF_ERR_ARG;
    const bool want = (on != 0);
    if (want != ctx->overlap) {
        // the two modes size their trace scratch differently (one buffer of the whole budget vs NLANES buffers of a
        // share each): drop the buffers so that the next call re-reserves them and the total stays within the budget
        CUDA_TRY(ctx, cudaSetDevice(ctx->device));
        CUDA_TRY(ctx, cudaDeviceSynchronize());
        std::vector<DevBuf *> bufs = {&ctx->b_trace};
        for (DevBuf &b : ctx->b_trace_x) bufs.push_back(&b);
        for (DevBuf *b : bufs)
            if (b->p) { cudaFree(b->p); b->p = nullptr; b->cap = 0; }
    }
    ctx->overlap = want;
    return ISOF_OK;
}

int isof_profile_enable(isof_ctx *ctx, int on)
{
    if (!ctx) return ISOF_ERR_ARG;
    flush_timers(ctx);
    ctx->profiling = on != 0;
    return ISOF_OK;
}

int isof_profile_reset(isof_ctx *ctx)
{
    if (!ctx) return ISOF_ERR_ARG;
    flush_timers(ctx);
    memset(&ctx->prof, 0, sizeof(ctx->prof));
    return ISOF_OK;
}

int isof_profile_get(isof_ctx *ctx, isof_profile *out)
{
    if (!ctx || !out) return ISOF_ERR_ARG;
    TRY(flush_timers(ctx));
    *out = ctx->prof;
    return ISOF_OK;
}

// ------------------------------------------------------------------------------------------ (a)
int isof_minimizers_batch_dev(isof_ctx *ctx, const uint8_t *bases, const int64_t *read_off, int32_t n_reads,
                              int64_t n_bases, int32_t k, int32_t w_size, int32_t *out_pos, int64_t out_cap,
                              int64_t *out_off, int64_t *n_out)
{
    if (!ctx || !n_out) return ISOF_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return isof_minimizers_core(ctx, bases, read_off, n_reads, n_bases, k, w_size, out_pos, out_cap, out_off, n_out);
}

int isof_minimizers_batch(isof_ctx *ctx, const uint8_t *bases, const int64_t *read_off, int32_t n_reads, int32_t k,
                          int32_t w_size, int32_t *out_pos, int64_t out_cap, int64_t *out_off, int64_t *n_out)
{
    if (!ctx || !read_off || !out_off || !n_out || n_reads < 0) return isof_fail(ctx, ISOF_ERR_ARG, "null argument");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int64_t n_bases = read_off[n_reads];
    if (read_off[0] != 0 || n_bases < 0) return isof_fail(ctx, ISOF_ERR_ARG, "read_off must start at 0 and be non-decreasing");
    TRY(h2d(ctx, ctx->b_h_bases, bases, (size_t)n_bases));
    TRY(h2d(ctx, ctx->b_h_off, read_off, sizeof(int64_t) * (size_t)(n_reads + 1)));
    // device outputs: worst case one minimizer per base + one per read
    const int64_t cap_dev = std::max<int64_t>(std::min<int64_t>(out_cap, n_bases + n_reads), 1);
    TRY(dev_reserve(ctx, ctx->b_h_cig, sizeof(int32_t) * (size_t)cap_dev));
    TRY(dev_reserve(ctx, ctx->b_h_cigoff, sizeof(int64_t) * (size_t)(n_reads + 1)));
    int rc = isof_minimizers_core(ctx, (const uint8_t *)ctx->b_h_bases.p, (const int64_t *)ctx->b_h_off.p, n_reads, n_bases,
                                  k, w_size, (int32_t *)ctx->b_h_cig.p, cap_dev, (int64_t *)ctx->b_h_cigoff.p, n_out);
    if (rc != ISOF_OK && rc != ISOF_ERR_CAPACITY) return rc;
    TRY(d2h(ctx, out_off, ctx->b_h_cigoff.p, sizeof(int64_t) * (size_t)(n_reads + 1)));
    if (rc == ISOF_OK) TRY(d2h(ctx, out_pos, ctx->b_h_cig.p, sizeof(int32_t) * (size_t)*n_out));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return rc;
}

// ------------------------------------------------------------------------------------------ (c)
int isof_spans_batch_dev(isof_ctx *ctx, const uint8_t *bases, const int64_t *read_off, int32_t n_reads, const int32_t *min_pos,
                         const int64_t *min_off, int64_t n_min, int32_t k, int32_t x_low, int32_t x_high, int32_t delta_len,
                         int64_t rec_cap, int32_t *rec_read, int32_t *rec_p1, int32_t *rec_p2, int32_t *rec_count,
                         int32_t *rec_bucket, int32_t *sorted_rec, int32_t *bucket_off, uint8_t *bucket_dropped,
                         int64_t *n_rec, int64_t *n_buckets)
{
    if (!ctx || !n_rec || !n_buckets) return ISOF_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return isof_spans_core(ctx, bases, read_off, n_reads, min_pos, min_off, n_min, k, x_low, x_high, delta_len, rec_cap, rec_read,
                           rec_p1, rec_p2, rec_count, rec_bucket, sorted_rec, bucket_off, bucket_dropped, n_rec, n_buckets);
}

int isof_spans_batch(isof_ctx *ctx, const uint8_t *bases, const int64_t *read_off, int32_t n_reads, const int32_t *min_pos,
                     const int64_t *min_off, int32_t k, int32_t x_low, int32_t x_high, int32_t delta_len, int64_t rec_cap,
                     int32_t *rec_read, int32_t *rec_p1, int32_t *rec_p2, int32_t *rec_count, int32_t *rec_bucket,
                     int32_t *sorted_rec, int32_t *bucket_off, uint8_t *bucket_dropped, int64_t *n_rec, int64_t *n_buckets)
{
    return isof_spans_multi(ctx, bases, read_off, n_reads, nullptr, 1, min_pos, min_off, k, x_low, x_high, delta_len, rec_cap,
                            rec_read, rec_p1, rec_p2, rec_count, rec_bucket, sorted_rec, bucket_off, bucket_dropped, n_rec,
                            n_buckets);
}

int isof_spans_multi(isof_ctx *ctx, const uint8_t *bases, const int64_t *read_off, int32_t n_reads,
                     const int32_t *batch_read_off, int32_t n_batches, const int32_t *min_pos, const int64_t *min_off, int32_t k,
                     int32_t x_low, int32_t x_high, int32_t delta_len, int64_t rec_cap, int32_t *rec_read, int32_t *rec_p1,
                     int32_t *rec_p2, int32_t *rec_count, int32_t *rec_bucket, int32_t *sorted_rec, int32_t *bucket_off,
                     uint8_t *bucket_dropped, int64_t *n_rec, int64_t *n_buckets)
{
    if (!ctx || !read_off || !min_off || !n_rec || !n_buckets || n_reads < 0) return isof_fail(ctx, ISOF_ERR_ARG, "null argument");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int64_t n_bases = read_off[n_reads], M = min_off[n_reads];
    TRY(h2d(ctx, ctx->b_h_bases, bases, (size_t)n_bases));
    TRY(h2d(ctx, ctx->b_h_off, read_off, sizeof(int64_t) * (size_t)(n_reads + 1)));
    TRY(h2d(ctx, ctx->b_sh_minpos, min_pos, sizeof(int32_t) * (size_t)M));
    TRY(h2d(ctx, ctx->b_sh_minoff, min_off, sizeof(int64_t) * (size_t)(n_reads + 1)));
    const int32_t *d_batch_off = nullptr;
    if (batch_read_off && n_batches > 1) {
        if (batch_read_off[0] != 0 || batch_read_off[n_batches] != n_reads)
            return isof_fail(ctx, ISOF_ERR_ARG, "batch_read_off must cover reads 0..n_reads");
        TRY(h2d(ctx, ctx->b_sh_batchoff, batch_read_off, sizeof(int32_t) * (size_t)(n_batches + 1)));
        d_batch_off = (const int32_t *)ctx->b_sh_batchoff.p;
    }
    const size_t cap = (size_t)std::max<int64_t>(rec_cap, 1);
    TRY(dev_reserve(ctx, ctx->b_sh_rread, 4 * cap));
    TRY(dev_reserve(ctx, ctx->b_sh_rp1, 4 * cap));
    TRY(dev_reserve(ctx, ctx->b_sh_rp2, 4 * cap));
    TRY(dev_reserve(ctx, ctx->b_sh_rcount, 4 * cap));
    TRY(dev_reserve(ctx, ctx->b_sh_rbucket, 4 * cap));
    TRY(dev_reserve(ctx, ctx->b_sh_sorted, 4 * cap));
    TRY(dev_reserve(ctx, ctx->b_sh_boff, 4 * (cap + 1)));
    TRY(dev_reserve(ctx, ctx->b_sh_bdrop, cap));
    int rc = isof_spans_core(ctx, (const uint8_t *)ctx->b_h_bases.p, (const int64_t *)ctx->b_h_off.p, n_reads,
                             (const int32_t *)ctx->b_sh_minpos.p, (const int64_t *)ctx->b_sh_minoff.p, M, k, x_low, x_high,
                             delta_len, rec_cap, (int32_t *)ctx->b_sh_rread.p, (int32_t *)ctx->b_sh_rp1.p,
                             (int32_t *)ctx->b_sh_rp2.p, (int32_t *)ctx->b_sh_rcount.p, (int32_t *)ctx->b_sh_rbucket.p,
                             (int32_t *)ctx->b_sh_sorted.p, (int32_t *)ctx->b_sh_boff.p, (uint8_t *)ctx->b_sh_bdrop.p, n_rec,
                             n_buckets, d_batch_off, n_batches);
    if (rc != ISOF_OK) return rc;
    const size_t G = (size_t)*n_rec, B = (size_t)*n_buckets;
    TRY(d2h(ctx, rec_read, ctx->b_sh_rread.p, 4 * G));
    TRY(d2h(ctx, rec_p1, ctx->b_sh_rp1.p, 4 * G));
    TRY(d2h(ctx, rec_p2, ctx->b_sh_rp2.p, 4 * G));
    TRY(d2h(ctx, rec_count, ctx->b_sh_rcount.p, 4 * G));
    TRY(d2h(ctx, rec_bucket, ctx->b_sh_rbucket.p, 4 * G));
    TRY(d2h(ctx, sorted_rec, ctx->b_sh_sorted.p, 4 * G));
    if (G) TRY(d2h(ctx, bucket_off, ctx->b_sh_boff.p, 4 * (B + 1)));
    TRY(d2h(ctx, bucket_dropped, ctx->b_sh_bdrop.p, B));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return ISOF_OK;
}

// ------------------------------------------------------------------------------------------ (b)
int isof_sg_align_pairs_dev(isof_ctx *ctx, const uint8_t *seq_cat, const int64_t *seq_off, int32_t n_seqs, int64_t n_bases,
                            const int32_t *pair_a, const int32_t *pair_b, int64_t n_pairs, int32_t match, int32_t mismatch,
                            int32_t gap_open, int32_t gap_ext, int32_t *score, int32_t *end_q, int32_t *end_r,
                            uint32_t *cigar_runs, int64_t cigar_cap, int64_t *cigar_off, int32_t *features,
                            int32_t delta_len, int64_t *n_runs_total)
{
    if (!ctx || !score || !end_q || !end_r) return isof_fail(ctx, ISOF_ERR_ARG, "null argument");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return isof_sg_align_core(ctx, seq_cat, seq_off, n_seqs, n_bases, pair_a, pair_b, n_pairs, match, mismatch, gap_open,
                              gap_ext, score, end_q, end_r, cigar_runs, cigar_cap, cigar_off, features, delta_len,
                              n_runs_total);
}

static int align_host(isof_ctx *ctx, const uint8_t *seq_cat, const int64_t *seq_off, int32_t n_seqs, const int32_t *pair_a,
                      const int32_t *pair_b, int64_t P, int32_t match, int32_t mismatch, int32_t gap_open, int32_t gap_ext,
                      int32_t delta_len, int32_t *score, int32_t *end_q, int32_t *end_r, uint32_t *cigar_runs,
                      int64_t cigar_cap, int64_t *cigar_off, int32_t *features)
{
    if (!ctx || !seq_off || n_seqs < 0 || P < 0) return isof_fail(ctx, ISOF_ERR_ARG, "null argument");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int64_t n_bases = seq_off[n_seqs];
    if (seq_off[0] != 0 || n_bases < 0) return isof_fail(ctx, ISOF_ERR_ARG, "seq_off must start at 0");
    if (P == 0) { if (cigar_off) cigar_off[0] = 0; return ISOF_OK; }
    if (!pair_a || !pair_b) return isof_fail(ctx, ISOF_ERR_ARG, "null pair list");
    if (!ctx->disable_fused) {
        // tiny calls (the per-pair seams): one fused launch, trace in shared memory, mapped I/O
        const int frc = isof_sg_fused_small(ctx, seq_cat, seq_off, n_seqs, pair_a, pair_b, P, match, mismatch, gap_open, gap_ext,
                                            delta_len, score, end_q, end_r, cigar_runs, cigar_cap, cigar_off, features);
        if (frc != 1) return frc;
    }
    TRY(h2d(ctx, ctx->b_h_bases, seq_cat, (size_t)n_bases));
    TRY(h2d(ctx, ctx->b_h_off, seq_off, sizeof(int64_t) * (size_t)(n_seqs + 1)));
    TRY(h2d(ctx, ctx->b_h_pa, pair_a, sizeof(int32_t) * (size_t)P));
    TRY(h2d(ctx, ctx->b_h_pb, pair_b, sizeof(int32_t) * (size_t)P));
    TRY(dev_reserve(ctx, ctx->b_h_score, sizeof(int32_t) * (size_t)P));
    TRY(dev_reserve(ctx, ctx->b_h_endq, sizeof(int32_t) * (size_t)P));
    TRY(dev_reserve(ctx, ctx->b_h_endr, sizeof(int32_t) * (size_t)P));
    TRY(dev_reserve(ctx, ctx->b_h_cigoff, sizeof(int64_t) * (size_t)(P + 1)));
    if (cigar_runs && cigar_cap > 0) TRY(dev_reserve(ctx, ctx->b_h_cig, sizeof(uint32_t) * (size_t)cigar_cap));
    if (features) TRY(dev_reserve(ctx, ctx->b_h_feat, sizeof(int32_t) * (size_t)P * ISOF_N_FEATURES));
    int64_t total = 0;
    uint32_t *d_cig = (cigar_runs && cigar_cap > 0) ? (uint32_t *)ctx->b_h_cig.p : nullptr;
    int rc = isof_sg_align_core(ctx, (const uint8_t *)ctx->b_h_bases.p, (const int64_t *)ctx->b_h_off.p, n_seqs, n_bases,
                                (const int32_t *)ctx->b_h_pa.p, (const int32_t *)ctx->b_h_pb.p, P, match, mismatch, gap_open,
                                gap_ext, (int32_t *)ctx->b_h_score.p, (int32_t *)ctx->b_h_endq.p, (int32_t *)ctx->b_h_endr.p,
                                d_cig, cigar_cap, (int64_t *)ctx->b_h_cigoff.p,
                                features ? (int32_t *)ctx->b_h_feat.p : nullptr, delta_len, &total);
    if (rc != ISOF_OK && rc != ISOF_ERR_CAPACITY) return rc;
    TRY(d2h(ctx, score, ctx->b_h_score.p, sizeof(int32_t) * (size_t)P));
    TRY(d2h(ctx, end_q, ctx->b_h_endq.p, sizeof(int32_t) * (size_t)P));
    TRY(d2h(ctx, end_r, ctx->b_h_endr.p, sizeof(int32_t) * (size_t)P));
    TRY(d2h(ctx, cigar_off, ctx->b_h_cigoff.p, sizeof(int64_t) * (size_t)(P + 1)));
    if (features) TRY(d2h(ctx, features, ctx->b_h_feat.p, sizeof(int32_t) * (size_t)P * ISOF_N_FEATURES));
    if (d_cig && rc == ISOF_OK) TRY(d2h(ctx, cigar_runs, d_cig, sizeof(uint32_t) * (size_t)total));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return rc;
}

int isof_sg_align_pairs(isof_ctx *ctx, const uint8_t *seq_cat, const int64_t *seq_off, int32_t n_seqs, const int32_t *pair_a,
                        const int32_t *pair_b, int64_t n_pairs, int32_t match, int32_t mismatch, int32_t gap_open,
                        int32_t gap_ext, int32_t *score, int32_t *end_q, int32_t *end_r, uint32_t *cigar_runs,
                        int64_t cigar_cap, int64_t *cigar_off)
{
    if (!score || !end_q || !end_r) return isof_fail(ctx, ISOF_ERR_ARG, "null output");
    return align_host(ctx, seq_cat, seq_off, n_seqs, pair_a, pair_b, n_pairs, match, mismatch, gap_open, gap_ext, 0, score,
                      end_q, end_r, cigar_runs, cigar_cap, cigar_off, nullptr);
}

int isof_sg_features_pairs(isof_ctx *ctx, const uint8_t *seq_cat, const int64_t *seq_off, int32_t n_seqs,
                           const int32_t *pair_a, const int32_t *pair_b, int64_t n_pairs, int32_t match, int32_t mismatch,
                           int32_t gap_open, int32_t gap_ext, int32_t delta_len, int32_t *score, int32_t *features,
                           uint32_t *cigar_runs, int64_t cigar_cap, int64_t *cigar_off)
{
    if (!score || !features) return isof_fail(ctx, ISOF_ERR_ARG, "null output");
    return align_host(ctx, seq_cat, seq_off, n_seqs, pair_a, pair_b, n_pairs, match, mismatch, gap_open, gap_ext, delta_len,
                      score, nullptr, nullptr, cigar_runs, cigar_cap, cigar_off, features);
}

int isof_sg_align_batch(isof_ctx *ctx, const uint8_t *s1_cat, const int64_t *s1_off, const uint8_t *s2_cat,
                        const int64_t *s2_off, int32_t n_pairs, int32_t match, int32_t mismatch, int32_t gap_open,
                        int32_t gap_ext, int32_t *score, int32_t *end_q, int32_t *end_r, uint32_t *cigar_runs,
                        int64_t cigar_cap, int64_t *cigar_off)
{
    if (!ctx || !s1_off || !s2_off || n_pairs < 0) return isof_fail(ctx, ISOF_ERR_ARG, "null argument");
    // build a pool: [s1 sequences..., s2 sequences...]
    const int64_t n1 = s1_off[n_pairs], n2 = s2_off[n_pairs];
    std::vector<uint8_t> cat((size_t)(n1 + n2));
    if (n1) memcpy(cat.data(), s1_cat, (size_t)n1);
    if (n2) memcpy(cat.data() + n1, s2_cat, (size_t)n2);
    std::vector<int64_t> off((size_t)2 * n_pairs + 1);
    for (int32_t p = 0; p <= n_pairs; ++p) off[p] = s1_off[p];
    for (int32_t p = 0; p <= n_pairs; ++p) off[(size_t)n_pairs + p] = n1 + s2_off[p];
    std::vector<int32_t> pa((size_t)n_pairs), pb((size_t)n_pairs);
    for (int32_t p = 0; p < n_pairs; ++p) { pa[p] = p; pb[p] = n_pairs + p; }
    return isof_sg_align_pairs(ctx, cat.data(), off.data(), 2 * n_pairs, pa.data(), pb.data(), n_pairs, match, mismatch,
                               gap_open, gap_ext, score, end_q, end_r, cigar_runs, cigar_cap, cigar_off);
}

// ---- (b') unit-cost global edit distance (myers.cu)
int isof_edit_distance_pairs_dev(isof_ctx *ctx, const uint8_t *seq_cat, const int64_t *seq_off, int32_t n_seqs, int64_t n_bases,
                                 const int32_t *pair_a, const int32_t *pair_b, int64_t n_pairs, int32_t k, int32_t *dist,
                                 uint32_t *cigar_runs, int64_t cigar_cap, int64_t *cigar_off, int64_t *n_runs_total)
{
    if (!ctx || !dist || !seq_off || (n_pairs > 0 && (!pair_a || !pair_b))) return isof_fail(ctx, ISOF_ERR_ARG, "null argument");
    if (cigar_runs && !cigar_off) return isof_fail(ctx, ISOF_ERR_ARG, "cigar_runs without cigar_off");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return isof_edit_distance_core(ctx, seq_cat, seq_off, n_seqs, n_bases, pair_a, pair_b, n_pairs, k, dist, cigar_runs, cigar_cap,
                                   cigar_off, n_runs_total);
}

int isof_edit_distance_pairs(isof_ctx *ctx, const uint8_t *seq_cat, const int64_t *seq_off, int32_t n_seqs, const int32_t *pair_a,
                             const int32_t *pair_b, int64_t P, int32_t k, int32_t *dist, uint32_t *cigar_runs, int64_t cigar_cap,
                             int64_t *cigar_off)
{
    if (!ctx || !seq_off || n_seqs < 0 || P < 0 || !dist) return isof_fail(ctx, ISOF_ERR_ARG, "null argument");
    if (cigar_runs && !cigar_off) return isof_fail(ctx, ISOF_ERR_ARG, "cigar_runs without cigar_off");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int64_t n_bases = seq_off[n_seqs];
    if (seq_off[0] != 0 || n_bases < 0) return isof_fail(ctx, ISOF_ERR_ARG, "seq_off must start at 0");
    if (P == 0) { if (cigar_off) cigar_off[0] = 0; return ISOF_OK; }
    if (!pair_a || !pair_b) return isof_fail(ctx, ISOF_ERR_ARG, "null pair list");
    TRY(h2d(ctx, ctx->b_h_bases, seq_cat, (size_t)n_bases));
    TRY(h2d(ctx, ctx->b_h_off, seq_off, sizeof(int64_t) * (size_t)(n_seqs + 1)));
    TRY(h2d(ctx, ctx->b_h_pa, pair_a, sizeof(int32_t) * (size_t)P));
    TRY(h2d(ctx, ctx->b_h_pb, pair_b, sizeof(int32_t) * (size_t)P));
    TRY(dev_reserve(ctx, ctx->b_h_dist, sizeof(int32_t) * (size_t)P));
    if (cigar_off) TRY(dev_reserve(ctx, ctx->b_h_cigoff, sizeof(int64_t) * (size_t)(P + 1)));
    uint32_t *d_cig = nullptr;
    if (cigar_runs && cigar_cap > 0) {
        TRY(dev_reserve(ctx, ctx->b_h_cig, sizeof(uint32_t) * (size_t)cigar_cap));
        d_cig = (uint32_t *)ctx->b_h_cig.p;
    }
    int64_t total = 0;
    const int rc = isof_edit_distance_core(ctx, (const uint8_t *)ctx->b_h_bases.p, (const int64_t *)ctx->b_h_off.p, n_seqs, n_bases,
                                           (const int32_t *)ctx->b_h_pa.p, (const int32_t *)ctx->b_h_pb.p, P, k,
                                           (int32_t *)ctx->b_h_dist.p, d_cig, d_cig ? cigar_cap : 0,
                                           cigar_off ? (int64_t *)ctx->b_h_cigoff.p : nullptr, &total);
    if (rc != ISOF_OK && rc != ISOF_ERR_CAPACITY) return rc;
    int rc2 = rc;
    if (cigar_runs && !d_cig && total > 0) rc2 = isof_fail(ctx, ISOF_ERR_CAPACITY, "cigar buffer too small: need %lld runs", (long long)total);
    TRY(d2h(ctx, dist, ctx->b_h_dist.p, sizeof(int32_t) * (size_t)P));
    if (cigar_off) TRY(d2h(ctx, cigar_off, ctx->b_h_cigoff.p, sizeof(int64_t) * (size_t)(P + 1)));
    if (d_cig && rc == ISOF_OK) TRY(d2h(ctx, cigar_runs, d_cig, sizeof(uint32_t) * (size_t)total));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return rc2;
}

}  // extern "C"
