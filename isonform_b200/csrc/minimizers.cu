// minimizers.cu -- kernel (a): batched minimizer extraction, bit-exact with
// get_kmer_minimizers (reference main:81-110) under PYTHONHASHSEED=0 / CPython >= 3.11.
//
// One warp per read.  The read is consumed in chunks of CHUNK k-mer start positions:
//   1. the chunk's bases (+k-1 lookahead) are staged in shared memory with coalesced loads;
//   2. every lane hashes CHUNK/32 k-mers with SipHash-1-3 (zero key, the CPython `hash(str)` the
//      reference orders k-mers by, main:85,94); hashes go to a shared ring that also keeps the
//      w previous hashes;
//   3. selection: every lane computes, for the 8 positions q = u*32 + lane (interleaved, so the hash
//      ring is read without bank conflicts), the window minimum, its LAST arg-min and whether the
//      minimum occurs more than once in the window.  If no window of the
//      chunk has a duplicated minimum the reference's state machine is history-free
//      ("emit when the last arg-min changes", DESIGN.md section minimizers) and the chunk is
//      emitted in parallel (one ballot per group of 32 positions orders the output); otherwise lane 0 replays the chunk with the
//      reference's sequential rule (older equal hash is kept until it leaves the window).
// A second kernel compacts the per-read results into the caller's (out_pos, out_off) layout.
#include "common.cuh"
#include "siphash.cuh"
#include <cub/device/device_scan.cuh>

namespace {

constexpr int CHUNK = 256;          // k-mer positions per warp iteration (8 per lane)
constexpr int PER_LANE = CHUNK / 32;
constexpr int KMAX = 128;           // largest k (bytes per k-mer)
constexpr int WMAX = 128;           // largest w = w_size - k
constexpr int WARPS = 4;

// CPython siphash13 with k0=k1=0 over `len` bytes starting at byte offset `a` of the 4-byte aligned
// shared buffer `cw`.  Result as Py_hash_t (-1 -> -2), hash of the empty slice is 0.
__device__ __forceinline__ int64_t py_siphash13(const uint32_t *cw, int a, int len)
{
    if (len <= 0) return 0;
    const int wi = a >> 2;
    const int sh = (a & 3) * 8;
    uint64_t v0 = 0x736f6d6570736575ULL, v1 = 0x646f72616e646f6dULL;
    uint64_t v2 = 0x6c7967656e657261ULL, v3 = 0x7465646279746573ULL;
    const int nfull = len >> 3;
    uint32_t w0 = cw[wi];
    int j = 0;
    for (; j < nfull; ++j) {
        uint32_t w1 = cw[wi + 2 * j + 1], w2 = cw[wi + 2 * j + 2];
        uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
        uint64_t m = ((uint64_t)hi << 32) | lo;
        v3 ^= m;
        SIPROUND;
        v0 ^= m;
        w0 = w2;
    }
    const int rem = len & 7;
    uint64_t b = (uint64_t)len << 56;
    if (rem) {
        uint32_t w1 = cw[wi + 2 * j + 1], w2 = cw[wi + 2 * j + 2];
        uint32_t lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
        uint64_t t = ((uint64_t)hi << 32) | lo;
        t &= (~0ULL) >> (64 - 8 * rem);
        b |= t;
    }
    v3 ^= b;
    SIPROUND;
    v0 ^= b;
    v2 ^= 0xff;
    SIPROUND;
    SIPROUND;
    SIPROUND;
    int64_t x = (int64_t)((v0 ^ v1) ^ (v2 ^ v3));
    return x == -1 ? -2 : x;
}

// Per read r the kernel writes its minimizer positions to tmp_pos[off[r] + r ...] and the count
// to cnt[r].  Capacity per read is len+1 (a read of any length emits at least the initial one).
__global__ void __launch_bounds__(WARPS * 32)
minimizer_kernel(const uint8_t *__restrict__ bases, const int64_t *__restrict__ off, int R, int k, int w,
                 int32_t *__restrict__ tmp_pos, int32_t *__restrict__ cnt)
{
    __shared__ __align__(16) uint32_t s_chars[WARPS][(CHUNK + KMAX + 16) / 4];
    __shared__ __align__(16) int64_t s_hash[WARPS][CHUNK + WMAX];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t *cw = s_chars[warp];
    uint8_t *cb = reinterpret_cast<uint8_t *>(cw);
    int64_t *hs = s_hash[warp];
    const int nbuf = CHUNK + k + 8;     // bytes staged per chunk (multiple of 4 not required)

    for (int r = blockIdx.x * WARPS + warp; r < R; r += gridDim.x * WARPS) {
        const int64_t base = off[r];
        const int n = (int)(off[r + 1] - base);
        const uint8_t *seq = bases + base;
        int32_t *out = tmp_pos + base + r;
        // positions hashed: 0..i_max ; selection runs for i in [w, i_last]
        const int i_last = max(w, n - k - 1);
        int emitted = 0;
        int pos = 0;                   // carried minimizer position (valid once i >= w)
        for (int c0 = 0; c0 <= i_last; c0 += CHUNK) {
            // ---- 1. stage bases [c0, c0+nbuf)
            for (int x = lane; x < nbuf; x += 32) {
                int g = c0 + x;
                cb[x] = (g < n) ? seq[g] : (uint8_t)0;
            }
            __syncwarp();
            // ---- 2. hashes
#pragma unroll 1
            for (int u = 0; u < PER_LANE; ++u) {
                int q = u * 32 + lane;
                int i = c0 + q;
                int64_t h = 0;
                if (i <= i_last) {
                    int len = min(k, n - i);
                    h = py_siphash13(cw, q, len);
                }
                ISOF_ASSERT(w + q < CHUNK + WMAX, CHK_MIN_SMEM);
                hs[w + q] = h;
            }
            __syncwarp();
            // ---- 3. selection: lane l takes positions q = u*32 + l, so consecutive lanes read consecutive
            // hashes (no shared-memory bank conflicts) and a ballot per u orders the emitted positions
            int la[PER_LANE];                     // last arg-min of window(i); -1 when i not selectable
            bool tie = false;
#pragma unroll
            for (int u = 0; u < PER_LANE; ++u) {
                int q = u * 32 + lane, i = c0 + q;
                la[u] = -1;
                if (i >= w && i <= i_last) {
                    int64_t mn = hs[w + q];
                    int arg = i, dup = 0;
                    for (int x = 1; x <= w; ++x) {          // newest -> oldest
                        int64_t h = hs[w + q - x];
                        if (h < mn) { mn = h; arg = i - x; dup = 0; }
                        else if (h == mn) dup = 1;
                    }
                    la[u] = arg;
                    if (dup && i > w) tie = true;           // the initial window is always canonical
                }
            }
            const bool need_seq = __any_sync(0xffffffffu, tie);
            if (!need_seq) {
#pragma unroll
                for (int u = 0; u < PER_LANE; ++u) {
                    const int q = u * 32 + lane, i = c0 + q;
                    // last arg-min of the previous position: lane-1 of this u, or lane 31 of u-1
                    int pl = __shfl_up_sync(0xffffffffu, la[u], 1);
                    if (u > 0) {
                        const int t = __shfl_sync(0xffffffffu, la[u - 1], 31);
                        if (lane == 0) pl = t;
                    }
                    bool em = false;
                    if (la[u] >= 0) {
                        if (i == w) em = true;                                  // initial minimizer
                        else if (q == 0)                                        // exact sequential rule against the carried state (main:100-108)
                            em = (pos == i - w - 1) || (hs[w] < hs[w + (pos - c0)]);
                        else em = (la[u] != pl);
                    }
                    const unsigned m = __ballot_sync(0xffffffffu, em);
                    if (em) {
                        const int slot = emitted + __popc(m & ((1u << lane) - 1u));
                        ISOF_ASSERT(slot <= n, CHK_MIN_OUT);
                        out[slot] = la[u];
                    }
                    emitted += __popc(m);
                }
                // carried state = last arg-min of the chunk's last selectable position
                const int i_end = min(c0 + CHUNK - 1, i_last);
                if (i_end >= w) {
                    const int ql = i_end - c0, ul = ql >> 5;
                    int v = la[0];
#pragma unroll
                    for (int u = 1; u < PER_LANE; ++u) if (u == ul) v = la[u];
                    pos = __shfl_sync(0xffffffffu, v, ql & 31);
                }
            } else {
                if (lane == 0) {
                    int i_beg = max(c0, w), i_end = min(c0 + CHUNK - 1, i_last);
                    for (int i = i_beg; i <= i_end; ++i) {
                        const int q = i - c0;
                        bool recompute = (i == w) || (pos == i - w - 1);
                        if (recompute) {
                            int64_t mn = hs[w + q];
                            int arg = i;
                            for (int x = 1; x <= w; ++x) {
                                int64_t h = hs[w + q - x];
                                if (h < mn) { mn = h; arg = i - x; }
                            }
                            pos = arg;
                            ISOF_ASSERT(emitted <= n, CHK_MIN_OUT);
                            out[emitted++] = pos;
                        } else if (hs[w + q] < hs[w + (pos - c0)]) {
                            pos = i;
                            out[emitted++] = pos;
                        }
                    }
                }
                emitted = __shfl_sync(0xffffffffu, emitted, 0);
                pos = __shfl_sync(0xffffffffu, pos, 0);
            }
            __syncwarp();
            // ---- 4. keep the last w hashes as history for the next chunk
            int64_t keep[WMAX / 32];
#pragma unroll
            for (int u = 0; u < WMAX / 32; ++u) {
                int x = u * 32 + lane;
                keep[u] = (x < w) ? hs[CHUNK + x] : 0;
            }
            __syncwarp();
#pragma unroll
            for (int u = 0; u < WMAX / 32; ++u) {
                int x = u * 32 + lane;
                if (x < w) hs[x] = keep[u];
            }
            __syncwarp();
        }
        if (lane == 0) cnt[r] = emitted;
    }
}

// out_off = exclusive scan of cnt is done by cub; this kernel copies each read's positions.
__global__ void minimizer_compact_kernel(const int64_t *__restrict__ off, int R, const int32_t *__restrict__ tmp_pos,
                                         const int32_t *__restrict__ cnt, const int64_t *__restrict__ out_off,
                                         int32_t *__restrict__ out_pos, int64_t out_cap)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int r = warp; r < R; r += nwarps) {
        const int32_t *src = tmp_pos + off[r] + r;
        const int64_t dst = out_off[r];
        const int c = cnt[r];
        for (int x = lane; x < c; x += 32)
            if (dst + x < out_cap) out_pos[dst + x] = src[x];
    }
}

__global__ void cnt_to_i64_kernel(const int32_t *cnt, int64_t *out, int R)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= R) out[i] = (i < R) ? (int64_t)cnt[i] : 0;
}

}  // namespace

unsigned isof_viol_minimizers() { return isof_read_violations_tu(); }

int isof_minimizers_core(isof_ctx *ctx, const uint8_t *d_bases, const int64_t *d_off, int32_t R, int64_t n_bases,
                         int32_t k, int32_t w_size, int32_t *d_out_pos, int64_t out_cap, int64_t *d_out_off,
                         int64_t *h_n_out)
{
    const int w = w_size - k;
    if (R < 0 || k <= 0 || w < 0) return isof_fail(ctx, ISOF_ERR_ARG, "bad k/w_size/n_reads (k=%d w_size=%d)", k, w_size);
    if (k > KMAX || w > WMAX - 1)
        return isof_fail(ctx, ISOF_ERR_ARG, "k=%d (max %d) or w_size-k=%d (max %d) too large", k, KMAX, w, WMAX - 1);
    if (R == 0) {
        CUDA_TRY(ctx, cudaMemsetAsync(d_out_off, 0, sizeof(int64_t), ctx->stream));
        *h_n_out = 0;
        return ISOF_OK;
    }
    TRY(dev_reserve(ctx, ctx->b_m_tmp, sizeof(int32_t) * (size_t)(n_bases + R)));
    TRY(dev_reserve(ctx, ctx->b_m_cnt, sizeof(int32_t) * (size_t)(R + 1)));
    TRY(dev_reserve(ctx, ctx->b_m_off, sizeof(int64_t) * (size_t)(R + 1)));
    int32_t *tmp = (int32_t *)ctx->b_m_tmp.p;
    int32_t *cnt = (int32_t *)ctx->b_m_cnt.p;
    int64_t *cnt64 = (int64_t *)ctx->b_m_off.p;
    const int blocks = std::min((R + WARPS - 1) / WARPS, ctx->sm_count * 16);
    {
        LaunchTimer t(ctx, SLOT_MIN);
        minimizer_kernel<<<blocks, WARPS * 32, 0, ctx->stream>>>(d_bases, d_off, R, k, w, tmp, cnt);
    }
    CUDA_TRY(ctx, cudaGetLastError());
    {
        LaunchTimer t(ctx, SLOT_MIN);
        cnt_to_i64_kernel<<<(R + 256) / 256, 256, 0, ctx->stream>>>(cnt, cnt64, R);
    }
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cnt64, d_out_off, R + 1, ctx->stream);
    TRY(dev_reserve(ctx, ctx->b_scan_tmp, tmp_bytes));
    {
        LaunchTimer t(ctx, SLOT_MIN);
        CUDA_TRY(ctx, cub::DeviceScan::ExclusiveSum(ctx->b_scan_tmp.p, tmp_bytes, cnt64, d_out_off, R + 1, ctx->stream));
    }
    {
        LaunchTimer t(ctx, SLOT_MIN);
        const int cblocks = std::min((R * 32 + 255) / 256, ctx->sm_count * 8);
        minimizer_compact_kernel<<<cblocks, 256, 0, ctx->stream>>>(d_off, R, tmp, cnt, d_out_off, d_out_pos, out_cap);
    }
    CUDA_TRY(ctx, cudaGetLastError());
    int64_t *pin = (int64_t *)ctx->pinned;
    CUDA_TRY(ctx, cudaMemcpyAsync(pin, d_out_off + R, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    *h_n_out = pin[0];
    if (pin[0] > out_cap)
        return isof_fail(ctx, ISOF_ERR_CAPACITY, "out_pos holds %lld entries, %lld needed", (long long)out_cap,
                         (long long)pin[0]);
    return ISOF_OK;
}
