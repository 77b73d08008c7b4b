// spans.cu -- kernel (c): the minimizer-pair database and the all-pairs span comparison, i.e.
//   minimizers_comb_iterator                   (reference main:215-224)
//   get_minimizer_combinations_database        (main:174-212)
//   find_most_supported_span                   (main:308-338)
// for one batch, as a sort-based group-by on the GPU.
//
// A *record* is one (anchor, partner) minimizer pair of one read: (read, p1, p2) with
// x_low < p2-p1 <= x_high.  Records are emitted in exactly the order the reference appends them to
// its dictionaries: reads in batch order, anchors left to right, partners FARTHEST first
// (main:224 reverses the list).  A *bucket* is the set of records with the same (kmer1, kmer2):
// a stable two-pass LSD radix sort on the two 64-bit k-mer keys groups them while keeping the
// reference's insertion order inside every bucket.  k-mer identity: k <= 31 and pure upper-case
// ACGT -> exact 2-bit packing; anything else (N, lower case, truncated k-mers of very short reads,
// k >= 32) -> a 63-bit SipHash brings candidates together and a resolution pass compares BYTES: the
// final key of such a k-mer is the index of the first minimizer of the batch with the same string, so
// identity is exact (span_resolve_kernel; runs only when a batch holds such k-mers).  Buckets that the reference drops (the polyA/polyA pair main:181, more than
// 3*len(reads) occurrences main:206) are flagged; singletons are kept with size 1 (the host view
// hides them as the reference deletes them, main:201-203).
// The span search then counts, for every record, the other-read records of its bucket whose span
// differs by less than delta_len: count = 1 + matches if the bucket holds >= 3 records, else 0
// (no interval, main:323).  Instance arrays are materialised on the host only for the intervals the
// weighted-interval-scheduling step picks.
#include "common.cuh"
#include "siphash.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace {

__device__ __forceinline__ int find_read(const int64_t *__restrict__ min_off, int R, int64_t x)
{
    int lo = 0, hi = R;                 // largest r with min_off[r] <= x
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (min_off[mid] <= x) lo = mid; else hi = mid;
    }
    return lo;
}

// per minimizer: read index, identity key, polyA flag, number of partners
__global__ void span_key_count_kernel(const uint8_t *__restrict__ bases, const int64_t *__restrict__ read_off, int R,
                                      const int32_t *__restrict__ min_pos, const int64_t *__restrict__ min_off, int64_t M,
                                      int k, int x_low, int x_high, int32_t *__restrict__ mread,
                                      uint64_t *__restrict__ mkey, uint8_t *__restrict__ mpolya, int64_t *__restrict__ mcnt,
                                      uint64_t hash_mask, unsigned long long *__restrict__ any_hashed)
{
    const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x > M) return;
    if (x == M) { mcnt[M] = 0; return; }
    const int r = find_read(min_off, R, x);
    mread[x] = r;
    const int64_t rb = read_off[r];
    const int n = (int)(read_off[r + 1] - rb);
    const int p = min_pos[x];
    const int len = max(0, min(k, n - p));
    const uint8_t *s = bases + rb + p;
    bool packable = (len == k) && (k <= 31);
    bool polya = (len == k);
    uint64_t key = 0;
    for (int q = 0; q < len; ++q) {
        const uint8_t c = s[q];
        int code;
        switch (c) {
            case 'A': code = 0; break;
            case 'C': code = 1; break;
            case 'G': code = 2; break;
            case 'T': code = 3; break;
            default: code = -1; break;
        }
        if (code < 0) packable = false;
        if (c != 'A') polya = false;
        key = (key << 2) | (uint64_t)(code & 3);
    }
    if (!packable) {
        // hash_mask < 2^63 - 1 is a test knob: it forces collisions so that the byte comparison below is exercised
        key = ((uint64_t)py_siphash13_bytes(s, len) & 0x7fffffffffffffffULL & hash_mask) | 0x8000000000000000ULL;
        atomicOr(any_hashed, 1ull);
    }
    mkey[x] = key;
    mpolya[x] = polya ? 1 : 0;
    // partners: later minimizers of the same read with x_low < d <= x_high (positions increase)
    const int64_t e = min_off[r + 1];
    int64_t c = 0;
    for (int64_t y = x + 1; y < e; ++y) {
        const int d = min_pos[y] - p;
        if (d > x_high) break;
        if (d > x_low) ++c;
    }
    mcnt[x] = c;
}

// ---- exact identity for k-mers that do not pack into 2 bits per base
__global__ void span_hashkey_kernel(const uint64_t *__restrict__ mkey, int64_t M, uint64_t *__restrict__ skey,
                                    uint32_t *__restrict__ sidx)
{
    const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= M) return;
    skey[x] = (mkey[x] >> 63) ? mkey[x] : 0ull;         // packed keys sort to the front and are left alone
    sidx[x] = (uint32_t)x;
}

// One thread per run of equal hash keys (stable sort: minimizer indices ascend inside a run).  Every element gets the
// key TOPBIT | (index of the first minimizer of the run with the same bytes): equal strings <=> equal keys, whatever
// the hash did.  A run of identical strings (the common case: one repeated N-containing k-mer) costs one comparison
// per element; distinct strings under one hash value cost one comparison per distinct representative.
__global__ void span_resolve_kernel(const uint8_t *__restrict__ bases, const int64_t *__restrict__ read_off,
                                    const int32_t *__restrict__ mread, const int32_t *__restrict__ min_pos, int k,
                                    const uint64_t *__restrict__ skey, const uint32_t *__restrict__ sidx, int64_t M,
                                    uint64_t *__restrict__ mkey)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= M) return;
    const uint64_t key = skey[s];
    if (!(key >> 63) || (s > 0 && skey[s - 1] == key)) return;          // not the start of a run of hashed keys
    const uint64_t TOP = 0x8000000000000000ULL;
    for (int64_t e = s; e < M && skey[e] == key; ++e) {
        const uint32_t xe = sidx[e];
        const int64_t rbe = read_off[mread[xe]];
        const int pe = min_pos[xe];
        const int le = max(0, min(k, (int)(read_off[mread[xe] + 1] - rbe) - pe));
        const uint8_t *be = bases + rbe + pe;
        uint32_t repx = xe;
        for (int64_t f = s; f < e; ++f) {
            const uint32_t xf = sidx[f];
            if (mkey[xf] != (TOP | (uint64_t)xf)) continue;              // not a representative
            const int64_t rbf = read_off[mread[xf]];
            const int pf = min_pos[xf];
            const int lf = max(0, min(k, (int)(read_off[mread[xf] + 1] - rbf) - pf));
            if (lf != le) continue;
            const uint8_t *bf = bases + rbf + pf;
            bool same = true;
            for (int q = 0; q < le; ++q) if (be[q] != bf[q]) { same = false; break; }
            if (same) { repx = xf; break; }
        }
        mkey[xe] = TOP | (uint64_t)repx;
    }
}

__global__ void span_emit_kernel(const int32_t *__restrict__ min_pos, const int64_t *__restrict__ min_off, int64_t M,
                                 int x_low, int x_high, const int32_t *__restrict__ mread, const uint64_t *__restrict__ mkey,
                                 const uint8_t *__restrict__ mpolya, const int64_t *__restrict__ moff_rec,
                                 uint64_t *__restrict__ key1, uint64_t *__restrict__ key2, int32_t *__restrict__ rread,
                                 int32_t *__restrict__ rp1, int32_t *__restrict__ rp2, uint8_t *__restrict__ rforb,
                                 uint32_t *__restrict__ iota, const int32_t *__restrict__ batch_read_off, int n_batches,
                                 uint32_t *__restrict__ rbatch)
{
    const int64_t x = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= M) return;
    const int64_t cnt = moff_rec[x + 1] - moff_rec[x];
    if (cnt == 0) return;
    const int r = mread[x];
    const int p = min_pos[x];
    const int64_t e = min_off[r + 1];
    int64_t first = -1, last = -1;
    for (int64_t y = x + 1; y < e; ++y) {
        const int d = min_pos[y] - p;
        if (d > x_high) break;
        if (d > x_low) { if (first < 0) first = y; last = y; }
    }
    int64_t g = moff_rec[x];
    const uint64_t k1 = mkey[x];
    const uint8_t pa = mpolya[x];
    uint32_t bt = 0;
    if (batch_read_off) {                       // largest b with batch_read_off[b] <= r
        int lo = 0, hi = n_batches;
        while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (batch_read_off[mid] <= r) lo = mid; else hi = mid; }
        bt = (uint32_t)lo;
    }
    for (int64_t y = last; y >= first; --y, ++g) {          // farthest partner first
        ISOF_ASSERT(g < moff_rec[x + 1], CHK_SPAN_REC);
        key1[g] = k1;
        key2[g] = mkey[y];
        rread[g] = r;
        rp1[g] = p;
        rp2[g] = min_pos[y];
        rforb[g] = pa & mpolya[y];
        iota[g] = (uint32_t)g;
        if (rbatch) rbatch[g] = bt;
    }
}

__global__ void gather_u64_kernel(const uint64_t *__restrict__ src, const uint32_t *__restrict__ idx, int64_t G,
                                  uint64_t *__restrict__ dst)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < G) dst[s] = src[idx[s]];
}

// head[s] = 1 where a new bucket starts in sorted order
__global__ void span_head_kernel(const uint64_t *__restrict__ key1, const uint64_t *__restrict__ key2,
                                 const uint32_t *__restrict__ rbatch, const uint32_t *__restrict__ perm, int64_t G,
                                 int32_t *__restrict__ head)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= G) return;
    int h = 1;
    if (s > 0) {
        const uint32_t a = perm[s], b = perm[s - 1];
        h = (key1[a] != key1[b]) || (key2[a] != key2[b]) || (rbatch && rbatch[a] != rbatch[b]);
    }
    head[s] = h;
}

__global__ void gather_u32_kernel(const uint32_t *__restrict__ src, const uint32_t *__restrict__ idx, int64_t G,
                                  uint32_t *__restrict__ dst)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < G) dst[s] = src[idx[s]];
}

// bid_incl = inclusive sum of head; bucket b = bid_incl-1
__global__ void span_bucket_kernel(const int32_t *__restrict__ head, const int32_t *__restrict__ bid_incl,
                                   const uint32_t *__restrict__ perm, int64_t G, int32_t *__restrict__ rec_bucket,
                                   int32_t *__restrict__ bucket_off, int32_t *__restrict__ sorted_rec)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= G) return;
    const int b = bid_incl[s] - 1;
    const uint32_t g = perm[s];
    rec_bucket[g] = b;
    sorted_rec[s] = (int32_t)g;
    if (head[s]) bucket_off[b] = (int32_t)s;
    if (s == G - 1) bucket_off[b + 1] = (int32_t)G;
}

__global__ void span_drop_kernel(const int32_t *__restrict__ bucket_off, int32_t B, const int32_t *__restrict__ sorted_rec,
                                 const uint8_t *__restrict__ rforb, int R, const int32_t *__restrict__ batch_read_off,
                                 const uint32_t *__restrict__ rbatch, uint8_t *__restrict__ dropped)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int st = bucket_off[b], sz = bucket_off[b + 1] - st;
    int reads_in_batch = R;                      // main:206 uses len(reads) of the batch the bucket belongs to
    if (batch_read_off) { const uint32_t bt = rbatch[sorted_rec[st]]; reads_in_batch = batch_read_off[bt + 1] - batch_read_off[bt]; }
    dropped[b] = (rforb[sorted_rec[st]] || sz > 3 * reads_in_batch) ? 1 : 0;
}

// find_most_supported_span: per record, 1 + number of other-read records of the bucket within delta_len
__global__ void span_count_kernel(const int32_t *__restrict__ rread, const int32_t *__restrict__ rp1,
                                  const int32_t *__restrict__ rp2, const int32_t *__restrict__ rec_bucket,
                                  const int32_t *__restrict__ bucket_off, const uint8_t *__restrict__ dropped,
                                  const int32_t *__restrict__ sorted_rec, int64_t G, int delta_len,
                                  int32_t *__restrict__ rec_count)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    const int b = rec_bucket[g];
    const int st = bucket_off[b], sz = bucket_off[b + 1] - st;
    int cnt = 0;
    if (!dropped[b] && sz >= 3) {
        const int r = rread[g], span = rp2[g] - rp1[g];
        cnt = 1;
        for (int s = st; s < st + sz; ++s) {
            const int h = sorted_rec[s];
            if (rread[h] != r && abs(span - (rp2[h] - rp1[h])) < delta_len) ++cnt;
        }
    }
    rec_count[g] = cnt;
}

}  // namespace

unsigned isof_viol_spans() { return isof_read_violations_tu(); }

// All pointers device pointers.  h_n_rec / h_n_buckets are host pointers.
int isof_spans_core(isof_ctx *ctx, const uint8_t *d_bases, const int64_t *d_read_off, int32_t R, const int32_t *d_min_pos,
                    const int64_t *d_min_off, int64_t M, int32_t k, int32_t x_low, int32_t x_high, int32_t delta_len,
                    int64_t rec_cap, int32_t *d_rec_read, int32_t *d_rec_p1, int32_t *d_rec_p2, int32_t *d_rec_count,
                    int32_t *d_rec_bucket, int32_t *d_sorted_rec, int32_t *d_bucket_off, uint8_t *d_bucket_dropped,
                    int64_t *h_n_rec, int64_t *h_n_buckets, const int32_t *d_batch_read_off, int32_t n_batches)
{
    cudaStream_t st = ctx->stream;
    if (n_batches <= 1) d_batch_read_off = nullptr;
    *h_n_rec = 0;
    *h_n_buckets = 0;
    if (R < 0 || M < 0 || k <= 0) return isof_fail(ctx, ISOF_ERR_ARG, "bad span arguments");
    if (R == 0 || M == 0) return ISOF_OK;
    // ---- per-minimizer scratch
    TRY(dev_reserve(ctx, ctx->b_s_mread, sizeof(int32_t) * (size_t)M));
    TRY(dev_reserve(ctx, ctx->b_s_mkey, sizeof(uint64_t) * (size_t)M));
    TRY(dev_reserve(ctx, ctx->b_s_mpolya, (size_t)M));
    TRY(dev_reserve(ctx, ctx->b_s_mcnt, sizeof(int64_t) * (size_t)(M + 1)));
    TRY(dev_reserve(ctx, ctx->b_s_moff, sizeof(int64_t) * (size_t)(M + 1)));
    TRY(dev_reserve(ctx, ctx->b_misc, 128));
    unsigned long long *any_hashed = (unsigned long long *)ctx->b_misc.p + 12;
    CUDA_TRY(ctx, cudaMemsetAsync(any_hashed, 0, 8, st));
    int32_t *mread = (int32_t *)ctx->b_s_mread.p;
    uint64_t *mkey = (uint64_t *)ctx->b_s_mkey.p;
    uint8_t *mpolya = (uint8_t *)ctx->b_s_mpolya.p;
    int64_t *mcnt = (int64_t *)ctx->b_s_mcnt.p, *moff = (int64_t *)ctx->b_s_moff.p;
    {
        LaunchTimer t(ctx, SLOT_OTHER);
        span_key_count_kernel<<<(unsigned)((M + 1 + 255) / 256), 256, 0, st>>>(d_bases, d_read_off, R, d_min_pos, d_min_off, M,
                                                                               k, x_low, x_high, mread, mkey, mpolya, mcnt,
                                                                               ctx->span_hash_mask, any_hashed);
    }
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, mcnt, moff, M + 1, st);
    TRY(dev_reserve(ctx, ctx->b_scan_tmp, tb));
    {
        LaunchTimer t(ctx, SLOT_OTHER);
        CUDA_TRY(ctx, cub::DeviceScan::ExclusiveSum(ctx->b_scan_tmp.p, tb, mcnt, moff, M + 1, st));
    }
    int64_t *pin = (int64_t *)ctx->pinned;
    CUDA_TRY(ctx, cudaMemcpyAsync(pin, moff + M, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(pin + 1, any_hashed, 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    const int64_t G = pin[0];
    if (pin[1] && G > 0) {
        // some k-mers of this batch do not pack into 2 bits per base: replace their hash keys by exact identities
        TRY(dev_reserve(ctx, ctx->b_s_keyt, sizeof(uint64_t) * (size_t)M));
        TRY(dev_reserve(ctx, ctx->b_s_keyo, sizeof(uint64_t) * (size_t)M));
        TRY(dev_reserve(ctx, ctx->b_s_idx0, sizeof(uint32_t) * (size_t)M));
        TRY(dev_reserve(ctx, ctx->b_s_idx1, sizeof(uint32_t) * (size_t)M));
        uint64_t *skey = (uint64_t *)ctx->b_s_keyt.p, *skey_o = (uint64_t *)ctx->b_s_keyo.p;
        uint32_t *sidx = (uint32_t *)ctx->b_s_idx0.p, *sidx_o = (uint32_t *)ctx->b_s_idx1.p;
        {
            LaunchTimer t(ctx, SLOT_OTHER);
            span_hashkey_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(mkey, M, skey, sidx);
        }
        size_t th = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, th, skey, skey_o, sidx, sidx_o, (int)M, 0, 64, st);
        TRY(dev_reserve(ctx, ctx->b_s_sorttmp, th));
        {
            LaunchTimer t(ctx, SLOT_OTHER);
            CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(ctx->b_s_sorttmp.p, th, skey, skey_o, sidx, sidx_o, (int)M, 0, 64, st));
        }
        {
            LaunchTimer t(ctx, SLOT_OTHER);
            span_resolve_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(d_bases, d_read_off, mread, d_min_pos, k, skey_o, sidx_o,
                                                                             M, mkey);
        }
        CUDA_TRY(ctx, cudaGetLastError());
    }
    *h_n_rec = G;
    if (G > rec_cap) return isof_fail(ctx, ISOF_ERR_CAPACITY, "record arrays hold %lld, %lld needed", (long long)rec_cap, (long long)G);
    if (G >= (1LL << 31)) return isof_fail(ctx, ISOF_ERR_RANGE, "more than 2^31 minimizer pairs in one batch");
    if (G == 0) return ISOF_OK;
    // ---- records
    TRY(dev_reserve(ctx, ctx->b_s_key1, sizeof(uint64_t) * (size_t)G));
    TRY(dev_reserve(ctx, ctx->b_s_key2, sizeof(uint64_t) * (size_t)G));
    TRY(dev_reserve(ctx, ctx->b_s_keyt, sizeof(uint64_t) * (size_t)G));
    TRY(dev_reserve(ctx, ctx->b_s_keyo, sizeof(uint64_t) * (size_t)G));
    TRY(dev_reserve(ctx, ctx->b_s_idx0, sizeof(uint32_t) * (size_t)G));
    TRY(dev_reserve(ctx, ctx->b_s_idx1, sizeof(uint32_t) * (size_t)G));
    TRY(dev_reserve(ctx, ctx->b_s_idx2, sizeof(uint32_t) * (size_t)G));
    TRY(dev_reserve(ctx, ctx->b_s_forb, (size_t)G));
    TRY(dev_reserve(ctx, ctx->b_s_head, sizeof(int32_t) * (size_t)G));
    uint32_t *rbatch = nullptr, *rbatch_t = nullptr;
    if (d_batch_read_off) {
        TRY(dev_reserve(ctx, ctx->b_s_rbatch, sizeof(uint32_t) * (size_t)G * 2));
        rbatch = (uint32_t *)ctx->b_s_rbatch.p;
        rbatch_t = rbatch + G;
    }
    TRY(dev_reserve(ctx, ctx->b_s_bid, sizeof(int32_t) * (size_t)G));
    uint64_t *key1 = (uint64_t *)ctx->b_s_key1.p, *key2 = (uint64_t *)ctx->b_s_key2.p;
    uint64_t *keyt = (uint64_t *)ctx->b_s_keyt.p, *keyo = (uint64_t *)ctx->b_s_keyo.p;
    uint32_t *idx0 = (uint32_t *)ctx->b_s_idx0.p, *idx1 = (uint32_t *)ctx->b_s_idx1.p, *idx2 = (uint32_t *)ctx->b_s_idx2.p;
    uint8_t *rforb = (uint8_t *)ctx->b_s_forb.p;
    int32_t *head = (int32_t *)ctx->b_s_head.p, *bid = (int32_t *)ctx->b_s_bid.p;
    const unsigned gb = (unsigned)((G + 255) / 256);
    {
        LaunchTimer t(ctx, SLOT_OTHER);
        span_emit_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(d_min_pos, d_min_off, M, x_low, x_high, mread, mkey, mpolya,
                                                                      moff, key1, key2, d_rec_read, d_rec_p1, d_rec_p2, rforb, idx0,
                                                                      d_batch_read_off, n_batches, rbatch);
    }
    // ---- stable LSD sort: by key2, then by key1
    const int end_bit = 64;          // packed keys use 2k bits, but a hashed (fallback) key sets bit 63
    size_t ts = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, ts, key2, keyo, idx0, idx1, (int)G, 0, end_bit, st);
    TRY(dev_reserve(ctx, ctx->b_s_sorttmp, ts));
    {
        LaunchTimer t(ctx, SLOT_OTHER);
        CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(ctx->b_s_sorttmp.p, ts, key2, keyo, idx0, idx1, (int)G, 0, end_bit, st));
    }
    {
        LaunchTimer t(ctx, SLOT_OTHER);
        gather_u64_kernel<<<gb, 256, 0, st>>>(key1, idx1, G, keyt);
    }
    {
        LaunchTimer t(ctx, SLOT_OTHER);
        CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(ctx->b_s_sorttmp.p, ts, keyt, keyo, idx1, idx2, (int)G, 0, end_bit, st));
    }
    if (rbatch) {
        // third (most significant) pass: the batch a record belongs to, so that buckets never mix batches
        {
            LaunchTimer t(ctx, SLOT_OTHER);
            gather_u32_kernel<<<gb, 256, 0, st>>>(rbatch, idx2, G, rbatch_t);
        }
        int nbits = 1;
        while ((1 << nbits) < n_batches) ++nbits;
        size_t t4 = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, t4, rbatch_t, (uint32_t *)keyo, idx2, idx1, (int)G, 0, nbits, st);
        TRY(dev_reserve(ctx, ctx->b_s_sorttmp, std::max(ts, t4)));
        {
            LaunchTimer t(ctx, SLOT_OTHER);
            CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(ctx->b_s_sorttmp.p, t4, rbatch_t, (uint32_t *)keyo, idx2, idx1, (int)G, 0,
                                                          nbits, st));
        }
        std::swap(idx1, idx2);                  // idx2 again holds the final permutation
    }
    // ---- buckets
    {
        LaunchTimer t(ctx, SLOT_OTHER);
        span_head_kernel<<<gb, 256, 0, st>>>(key1, key2, rbatch, idx2, G, head);
    }
    size_t t3 = 0;
    cub::DeviceScan::InclusiveSum(nullptr, t3, head, bid, (int)G, st);
    TRY(dev_reserve(ctx, ctx->b_scan_tmp, t3));
    {
        LaunchTimer t(ctx, SLOT_OTHER);
        CUDA_TRY(ctx, cub::DeviceScan::InclusiveSum(ctx->b_scan_tmp.p, t3, head, bid, (int)G, st));
    }
    {
        LaunchTimer t(ctx, SLOT_OTHER);
        span_bucket_kernel<<<gb, 256, 0, st>>>(head, bid, idx2, G, d_rec_bucket, d_bucket_off, d_sorted_rec);
    }
    int32_t *pin32 = (int32_t *)ctx->pinned;
    CUDA_TRY(ctx, cudaMemcpyAsync(pin32, bid + (G - 1), 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    const int32_t B = pin32[0];
    *h_n_buckets = B;
    {
        LaunchTimer t(ctx, SLOT_OTHER);
        span_drop_kernel<<<(B + 255) / 256, 256, 0, st>>>(d_bucket_off, B, d_sorted_rec, rforb, R, d_batch_read_off, rbatch,
                                                          d_bucket_dropped);
    }
    {
        LaunchTimer t(ctx, SLOT_OTHER);
        span_count_kernel<<<gb, 256, 0, st>>>(d_rec_read, d_rec_p1, d_rec_p2, d_rec_bucket, d_bucket_off, d_bucket_dropped,
                                              d_sorted_rec, G, delta_len, d_rec_count);
    }
    CUDA_TRY(ctx, cudaGetLastError());
    return ISOF_OK;
}
