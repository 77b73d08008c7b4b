/*
 * isof_oracle.c -- CPU restatement of the isONform hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (isonform_b200/) never links, imports or calls it.
 *
 * What is restated, and where it comes from (file:line relative to the reference tree):
 *   - isof_oracle_pyhash          CPython 3.12 hash(str) under PYTHONHASHSEED=0 == SipHash-1-3 with a
 *                                 zero key over the ASCII bytes (third-party: CPython Python/pyhash.c,
 *                                 3.11+ `siphash13`).  Used by main:85,94 as the minimizer order.
 *                                 PINNED: tests compare with the running interpreter's hash().
 *   - isof_oracle_minimizers      get_kmer_minimizers, main:81-110 (+ rindex main:72-79), restated
 *                                 literally (deque window, stale minimizer_pos).  PINNED by executing
 *                                 the reference (tests/golden/minimizers_*.json).
 *   - isof_oracle_pairs / spans   minimizers_comb_iterator main:215-224,
 *                                 get_minimizer_combinations_database main:174-212 and
 *                                 find_most_supported_span main:308-338 restated over integer k-mer
 *                                 ids.  PINNED by executing the reference.
 *   - isof_oracle_sg_align        consensus.parasail_alignment, modules/consensus.py:55-67, i.e.
 *                                 parasail.sg_trace_scan_16/_32 + result.cigar.  Third-party
 *                                 (PyPI parasail >=1.3.3 / jeffdaily/parasail 2.x), source NOT under
 *                                 the reference tree and not installable here.
 *                                 *** PARITY UNPINNED *** : restated from the library's published
 *                                 algorithm (Gotoh affine semi-global, all four end gaps free,
 *                                 trace table + cigar walk); every tie-break lives in
 *                                 isof_tiebreak_t below so it can be re-validated on a host that
 *                                 has parasail.
 *
 * Build: see oracle/Makefile (gcc -O3 -march=native -fopenmp -shared -fPIC).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#if defined(__GNUC__)
#define ISOF_API __attribute__((visibility("default")))
#else
#define ISOF_API
#endif

/* ------------------------------------------------------------------------------------------ */
/* SipHash-1-3, zero key  (CPython Python/pyhash.c `siphash13`)                                 */
/* ------------------------------------------------------------------------------------------ */

#define ROTL64(x, b) (uint64_t)(((x) << (b)) | ((x) >> (64 - (b))))
#define SIPROUND(v0, v1, v2, v3)                                                                  \
    do {                                                                                          \
        v0 += v1; v1 = ROTL64(v1, 13); v1 ^= v0; v0 = ROTL64(v0, 32);                             \
        v2 += v3; v3 = ROTL64(v3, 16); v3 ^= v2;                                                  \
        v0 += v3; v3 = ROTL64(v3, 21); v3 ^= v0;                                                  \
        v2 += v1; v1 = ROTL64(v1, 17); v1 ^= v2; v2 = ROTL64(v2, 32);                             \
    } while (0)

ISOF_API int64_t isof_oracle_pyhash(const uint8_t *s, int64_t len)
{
    if (len <= 0) return 0;                       /* hash("") == 0 */
    uint64_t k0 = 0, k1 = 0;                      /* PYTHONHASHSEED=0 */
    uint64_t b = (uint64_t)len << 56;
    uint64_t v0 = k0 ^ 0x736f6d6570736575ULL;
    uint64_t v1 = k1 ^ 0x646f72616e646f6dULL;
    uint64_t v2 = k0 ^ 0x6c7967656e657261ULL;
    uint64_t v3 = k1 ^ 0x7465646279746573ULL;
    int64_t n = len;
    const uint8_t *p = s;
    while (n >= 8) {
        uint64_t mi = 0;
        for (int q = 0; q < 8; ++q) mi |= (uint64_t)p[q] << (8 * q);   /* little endian */
        v3 ^= mi;
        SIPROUND(v0, v1, v2, v3);
        v0 ^= mi;
        p += 8; n -= 8;
    }
    uint64_t t = 0;
    for (int q = 0; q < (int)n; ++q) t |= (uint64_t)p[q] << (8 * q);
    b |= t;
    v3 ^= b;
    SIPROUND(v0, v1, v2, v3);
    v0 ^= b;
    v2 ^= 0xff;
    SIPROUND(v0, v1, v2, v3);
    SIPROUND(v0, v1, v2, v3);
    SIPROUND(v0, v1, v2, v3);
    uint64_t h = (v0 ^ v1) ^ (v2 ^ v3);
    int64_t x = (int64_t)h;
    if (x == -1) x = -2;                          /* CPython reserves -1 for errors */
    return x;
}

/* hash(seq[i:i+k]) with Python slice truncation */
static int64_t slice_hash(const uint8_t *seq, int64_t n, int64_t i, int k)
{
    if (i >= n) return 0;
    int64_t len = (i + k <= n) ? k : (n - i);
    return isof_oracle_pyhash(seq + i, len);
}

/* ------------------------------------------------------------------------------------------ */
/* get_kmer_minimizers  (main:81-110)                                                          */
/* ------------------------------------------------------------------------------------------ */

/* Returns the number of minimizers; writes at most `cap` positions (the k-mer string is
 * seq[pos:pos+k], rebuilt by the caller).  Literal restatement: the window is a ring buffer of
 * the last w+1 hashes, minimizer_pos is only refreshed in the recompute branch. */
ISOF_API int64_t isof_oracle_minimizers(const uint8_t *seq, int64_t n, int k, int w_size,
                                        int32_t *out_pos, int64_t cap)
{
    int w = w_size - k;
    if (w < 0) return -1;
    int W = w + 1;
    int64_t *win = (int64_t *)malloc(sizeof(int64_t) * (size_t)W);
    int64_t cnt = 0;
    /* window_kmers = deque(hash(seq[i:i+k]) for i in range(w+1)) */
    for (int i = 0; i < W; ++i) win[i] = slice_hash(seq, n, i, k);
    int64_t curr_min = win[0];
    for (int i = 1; i < W; ++i) if (win[i] < curr_min) curr_min = win[i];
    int64_t minimizer_pos = 0;
    for (int i = 0; i < W; ++i) if (win[i] == curr_min) minimizer_pos = i;   /* rindex: last */
    if (cnt < cap) out_pos[cnt] = (int32_t)minimizer_pos;
    ++cnt;
    int head = 0;                                   /* index of the oldest element */
    for (int64_t i = w + 1; i < n - k; ++i) {
        int64_t new_kmer = slice_hash(seq, n, i, k);
        int64_t discarded = win[head];
        win[head] = new_kmer;
        head = (head + 1) % W;                      /* oldest is now win[head] == position i-w */
        if (curr_min == discarded && minimizer_pos < i - w) {
            curr_min = win[head];
            for (int q = 1; q < W; ++q) {
                int64_t v = win[(head + q) % W];
                if (v < curr_min) curr_min = v;
            }
            int last = 0;
            for (int q = 0; q < W; ++q) if (win[(head + q) % W] == curr_min) last = q;
            minimizer_pos = last + i - w;
            if (cnt < cap) out_pos[cnt] = (int32_t)minimizer_pos;
            ++cnt;
        } else if (new_kmer < curr_min) {
            curr_min = new_kmer;
            if (cnt < cap) out_pos[cnt] = (int32_t)i;   /* minimizer_pos deliberately not updated */
            ++cnt;
        }
    }
    free(win);
    return cnt;
}

/* Batch form used by the CPU baseline: reads concatenated, off[R+1]; out_off[R+1] filled. */
ISOF_API int64_t isof_oracle_minimizers_batch(const uint8_t *cat, const int64_t *off, int R, int k,
                                              int w_size, int32_t *out_pos, int64_t *out_off,
                                              int64_t cap, int threads)
{
    /* two passes so that the output is packed in read order regardless of thread schedule */
    int64_t *cnt = (int64_t *)calloc((size_t)R + 1, sizeof(int64_t));
    int32_t **tmp = (int32_t **)calloc((size_t)R, sizeof(int32_t *));
    (void)threads;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel for schedule(dynamic, 16)
#endif
    for (int r = 0; r < R; ++r) {
        int64_t n = off[r + 1] - off[r];
        int64_t c = n > 0 ? n : 1;
        tmp[r] = (int32_t *)malloc(sizeof(int32_t) * (size_t)c);
        cnt[r] = isof_oracle_minimizers(cat + off[r], n, k, w_size, tmp[r], c);
    }
    int64_t tot = 0;
    for (int r = 0; r < R; ++r) {
        out_off[r] = tot;
        int64_t c = cnt[r] < 0 ? 0 : cnt[r];
        if (tot + c <= cap) memcpy(out_pos + tot, tmp[r], sizeof(int32_t) * (size_t)c);
        tot += c;
        free(tmp[r]);
    }
    out_off[R] = tot;
    free(tmp); free(cnt);
    return tot;
}

/* ------------------------------------------------------------------------------------------ */
/* parasail sg_trace + cigar  (modules/consensus.py:55-67)  *** PARITY UNPINNED ***             */
/* ------------------------------------------------------------------------------------------ */

/* Every tie-break of the restatement, in one place.  Defaults = the library as restated
 * (see DESIGN.md "parasail tie-break table"). */
typedef struct {
    /* H source on ties.  0: DIAG > F(walks query i) > E(walks ref j).  1: DIAG > E > F. */
    int32_t h_e_before_f;
    /* gap state on ties open==extend.  0: extend (open only if strictly greater).  1: open. */
    int32_t gap_open_on_tie;
    /* end cell.  0: last ROW first (smallest j, strict >), then last COLUMN only if strictly
     * greater (smallest i); equal score and end_ref==m-1 takes the smaller i.
     * 1: last COLUMN first (smallest i), then last row only if strictly greater (smallest j). */
    int32_t end_col_first;
    /* CIGAR letters of the two gap states.  0: SAM convention with s1 = query -- the gap that walks the
     * query axis i (state F) is 'I', the one that walks the ref axis j (state E) is 'D'; this is what
     * the reference's cigar_to_seq (consensus.py:35-47) assumes.  1: letters swapped (the library's own
     * state names: E = "INS" -> 'I', F = "DEL" -> 'D'). */
    int32_t swap_id;
} isof_tiebreak_t;

/* BAM op codes used in the run encoding (len << 4 | op), as parasail's cigar->seq does. */
enum { OP_M = 0, OP_I = 1, OP_D = 2, OP_EQ = 7, OP_X = 8 };

/* parasail_matrix_create("ACGT", match, mismatch): case-insensitive, everything outside the
 * alphabet maps to the wildcard row/column which scores 0. */
static inline int base_code(uint8_t c)
{
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return 4;
    }
}

static inline int upc(uint8_t c) { return (c >= 'a' && c <= 'z') ? c - 32 : c; }

#define T_SRC_DIAG 3
#define T_SRC_F 2       /* parasail "DEL": walks the query axis i -> SAM 'I' */
#define T_SRC_E 1       /* parasail "INS": walks the ref axis j   -> SAM 'D' */
#define T_E_EXT 1       /* bit0: E(i,j) came from E(i,j-1)-ext  */
#define T_F_EXT 2       /* bit1: F(i,j) came from F(i-1,j)-ext  */

/*
 * s1 = query (rows i), s2 = ref (columns j).  Outputs: score, end_query, end_ref (0-based cell),
 * cigar runs in alignment order.  Returns the number of runs (>=0), -1 on bad input, -2 if
 * cigar_cap is too small (n_runs still the true count via *n_runs_out).
 */
ISOF_API int32_t isof_oracle_sg_align_tb(const uint8_t *s1, int32_t n, const uint8_t *s2, int32_t m,
                                         int32_t match, int32_t mismatch, int32_t open, int32_t ext,
                                         const isof_tiebreak_t *tbk, int32_t *score_out,
                                         int32_t *end_q_out, int32_t *end_r_out, uint32_t *cigar,
                                         int32_t cigar_cap, int32_t *n_runs_out)
{
    if (n <= 0 || m <= 0) return -1;               /* parasail rejects empty sequences */
    isof_tiebreak_t tb = {0, 0, 0, 0};
    if (tbk) tb = *tbk;
    const int32_t NEG = -(1 << 29);
    int32_t S[5][5];
    for (int a = 0; a < 5; ++a)
        for (int b = 0; b < 5; ++b)
            S[a][b] = (a == 4 || b == 4) ? 0 : (a == b ? match : mismatch);

    uint8_t *trace = (uint8_t *)malloc((size_t)n * (size_t)m);
    int32_t *H = (int32_t *)malloc(sizeof(int32_t) * (size_t)(m + 1));   /* previous row, H[j+1]=H(i-1,j) */
    int32_t *F = (int32_t *)malloc(sizeof(int32_t) * (size_t)m);         /* F(i-1,j) */
    uint8_t *c2 = (uint8_t *)malloc((size_t)m);
    if (!trace || !H || !F || !c2) { free(trace); free(H); free(F); free(c2); return -3; }
    for (int j = 0; j < m; ++j) { c2[j] = (uint8_t)base_code(s2[j]); F[j] = NEG; }
    for (int j = 0; j <= m; ++j) H[j] = 0;                                /* free begin gaps */

    int32_t best_row = NEG, best_row_j = 0;     /* over the last row    */
    int32_t best_col = NEG, best_col_i = 0;     /* over the last column */

    for (int i = 0; i < n; ++i) {
        const int32_t *Srow = S[base_code(s1[i])];
        int32_t Hdiag = 0;            /* H(i-1,-1) = 0 */
        int32_t Hleft = 0;            /* H(i,-1)   = 0 */
        int32_t Eleft = NEG;          /* E(i,-1)        */
        uint8_t *trow = trace + (size_t)i * (size_t)m;
        for (int j = 0; j < m; ++j) {
            int32_t Hup = H[j + 1];
            int32_t e_opn = Hleft - open, e_ext = Eleft - ext;
            int32_t f_opn = Hup - open, f_ext = F[j] - ext;
            uint8_t t = 0;
            int32_t E, Fv;
            if (tb.gap_open_on_tie ? (e_opn >= e_ext) : (e_opn > e_ext)) E = e_opn; else { E = e_ext; t |= T_E_EXT; }
            if (tb.gap_open_on_tie ? (f_opn >= f_ext) : (f_opn > f_ext)) Fv = f_opn; else { Fv = f_ext; t |= T_F_EXT; }
            int32_t Hd = Hdiag + Srow[c2[j]];
            int32_t Hv = Hd;
            if (E > Hv) Hv = E;
            if (Fv > Hv) Hv = Fv;
            int src;
            if (Hv == Hd) src = T_SRC_DIAG;
            else if (tb.h_e_before_f) src = (Hv == E) ? T_SRC_E : T_SRC_F;
            else src = (Hv == Fv) ? T_SRC_F : T_SRC_E;
            trow[j] = (uint8_t)(t | (src << 2));
            Hdiag = Hup;
            H[j + 1] = Hv;          /* row i overwrites row i-1 */
            F[j] = Fv;
            Hleft = Hv;
            Eleft = E;
        }
        H[0] = 0;
        /* last column */
        if (H[m] > best_col) { best_col = H[m]; best_col_i = i; }
    }
    for (int j = 0; j < m; ++j)
        if (H[j + 1] > best_row) { best_row = H[j + 1]; best_row_j = j; }

    int32_t score, eq, er;
    if (!tb.end_col_first) {
        score = best_row; eq = n - 1; er = best_row_j;
        if (best_col > score) { score = best_col; eq = best_col_i; er = m - 1; }
        else if (best_col == score && er == m - 1 && best_col_i < eq) eq = best_col_i;
    } else {
        score = best_col; eq = best_col_i; er = m - 1;
        if (best_row > score) { score = best_row; eq = n - 1; er = best_row_j; }
    }
    *score_out = score; *end_q_out = eq; *end_r_out = er;

    /* ---- cigar walk (reverse), run-length encoded on the fly ---- */
    int32_t nr = 0;                 /* runs emitted (reverse order) */
    uint32_t cur_op = 0xffffffffu, cur_len = 0;
    uint32_t *rev = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n + m + 2));
#define PUSH(op, len)                                                                              \
    do {                                                                                           \
        if ((len) > 0) {                                                                           \
            if (cur_op == (uint32_t)(op)) cur_len += (uint32_t)(len);                              \
            else { if (cur_len) rev[nr++] = (cur_len << 4) | cur_op; cur_op = (op); cur_len = (len); } \
        }                                                                                          \
    } while (0)
    int32_t i = eq, j = er;
    const uint32_t op_e = tb.swap_id ? OP_I : OP_D;   /* E walks the ref axis j   */
    const uint32_t op_f = tb.swap_id ? OP_D : OP_I;   /* F walks the query axis i */
    if (eq + 1 == n) PUSH(op_e, m - 1 - j);        /* unaligned ref tail   */
    else PUSH(op_f, n - 1 - i);                    /* unaligned query tail */
    int where = T_SRC_DIAG;
    while (i >= 0 || j >= 0) {
        if (i < 0) { PUSH(op_e, j + 1); break; }
        if (j < 0) { PUSH(op_f, i + 1); break; }
        uint8_t t = trace[(size_t)i * (size_t)m + (size_t)j];
        if (where == T_SRC_DIAG) {
            int src = t >> 2;
            if (src == T_SRC_DIAG) {
                PUSH(upc(s1[i]) == upc(s2[j]) ? OP_EQ : OP_X, 1);
                --i; --j;
            } else where = src;
        } else if (where == T_SRC_E) {
            PUSH(op_e, 1); --j;
            where = (t & T_E_EXT) ? T_SRC_E : T_SRC_DIAG;
        } else {
            PUSH(op_f, 1); --i;
            where = (t & T_F_EXT) ? T_SRC_F : T_SRC_DIAG;
        }
    }
    if (cur_len) rev[nr++] = (cur_len << 4) | cur_op;
#undef PUSH
    *n_runs_out = nr;
    int32_t rc = nr;
    if (nr > cigar_cap) rc = -2;
    else for (int q = 0; q < nr; ++q) cigar[q] = rev[nr - 1 - q];
    free(rev); free(trace); free(H); free(F); free(c2);
    return rc;
}

ISOF_API int32_t isof_oracle_sg_align(const uint8_t *s1, int32_t n, const uint8_t *s2, int32_t m,
                                      int32_t match, int32_t mismatch, int32_t open, int32_t ext,
                                      int32_t *score_out, int32_t *end_q_out, int32_t *end_r_out,
                                      uint32_t *cigar, int32_t cigar_cap, int32_t *n_runs_out)
{
    return isof_oracle_sg_align_tb(s1, n, s2, m, match, mismatch, open, ext, NULL, score_out,
                                   end_q_out, end_r_out, cigar, cigar_cap, n_runs_out);
}

/*
 * Batch over a sequence pool (cat/off) and index pairs; the CPU baseline of bench.py.
 * cigar_off[P+1] is filled with exclusive prefix sums; runs that do not fit `cigar_cap` are
 * counted but not written.  Returns total runs, or <0 on error.
 */
ISOF_API int64_t isof_oracle_sg_align_pairs(const uint8_t *cat, const int64_t *off, const int32_t *ia,
                                            const int32_t *ib, int64_t P, int32_t match,
                                            int32_t mismatch, int32_t open, int32_t ext,
                                            int32_t *score, int32_t *end_q, int32_t *end_r,
                                            uint32_t *cigar, int64_t *cigar_off, int64_t cigar_cap,
                                            int threads)
{
    uint32_t **runs = (uint32_t **)calloc((size_t)P, sizeof(uint32_t *));
    int32_t *nr = (int32_t *)calloc((size_t)P, sizeof(int32_t));
    int err = 0;
    (void)threads;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int64_t p = 0; p < P; ++p) {
        int32_t n = (int32_t)(off[ia[p] + 1] - off[ia[p]]);
        int32_t m = (int32_t)(off[ib[p] + 1] - off[ib[p]]);
        int32_t cap = n + m + 2;
        runs[p] = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)cap);
        int32_t rc = isof_oracle_sg_align(cat + off[ia[p]], n, cat + off[ib[p]], m, match, mismatch,
                                          open, ext, &score[p], &end_q[p], &end_r[p], runs[p], cap,
                                          &nr[p]);
        if (rc < 0) err = 1;
    }
    int64_t tot = 0;
    for (int64_t p = 0; p < P; ++p) {
        cigar_off[p] = tot;
        if (!err && cigar && tot + nr[p] <= cigar_cap)
            memcpy(cigar + tot, runs[p], sizeof(uint32_t) * (size_t)nr[p]);
        tot += nr[p];
        free(runs[p]);
    }
    cigar_off[P] = tot;
    free(runs); free(nr);
    return err ? -1 : tot;
}

/* lanes of the inter-pair SIMD aligner and the widest instruction set its dispatcher will pick on this CPU */
ISOF_API int32_t isof_oracle_simd_lanes(void) { return 32; }
ISOF_API const char *isof_oracle_simd_isa(void)
{
#if defined(__x86_64__) && defined(__GNUC__)
    __builtin_cpu_init();
    if (__builtin_cpu_supports("avx512bw")) return "AVX-512BW";
    if (__builtin_cpu_supports("avx2")) return "AVX2";
    return "baseline x86-64-v2";
#else
    return "scalar";
#endif
}

ISOF_API int32_t isof_oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------ */
/* Inter-pair SIMD batch aligner (CPU baseline of bench.py).                                   */
/*                                                                                            */
/* Same recurrences, tie-breaks and cigar walk as isof_oracle_sg_align_tb with the default    */
/* table, but 32 pairs advance in lock step in int16 lanes (the innermost loop runs over the   */
/* lanes and is auto-vectorised: AVX-512BW / AVX2 where the CPU has them, via target_clones).  This is the   */
/* shape of a production CPU aligner (libparasail is SIMD too); the scalar function above      */
/* stays the checker and tests/test_oracle_golden.py holds the two bit-identical.              */
/* Falls back to the scalar routine for pairs outside the int16 range.                         */
/* ------------------------------------------------------------------------------------------ */
#define VL 32
#define NEG16V (-16000)

typedef struct {
    int32_t n[VL], m[VL];
    const uint8_t *s1[VL], *s2[VL];
    int64_t pair[VL];
    int lanes;
} group_t;

#if defined(__x86_64__) && defined(__GNUC__)
#define ISOF_CLONES __attribute__((target_clones("arch=skylake-avx512", "avx2", "default")))
#else
#define ISOF_CLONES
#endif

ISOF_CLONES
static void group_fill(int N, int M, const int16_t *restrict q1 /* N x VL */, const int16_t *restrict c2 /* M x VL */,
                       int16_t *restrict H /* (M+1) x VL */, int16_t *restrict F /* M x VL */, uint8_t *restrict trace,
                       int16_t match, int16_t mismatch, int16_t open, int16_t ext, const int32_t *restrict nn,
                       const int32_t *restrict mm, int32_t *restrict best_col, int32_t *restrict best_col_i,
                       int32_t *restrict best_row, int32_t *restrict best_row_j)
{
    for (int j = 0; j <= M; ++j) for (int v = 0; v < VL; ++v) H[(size_t)j * VL + v] = 0;
    for (int j = 0; j < M; ++j) for (int v = 0; v < VL; ++v) F[(size_t)j * VL + v] = NEG16V;
    for (int i = 0; i < N; ++i) {
        int16_t Hdiag[VL], Hleft[VL], Eleft[VL];
        const int16_t *restrict qc = q1 + (size_t)i * VL;
        for (int v = 0; v < VL; ++v) { Hdiag[v] = 0; Hleft[v] = 0; Eleft[v] = NEG16V; }
        uint8_t *restrict trow = trace + (size_t)i * M * VL;
        for (int j = 0; j < M; ++j) {
            int16_t *restrict Hup = H + (size_t)(j + 1) * VL;
            int16_t *restrict Fp = F + (size_t)j * VL;
            const int16_t *restrict rc = c2 + (size_t)j * VL;
            uint8_t *restrict t = trow + (size_t)j * VL;
#pragma omp simd
            for (int v = 0; v < VL; ++v) {
                const int16_t eo = (int16_t)(Hleft[v] - open), ee = (int16_t)(Eleft[v] - ext);
                const int16_t E = eo > ee ? eo : ee;
                const uint8_t fe = eo > ee ? 0 : T_E_EXT;
                const int16_t fo = (int16_t)(Hup[v] - open), fx = (int16_t)(Fp[v] - ext);
                const int16_t Fv = fo > fx ? fo : fx;
                const uint8_t ff = fo > fx ? 0 : T_F_EXT;
                const int16_t wild = (int16_t)((qc[v] > 3) | (rc[v] > 3));
                const int16_t sm = qc[v] == rc[v] ? match : mismatch;
                const int16_t s = wild ? 0 : sm;
                const int16_t Hd = (int16_t)(Hdiag[v] + s);
                int16_t Hv = Hd > E ? Hd : E;
                Hv = Hv > Fv ? Hv : Fv;
                const uint8_t src = (Hv == Hd) ? (T_SRC_DIAG << 2) : ((Hv == Fv) ? (T_SRC_F << 2) : (T_SRC_E << 2));
                t[v] = (uint8_t)(src | ff | fe);
                Hdiag[v] = Hup[v];
                Hup[v] = Hv;
                Fp[v] = Fv;
                Hleft[v] = Hv;
                Eleft[v] = E;
            }
        }
        for (int v = 0; v < VL; ++v) {
            if (i < nn[v]) {                                   /* last column of lane v */
                const int32_t hv = H[(size_t)mm[v] * VL + v];
                if (hv > best_col[v]) { best_col[v] = hv; best_col_i[v] = i; }
            }
            if (i == nn[v] - 1) {                              /* last row of lane v */
                for (int j = 0; j < mm[v]; ++j) {
                    const int32_t hv = H[(size_t)(j + 1) * VL + v];
                    if (hv > best_row[v]) { best_row[v] = hv; best_row_j[v] = j; }
                }
            }
        }
    }
}

static int simd_ok(int32_t n, int32_t m, int32_t match, int32_t mismatch, int32_t open, int32_t ext)
{
    int64_t mn = n < m ? n : m, mx = n > m ? n : m;
    int64_t am = match < 0 ? -match : match, ax = mismatch < 0 ? -mismatch : mismatch;
    if (am * mn > 30000) return 0;
    if (2 * (int64_t)open + (int64_t)ext * mx + ax + am > 15000) return 0;
    return n > 0 && m > 0;
}

ISOF_API int64_t isof_oracle_sg_align_pairs_simd(const uint8_t *cat, const int64_t *off, const int32_t *ia,
                                                 const int32_t *ib, int64_t P, int32_t match, int32_t mismatch,
                                                 int32_t open, int32_t ext, int32_t *score, int32_t *end_q,
                                                 int32_t *end_r, uint32_t *cigar, int64_t *cigar_off,
                                                 int64_t cigar_cap, int threads)
{
    uint32_t **runs = (uint32_t **)calloc((size_t)P, sizeof(uint32_t *));
    int32_t *nr = (int32_t *)calloc((size_t)P, sizeof(int32_t));
    int64_t *order = (int64_t *)malloc(sizeof(int64_t) * (size_t)(P > 0 ? P : 1));
    int err = 0;
    /* SIMD-eligible pairs first, sorted by (n, m) so that lanes of a group have similar shapes */
    int64_t ne = 0;
    for (int64_t p = 0; p < P; ++p) {
        int32_t n = (int32_t)(off[ia[p] + 1] - off[ia[p]]), m = (int32_t)(off[ib[p] + 1] - off[ib[p]]);
        if (simd_ok(n, m, match, mismatch, open, ext)) order[ne++] = p;
    }
    /* insertion into buckets would be overkill: a simple shell sort on n*65536+m */
    for (int64_t gap = ne / 2; gap > 0; gap /= 2)
        for (int64_t a = gap; a < ne; ++a) {
            int64_t t = order[a], b = a;
            int64_t kt = ((off[ia[t] + 1] - off[ia[t]]) << 20) + (off[ib[t] + 1] - off[ib[t]]);
            while (b >= gap) {
                int64_t u = order[b - gap];
                int64_t ku = ((off[ia[u] + 1] - off[ia[u]]) << 20) + (off[ib[u] + 1] - off[ib[u]]);
                if (ku <= kt) break;
                order[b] = u; b -= gap;
            }
            order[b] = t;
        }
    const int64_t ngroups = (ne + VL - 1) / VL;
    (void)threads;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int64_t g = 0; g < ngroups; ++g) {
        int32_t nn[VL], mm[VL];
        int64_t pp[VL];
        int lanes = 0, N = 0, M = 0;
        for (int v = 0; v < VL; ++v) {
            int64_t idx = g * VL + v;
            if (idx < ne) {
                int64_t p = order[idx];
                pp[v] = p; nn[v] = (int32_t)(off[ia[p] + 1] - off[ia[p]]); mm[v] = (int32_t)(off[ib[p] + 1] - off[ib[p]]);
                if (nn[v] > N) N = nn[v];
                if (mm[v] > M) M = mm[v];
                ++lanes;
            } else { pp[v] = -1; nn[v] = 0; mm[v] = 0; }
        }
        /* per-thread grow-only scratch: a fresh 140 MB trace per group would spend a third of the time in page faults */
        static _Thread_local int16_t *q1 = NULL, *c2 = NULL, *H = NULL, *F = NULL;
        static _Thread_local uint8_t *trace = NULL;
        static _Thread_local size_t capN = 0, capM = 0, capT = 0;
        if ((size_t)N > capN) { free(q1); q1 = (int16_t *)malloc(sizeof(int16_t) * (size_t)N * VL); capN = q1 ? (size_t)N : 0; }
        if ((size_t)M > capM) {
            free(c2); free(H); free(F);
            c2 = (int16_t *)malloc(sizeof(int16_t) * (size_t)M * VL);
            H = (int16_t *)malloc(sizeof(int16_t) * (size_t)(M + 1) * VL);
            F = (int16_t *)malloc(sizeof(int16_t) * (size_t)M * VL);
            capM = (c2 && H && F) ? (size_t)M : 0;
        }
        if ((size_t)N * (size_t)M > capT) { free(trace); trace = (uint8_t *)malloc((size_t)N * (size_t)M * VL); capT = trace ? (size_t)N * (size_t)M : 0; }
        if (!capN || !capM || !capT) { err = 1; continue; }
        for (int v = 0; v < VL; ++v) {
            const uint8_t *a = pp[v] >= 0 ? cat + off[ia[pp[v]]] : NULL;
            const uint8_t *b = pp[v] >= 0 ? cat + off[ib[pp[v]]] : NULL;
            for (int i = 0; i < N; ++i) q1[(size_t)i * VL + v] = (int16_t)(i < nn[v] ? base_code(a[i]) : 6);
            for (int j = 0; j < M; ++j) c2[(size_t)j * VL + v] = (int16_t)(j < mm[v] ? base_code(b[j]) : 7);
        }
        int32_t bc[VL], bci[VL], br[VL], brj[VL];
        for (int v = 0; v < VL; ++v) { bc[v] = br[v] = -(1 << 30); bci[v] = brj[v] = 0; }
        group_fill(N, M, q1, c2, H, F, trace, (int16_t)match, (int16_t)mismatch, (int16_t)open, (int16_t)ext, nn, mm, bc, bci,
                   br, brj);
        for (int v = 0; v < VL; ++v) {
            if (pp[v] < 0) continue;
            const int64_t p = pp[v];
            const int32_t n = nn[v], m = mm[v];
            const uint8_t *s1 = cat + off[ia[p]], *s2 = cat + off[ib[p]];
            int32_t sc = br[v], eq = n - 1, er = brj[v];
            if (bc[v] > sc) { sc = bc[v]; eq = bci[v]; er = m - 1; }
            else if (bc[v] == sc && er == m - 1 && bci[v] < eq) eq = bci[v];
            score[p] = sc; end_q[p] = eq; end_r[p] = er;
            uint32_t *rev = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n + m + 2));
            int32_t cnt = 0;
            uint32_t cur_op = 0xffffffffu, cur_len = 0;
#define PUSH2(op, len)                                                                             \
    do {                                                                                           \
        if ((len) > 0) {                                                                           \
            if (cur_op == (uint32_t)(op)) cur_len += (uint32_t)(len);                              \
            else { if (cur_len) rev[cnt++] = (cur_len << 4) | cur_op; cur_op = (op); cur_len = (len); } \
        }                                                                                          \
    } while (0)
            int32_t i = eq, j = er;
            if (eq + 1 == n) PUSH2(OP_D, m - 1 - j); else PUSH2(OP_I, n - 1 - i);
            int where = T_SRC_DIAG;
            while (i >= 0 || j >= 0) {
                if (i < 0) { PUSH2(OP_D, j + 1); break; }
                if (j < 0) { PUSH2(OP_I, i + 1); break; }
                const uint8_t t = trace[((size_t)i * M + (size_t)j) * VL + v];
                if (where == T_SRC_DIAG) {
                    int src = t >> 2;
                    if (src == T_SRC_DIAG) { PUSH2(upc(s1[i]) == upc(s2[j]) ? OP_EQ : OP_X, 1); --i; --j; }
                    else where = src;
                } else if (where == T_SRC_E) { PUSH2(OP_D, 1); --j; where = (t & T_E_EXT) ? T_SRC_E : T_SRC_DIAG; }
                else { PUSH2(OP_I, 1); --i; where = (t & T_F_EXT) ? T_SRC_F : T_SRC_DIAG; }
            }
            if (cur_len) rev[cnt++] = (cur_len << 4) | cur_op;
#undef PUSH2
            runs[p] = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(cnt > 0 ? cnt : 1));
            for (int q = 0; q < cnt; ++q) runs[p][q] = rev[cnt - 1 - q];
            nr[p] = cnt;
            free(rev);
        }
    }
    /* everything the int16 lanes cannot hold goes through the scalar routine */
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int64_t p = 0; p < P; ++p) {
        if (runs[p]) continue;
        int32_t n = (int32_t)(off[ia[p] + 1] - off[ia[p]]), m = (int32_t)(off[ib[p] + 1] - off[ib[p]]);
        int32_t cap = n + m + 2;
        runs[p] = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(cap > 0 ? cap : 1));
        int32_t rc = isof_oracle_sg_align(cat + off[ia[p]], n, cat + off[ib[p]], m, match, mismatch, open, ext, &score[p],
                                          &end_q[p], &end_r[p], runs[p], cap, &nr[p]);
        if (rc < 0) err = 1;
    }
    int64_t tot = 0;
    for (int64_t p = 0; p < P; ++p) {
        cigar_off[p] = tot;
        if (!err && cigar && tot + nr[p] <= cigar_cap) memcpy(cigar + tot, runs[p], sizeof(uint32_t) * (size_t)nr[p]);
        tot += nr[p];
        free(runs[p]);
    }
    cigar_off[P] = tot;
    free(runs); free(nr); free(order);
    return err ? -1 : tot;
}

/* ---------------------------------------------------------------------------------------------
 * Global (Needleman-Wunsch, unit cost) edit distance by Myers' bit-vector algorithm in Hyyro's
 * block form -- the "edlib CPU" column of BASELINE config 5.  edlib is listed by the reference's
 * setup.py:81 but has ZERO call sites, so there is nothing to be exact against; this is
 * a restatement of the published algorithm (G. Myers, JACM 46(3) 1999; H. Hyyro, Nordic J.
 * Computing 10(1) 2003), checked in tests/ against a plain O(nm) DP.  It is a timing baseline
 * only: nothing in the product computes edit distances.
 *
 * The query is cut into blocks of 64 rows; per text character each block is advanced with the
 * horizontal delta (-1, 0, +1) of the block above as carry-in.  Row 0 of the DP (global mode)
 * grows by 1 per column, so the first block's carry-in is always +1.
 * ------------------------------------------------------------------------------------------- */
static inline int myers_block(uint64_t *pv_io, uint64_t *mv_io, uint64_t eq, int hin)
{
    uint64_t pv = *pv_io, mv = *mv_io;
    const uint64_t hneg = (uint64_t)(hin < 0);
    const uint64_t xv = eq | mv;
    eq |= hneg;
    const uint64_t xh = (((eq & pv) + pv) ^ pv) | eq;
    uint64_t ph = mv | ~(xh | pv);
    uint64_t mh = pv & xh;
    int hout = 0;
    if (ph >> 63) hout = 1;
    else if (mh >> 63) hout = -1;
    ph = (ph << 1) | (uint64_t)(hin > 0);
    mh = (mh << 1) | hneg;
    *pv_io = mh | ~(xv | ph);
    *mv_io = ph & xv;
    return hout;
}

ISOF_API int32_t isof_oracle_edit_distance(const uint8_t *q, int32_t n, const uint8_t *t, int32_t m)
{
    if (n == 0) return m;
    if (m == 0) return n;
    const int nb = (n + 63) / 64;
    uint64_t *peq = (uint64_t *)calloc((size_t)nb * 256, sizeof(uint64_t));   /* [block][char] */
    uint64_t *pv = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)nb);
    uint64_t *mv = (uint64_t *)calloc((size_t)nb, sizeof(uint64_t));
    if (!peq || !pv || !mv) { free(peq); free(pv); free(mv); return -1; }
    for (int32_t i = 0; i < n; ++i) peq[(size_t)(i >> 6) * 256 + q[i]] |= 1ull << (i & 63);
    for (int b = 0; b < nb; ++b) pv[b] = ~0ull;
    int64_t bottom = (int64_t)nb * 64;           /* D(64*nb, 0): rows below n are padding, never read upward */
    for (int32_t j = 0; j < m; ++j) {
        int h = 1;
        const uint64_t *pe = peq + t[j];
        for (int b = 0; b < nb; ++b) h = myers_block(&pv[b], &mv[b], pe[(size_t)b * 256], h);
        bottom += h;
    }
    /* walk the last column up from row 64*nb to row n with the vertical deltas of the last block */
    for (int r = 63; r >= ((n - 1) & 63) + 1; --r) {
        if ((pv[nb - 1] >> r) & 1) --bottom;
        else if ((mv[nb - 1] >> r) & 1) ++bottom;
    }
    free(peq); free(pv); free(mv);
    return (int32_t)bottom;
}

/* ---------------------------------------------------------------------------------------------
 * One optimal global unit-cost alignment path as (len << 4 | op) runs in alignment order, ops 7 '=' / 8 'X' /
 * 1 'I' (consumes q) / 2 'D' (consumes t) -- the extended CIGAR of edlib's task="path".  Plain O(nm) matrix
 * (Wagner-Fischer), walked back from (n, m) preferring the diagonal, then up ('I'), then left ('D').
 * edlib has no call site in the reference (setup.py:81 only) and documents no tie rule: the checker for
 * isof_edit_distance_pairs' path.  Returns the number of runs (also when it exceeds cap; nothing is
 * written beyond cap), -1 on allocation failure; *dist_out receives D(n, m).
 * ------------------------------------------------------------------------------------------- */
ISOF_API int64_t isof_oracle_edit_path(const uint8_t *q, int32_t n, const uint8_t *t, int32_t m, uint32_t *runs,
                                       int64_t cap, int32_t *dist_out)
{
    const size_t W = (size_t)m + 1;
    int32_t *D = (int32_t *)malloc(sizeof(int32_t) * ((size_t)n + 1) * W);
    uint32_t *rev = (uint32_t *)malloc(sizeof(uint32_t) * ((size_t)n + (size_t)m + 2));
    if (!D || !rev) { free(D); free(rev); return -1; }
    for (int32_t j = 0; j <= m; ++j) D[j] = j;
    for (int32_t i = 1; i <= n; ++i) {
        int32_t *row = D + (size_t)i * W;
        const int32_t *up = row - W;
        row[0] = i;
        for (int32_t j = 1; j <= m; ++j) {
            int32_t v = up[j - 1] + (q[i - 1] != t[j - 1]);
            if (up[j] + 1 < v) v = up[j] + 1;
            if (row[j - 1] + 1 < v) v = row[j - 1] + 1;
            row[j] = v;
        }
    }
    if (dist_out) *dist_out = D[(size_t)n * W + m];
    int64_t nr = 0;
    uint32_t cur_op = 0xff, cur_len = 0;
    int32_t i = n, j = m;
    while (i > 0 || j > 0) {
        uint32_t op;
        uint32_t len = 1;
        if (i > 0 && j > 0) {
            const int32_t here = D[(size_t)i * W + j];
            const int neq = q[i - 1] != t[j - 1];
            if (here == D[(size_t)(i - 1) * W + j - 1] + neq) { op = neq ? 8u : 7u; --i; --j; }
            else if (here == D[(size_t)(i - 1) * W + j] + 1) { op = 1u; --i; }
            else { op = 2u; --j; }
        } else if (i > 0) { op = 1u; len = (uint32_t)i; i = 0; }
        else { op = 2u; len = (uint32_t)j; j = 0; }
        if (op == cur_op) cur_len += len;
        else {
            if (cur_len) rev[nr++] = (cur_len << 4) | cur_op;
            cur_op = op;
            cur_len = len;
        }
    }
    if (cur_len) rev[nr++] = (cur_len << 4) | cur_op;
    for (int64_t x = 0; x < nr && x < cap; ++x) runs[x] = rev[nr - 1 - x];
    free(D); free(rev);
    return nr;
}

ISOF_API int64_t isof_oracle_edit_distance_pairs(const uint8_t *cat, const int64_t *off, const int32_t *ia,
                                                 const int32_t *ib, int64_t P, int32_t *dist, int threads)
{
    int err = 0;
    (void)threads;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel for schedule(dynamic, 16)
#endif
    for (int64_t p = 0; p < P; ++p) {
        dist[p] = isof_oracle_edit_distance(cat + off[ia[p]], (int32_t)(off[ia[p] + 1] - off[ia[p]]),
                                            cat + off[ib[p]], (int32_t)(off[ib[p] + 1] - off[ib[p]]));
        if (dist[p] < 0) err = 1;
    }
    return err ? -1 : P;
}
