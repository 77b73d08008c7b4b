/*
 * isonform_b200.h -- C ABI of libisonform_b200.so, the B200 (sm_100a) implementation of isONform's
 * data-parallel hot path.  Plain pointers and sizes only; no torch / C++ types.
 *
 * The reference (aljpetri/isONform v0.3.9) is pure Python and has no FFI of its own; the seams this
 * library replaces are module-level Python functions.  Each entry point
 * cites the reference interface it stands in for (file:line relative to the reference tree).
 * INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - every function returns 0 on success or a negative isof_status; isof_last_error(ctx) gives text.
 *   - the caller owns every buffer; the library owns only `ctx` (device scratch, one CUDA stream).
 *   - one ctx per (process, GPU); calls on a ctx are stream-ordered and not thread-safe.
 *   - "_dev" variants take DEVICE pointers for all array arguments and do not synchronise the stream
 *     except where documented; the others take HOST pointers, copy in/out and synchronise.
 *   - there is no CPU fallback: without a CUDA device isof_init fails with ISOF_ERR_CUDA.
 */
#ifndef ISONFORM_B200_H
#define ISONFORM_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define ISOF_API __attribute__((visibility("default")))
#else
#define ISOF_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct isof_ctx isof_ctx;

typedef enum {
    ISOF_OK = 0,
    ISOF_ERR_CUDA = -1,       /* CUDA runtime error (text in isof_last_error)                     */
    ISOF_ERR_ARG = -2,        /* bad argument (empty sequence, k > 64*8, w_size < k, ...)         */
    ISOF_ERR_CAPACITY = -3,   /* an output buffer is too small; required size reported            */
    ISOF_ERR_NOMEM = -4,      /* a single work item does not fit the scratch budget               */
    ISOF_ERR_RANGE = -5       /* scores would overflow the 32-bit tagged representation           */
} isof_status;

/* CIGAR run encoding: (length << 4) | op, BAM op codes, as parasail's Cigar.seq. */
enum { ISOF_CIGAR_M = 0, ISOF_CIGAR_I = 1, ISOF_CIGAR_D = 2, ISOF_CIGAR_EQ = 7, ISOF_CIGAR_X = 8 };

/* number of int32 features per pair written by isof_sg_features_* (see below) */
#define ISOF_N_FEATURES 12

/* ------------------------------------------------------------------------------------------ */
/* context                                                                                    */
/* ------------------------------------------------------------------------------------------ */
ISOF_API int isof_version(void);
/* Creates the context on CUDA device `device`.  Fails (ISOF_ERR_CUDA) when no device is present. */
ISOF_API int isof_init(int device, isof_ctx **ctx);
ISOF_API int isof_destroy(isof_ctx *ctx);
ISOF_API const char *isof_last_error(const isof_ctx *ctx);
/* Device scratch is grow-only (a call that needed a 50 GB trace keeps it for the next one).  This returns all of it to the
 * device -- after synchronising the ctx's streams -- e.g. between the stages of an in-process run that shares the GPU. */
ISOF_API int isof_release_scratch(isof_ctx *ctx);
/* Use an existing CUDA stream (cudaStream_t passed as void*), e.g. torch's current stream, so that
 * the caller's CUDA events bracket the library's kernels.  NULL restores the ctx-owned stream. */
ISOF_API int isof_set_stream(isof_ctx *ctx, void *cuda_stream);
/* Upper bound for the alignment trace scratch in bytes (default: 40% of the device memory free at first use, up to 55% for batches of very long pairs). */
ISOF_API int isof_set_trace_budget(isof_ctx *ctx, size_t bytes);
/* The DP kernel has two arithmetic paths with identical results: a packed 16-bit path (two alignments
 * per warp, used when both fit the 16-bit tagged range and hold only ACGT) and a 32-bit path (any
 * length, wildcards).  on=0 forces the 32-bit path (A/B measurements, tests); default on=1. */
ISOF_API int isof_set_packed(isof_ctx *ctx, int on);
/* Long pairs (beyond the plain 16-bit range, e.g. 10 kb reads) are occupancy-bound by trace memory: 37 MB per 8.5 kb pair.
 * In feature-only calls their trace is therefore BANDED: only the diagonals between the main one and the one the free end
 * gaps make room for (j - i between min(0, m-n) and max(0, m-n)), widened by `half_width` either side, are stored.  The DP
 * itself is complete (scores are exact); a pair whose optimal path leaves the band is detected by the traceback and aligned
 * again with the full trace, so results never depend on the band.  half_width = 0 turns banding off; default 1024. */
ISOF_API int isof_set_band(isof_ctx *ctx, int half_width);
/* Traceback of the regular (chunked) path.  mode 1 = one WARP per pair -- the trace holds the decision of every cell, so
 * the 32 lanes look ahead along the diagonal (state H) or along the gap and one iteration covers up to 32 steps of the
 * walk: 2x faster on its own, but its many warps take the slots the next chunk's DP waits for; mode 0 = one thread per
 * pair (light, latency-bound: fits the gaps of a pipelined step); mode 2 (default) = warp per pair for chunks of up to 4096
 * pairs (small calls, long reads), thread per pair beyond.  Same results in every mode.  The path walk of
 * isof_edit_distance_pairs follows the same switch (warp per pair for pairs beyond 128 rows; mode 2: while the call holds
 * at most 128 x SMs of them). */
ISOF_API int isof_set_traceback_warp(isof_ctx *ctx, int mode);
/* Host-buffer calls of at most 32 small pairs (trace <= 200 KB each, i.e. up to ~500 x 500) run as ONE fused launch: base
 * codes, DP, traceback and features in one kernel with the trace in shared memory and inputs / outputs in mapped pinned
 * memory -- the per-pair seams of the reference (a memo miss of consensus.parasail_alignment).  Identical results; on=0
 * sends such calls through the regular path (tests, A/B measurements); default on=1. */
ISOF_API int isof_set_fused_small(isof_ctx *ctx, int on);
/* Latency accounting of the last fused small call (development aid), 56 values: [0..3] SM clock cycles pair 0 spent
 * {fetching its inputs over PCIe + encoding, in the DP, in the traceback, in features + result stores}; [8 + 2g], [9 + 2g]
 * = begin / end of stripe g's DP in ns (%globaltimer) when the pair ran in chase mode (one block per stripe). */
ISOF_API int isof_debug_fused_phases(isof_ctx *ctx, int64_t *values56);
/* A call whose pairs are ALL short (both lengths <= 384, inside the 16-bit range) and numerous enough to fill the GPU
 * (about 64 x SMs x (length/60)^2 pairs: a thread per task trades latency for throughput) takes the short-pair regime:
 * one thread per task instead of a 32-lane wavefront per task (the bubble-popping sizes, SimplifyGraph.py:630-664).
 * Identical results.  on=0 never, on=1 automatic (default), on=2 every all-short call (tests, A/B measurements). */
ISOF_API int isof_set_short(isof_ctx *ctx, int on);
/* Tie-break table of the aligner.  parasail is third-party and not vendored by the reference, so every
 * tie-break of its sg_trace + cigar code (consensus.py:56-59) is a named switch here and in the oracle
 * (isof_tiebreak_t in oracle/isof_oracle.c) instead of a hard-coded constant; all zero (the default) is the
 * library as restated in DESIGN.md.  h_e_before_f: H source on ties DIAG > E > F instead of DIAG > F > E.
 * gap_open_on_tie: a gap opens instead of extending when both score the same.  end_col_first: the end cell is
 * searched in the last column first, then the last row.  swap_id: the CIGAR letters of the two gap states are
 * swapped ('I' consumes s2).  Results are bit-identical to the oracle under each of the 16 tables. */
ISOF_API int isof_set_tiebreak(isof_ctx *ctx, int h_e_before_f, int gap_open_on_tie, int end_col_first, int swap_id);
/* Alignment batches whose trace exceeds the scratch budget are cut into chunks.  on=1 (default):
 * chunks rotate over three streams, each with its own scratch, so a chunk's tail (idle SMs while
 * its last pairs finish) and its traceback overlap the next chunks' DP.  on=0: one stream, one
 * buffer (per-kernel event timings are then exact; bench.py uses it for the roofline step). */
ISOF_API int isof_set_overlap(isof_ctx *ctx, int on);
/* Debug build only (libisonform_b200_dbg.so, -DISOF_CHECK): every kernel carries device-side bounds
 * assertions (trace / boundary / sequence / output indices, CIGAR-tail vs unread-trace overlap) that record a
 * bit per violated class instead of trapping.  Returns and clears the mask; bit 31 is set iff the assertions
 * are compiled in (the release library always reports 0). */
ISOF_API int isof_debug_violations(isof_ctx *ctx, unsigned int *mask);
/* Test knob of kernel (c): k-mers that do not pack into 2 bits per base (non-ACGT, lower case, k >= 32) are brought
 * together by a 63-bit hash and then told apart by comparing bytes.  A small mask (e.g. 0xF) makes nearly all of them
 * collide, so a test can prove that identity never depends on the hash.  Default: all 63 bits. */
ISOF_API int isof_debug_set_span_hash_mask(isof_ctx *ctx, uint64_t mask);
/* Launches a kernel whose only assertion is false (class 30): proves the reporting path end to end. */
ISOF_API int isof_debug_selftest(isof_ctx *ctx);

/* Per-ctx counters, reset by isof_profile_reset.  Kernel times are measured with CUDA events on the
 * ctx stream around each launch when profiling is enabled (isof_profile_enable(ctx,1)); enabling
 * it adds one event pair per launch and a synchronise in isof_profile_get. */
typedef struct {
    int64_t launches;            /* kernels launched by this library                            */
    int64_t dp_launches;         /* launches of the alignment DP kernel                          */
    double dp_ms;                /* summed device time of the DP kernel                          */
    double traceback_ms;         /* summed device time of the traceback + cigar kernels          */
    double minimizer_ms;         /* summed device time of the minimizer kernels                  */
    double other_ms;             /* planning / encoding / span kernels                           */
    int64_t dp_cells;            /* sum len(s1)*len(s2) over aligned pairs                       */
    int64_t dp_cells_padded;     /* cells actually updated (stripe padding included)             */
    int64_t dp_cells_packed;     /* part of dp_cells that went through the packed 16-bit path    */
    int64_t dp_cells_profile;    /* part of dp_cells_packed that used the shared-memory query profile */
    int64_t trace_bytes;         /* trace bytes written by the DP kernel                         */
    int64_t h2d_bytes, d2h_bytes;
    int64_t band_pairs_retried;  /* pairs whose path left the banded trace and were aligned again with the full trace */
    int64_t trace_replans;       /* times the trace chunks were planned again, smaller, because a lane did not fit the device */
} isof_profile;
ISOF_API int isof_profile_enable(isof_ctx *ctx, int on);
ISOF_API int isof_profile_reset(isof_ctx *ctx);
ISOF_API int isof_profile_get(isof_ctx *ctx, isof_profile *out);
/* Measures the INT32 ALU-pipe peak of this GPU with a register-only micro-kernel of independent
 * VIADDMNMX (DPX add+max) chains, the instruction the DP kernel is built from: lane-operations per
 * second.  Denominator of the integer roofline reported by bench.py (there is no published or
 * driver-measured INT32 figure).  Takes ~20 ms. */
ISOF_API int isof_int32_peak(isof_ctx *ctx, double *lane_ops_per_s);

/* ------------------------------------------------------------------------------------------ */
/* (a) minimizers                                                                             */
/* ------------------------------------------------------------------------------------------ */
/*
 * Replaces get_minimizers_and_positions / get_kmer_minimizers (main:159-171, main:81-110).
 *   bases      concatenated ASCII reads (any byte values; hashed as CPython hashes str under
 *              PYTHONHASHSEED=0, i.e. SipHash-1-3 with a zero key, main:85,94)
 *   read_off   n_reads+1 offsets into `bases`
 *   k, w_size  the reference's k_size and w_size (window holds w_size-k+1 k-mers)
 *   out_pos    minimizer start positions, reads back to back (the k-mer string is
 *              read[pos:pos+k], rebuilt by the caller exactly like main:89,104,109)
 *   out_off    n_reads+1 offsets into out_pos
 *   n_out      total minimizers; on ISOF_ERR_CAPACITY the size out_pos must have
 */
ISOF_API int isof_minimizers_batch(isof_ctx *ctx, const uint8_t *bases, const int64_t *read_off, int32_t n_reads,
                          int32_t k, int32_t w_size, int32_t *out_pos, int64_t out_cap, int64_t *out_off,
                          int64_t *n_out);
/* Device-pointer variant; *n_out is a HOST pointer (one 8-byte read-back, synchronises). */
ISOF_API int isof_minimizers_batch_dev(isof_ctx *ctx, const uint8_t *bases, const int64_t *read_off, int32_t n_reads,
                              int64_t n_bases, int32_t k, int32_t w_size, int32_t *out_pos, int64_t out_cap,
                              int64_t *out_off, int64_t *n_out);

/* ------------------------------------------------------------------------------------------ */
/* (c) minimizer-pair database + span search                                                  */
/* ------------------------------------------------------------------------------------------ */
/*
 * Replaces, for one batch, minimizers_comb_iterator (main:215-224),
 * get_minimizer_combinations_database (main:174-212) and find_most_supported_span (main:308-338).
 * Input: the reads and their minimizers (min_pos/min_off exactly as isof_minimizers_batch returns).
 * A record g is one (anchor, partner) pair of one read with x_low < p2-p1 <= x_high; records come in
 * the reference's insertion order (reads in batch order, anchors left to right, partners farthest
 * first).  Outputs (arrays of capacity rec_cap; bucket_off needs rec_cap+1):
 *   rec_read (0-based read index; the reference's r_id is rec_read+1), rec_p1, rec_p2
 *   rec_bucket      bucket (= distinct (kmer1,kmer2) pair) of the record
 *   rec_count       find_most_supported_span's instance count for the record: 1 + number of
 *                   OTHER-read records of the bucket with |span difference| < delta_len, or 0 when
 *                   the reference emits no interval (bucket < 3 records, polyA pair, > 3*n_reads)
 *   sorted_rec      record indices grouped by bucket, reference insertion order inside a bucket
 *   bucket_off      n_buckets+1 offsets into sorted_rec;  bucket_dropped: 1 if the reference deletes
 *                   the bucket (polyA/polyA pair, over-abundant)
 * On ISOF_ERR_CAPACITY *n_rec holds the needed rec_cap.
 */
ISOF_API int isof_spans_batch(isof_ctx *ctx, const uint8_t *bases, const int64_t *read_off, int32_t n_reads,
                              const int32_t *min_pos, const int64_t *min_off, int32_t k, int32_t x_low, int32_t x_high,
                              int32_t delta_len, int64_t rec_cap, int32_t *rec_read, int32_t *rec_p1, int32_t *rec_p2,
                              int32_t *rec_count, int32_t *rec_bucket, int32_t *sorted_rec, int32_t *bucket_off,
                              uint8_t *bucket_dropped, int64_t *n_rec, int64_t *n_buckets);
/* Super-batched form for the in-process runner: the reads of n_batches independent batches back to back
 * (batch b owns reads batch_read_off[b] .. batch_read_off[b+1]-1).  Buckets never mix batches and the
 * over-abundance rule (main:206) uses the read count of the bucket's own batch; records of a batch are
 * contiguous (reads are), and so are its buckets.  One launch sequence instead of one per batch. */
ISOF_API int isof_spans_multi(isof_ctx *ctx, const uint8_t *bases, const int64_t *read_off, int32_t n_reads,
                              const int32_t *batch_read_off, int32_t n_batches, const int32_t *min_pos,
                              const int64_t *min_off, int32_t k, int32_t x_low, int32_t x_high, int32_t delta_len,
                              int64_t rec_cap, int32_t *rec_read, int32_t *rec_p1, int32_t *rec_p2, int32_t *rec_count,
                              int32_t *rec_bucket, int32_t *sorted_rec, int32_t *bucket_off, uint8_t *bucket_dropped,
                              int64_t *n_rec, int64_t *n_buckets);
/* Device-pointer variant (n_rec / n_buckets are HOST pointers; two scalar read-backs). */
ISOF_API int isof_spans_batch_dev(isof_ctx *ctx, const uint8_t *bases, const int64_t *read_off, int32_t n_reads,
                                  const int32_t *min_pos, const int64_t *min_off, int64_t n_min, int32_t k, int32_t x_low,
                                  int32_t x_high, int32_t delta_len, int64_t rec_cap, int32_t *rec_read, int32_t *rec_p1,
                                  int32_t *rec_p2, int32_t *rec_count, int32_t *rec_bucket, int32_t *sorted_rec,
                                  int32_t *bucket_off, uint8_t *bucket_dropped, int64_t *n_rec, int64_t *n_buckets);

/*
 * Host-side helpers of the front half (plain C++, no GPU work; SURVEY.md scopes M6 to the host).
 * isof_wis_picks replaces the per-read loop main:496-514 (sort by stop, fill_p2 main:227-239, solve_WIS
 * main:242-263) for ALL reads of a batch: read r owns records read_rec_off[r] .. read_rec_off[r+1]-1 of the
 * isof_spans_batch output; an interval is (rec_p1+k, rec_p2, weight rec_count) for records with rec_count > 0.
 * Same IEEE-754 double arithmetic and tie-breaks as the reference's Python floats.  pick_rec receives the
 * record indices WIS keeps, in increasing stop order per read; pick_off[n_reads+1] the per-read offsets.
 * ISOF_ERR_RANGE: an interval start lies beyond the last stop (the reference raises IndexError there).
 * isof_span_instances writes, for each picked record, the `seqs` array find_most_supported_span builds
 * (main:324-337): (r_id, p1, p2) of the record, then of every other-read record of its bucket whose span differs
 * by less than delta_len -- 3*rec_count values, so inst_off is the prefix sum of 3*rec_count[pick_rec].
 * read_id maps the 0-based read index to the reference's r_id.
 */
ISOF_API int isof_wis_picks(int32_t n_reads, const int64_t *read_rec_off, const int32_t *rec_p1, const int32_t *rec_p2,
                            const int32_t *rec_count, int32_t k, int32_t *pick_rec, int64_t pick_cap, int64_t *pick_off,
                            int64_t *n_picks);
ISOF_API int isof_span_instances(int64_t n_picks, const int32_t *pick_rec, const int32_t *rec_read, const int32_t *rec_p1,
                                 const int32_t *rec_p2, const int32_t *rec_bucket, const int32_t *sorted_rec,
                                 const int32_t *bucket_off, const int64_t *read_id, int32_t delta_len,
                                 const int64_t *inst_off, uint32_t *inst);

/*
 * In-process partial-order-alignment consensus (host C++, csrc/host_poa.cpp): stands where the
 * reference spawns the external `spoa <reads.fa> -l 0 -r 0 -g -2` (IsoformGeneration.run_spoa, IsoformGeneration.py:143-156).
 * Local alignment of each sequence to the growing graph, linear gap, heaviest-bundle consensus; an independent
 * implementation with that command line's scoring (match 5, mismatch -4, gap -2), NOT a bit-exact spoa: use it on both
 * sides of a comparison (scripts/isof_spoa is the same code as an executable).  Sequences are upper-cased; empty ones are
 * skipped.  *out_len receives the consensus length (also on ISOF_ERR_CAPACITY).
 */
ISOF_API int isof_poa_consensus(const uint8_t *seq_cat, const int64_t *seq_off, int32_t n_seqs, int32_t match,
                                int32_t mismatch, int32_t gap, uint8_t *out, int64_t out_cap, int64_t *out_len);

/* ------------------------------------------------------------------------------------------ */
/* (b) semi-global affine alignment with traceback                                            */
/* ------------------------------------------------------------------------------------------ */
/*
 * Replaces consensus.parasail_alignment (modules/consensus.py:55-67): parasail.matrix_create("ACGT",
 * match, mismatch) + parasail.sg_trace_scan_16/_32(s1, s2, open, ext) + result.cigar, for a batch of
 * pairs drawn from a sequence pool.  s1 = pool[pair_a[p]] is the query (CIGAR 'I' consumes it),
 * s2 = pool[pair_b[p]] the reference ('D' consumes it).  Call sites in the reference:
 * SimplifyGraph.py:644 (2,-8,12,1), IsoformGeneration.py:381 (2,-2,12,1), reached again from
 * batch_merging_parallel.py:193.
 *   seq_cat/seq_off   concatenated ASCII sequences and n_seqs+1 offsets
 *   score,end_q,end_r per pair: score, 0-based end cell (parasail's end_query / end_ref)
 *   cigar_runs        (len<<4|op) runs of all pairs back to back, alignment order, end gaps included
 *   cigar_off         n_pairs+1 offsets into cigar_runs (filled even on ISOF_ERR_CAPACITY)
 *   cigar_runs may be NULL (cigar_cap 0) when only scores/end cells are wanted: cigar_off is still filled.
 * Empty sequences are rejected (ISOF_ERR_ARG), as parasail does.
 */
ISOF_API int isof_sg_align_pairs(isof_ctx *ctx, const uint8_t *seq_cat, const int64_t *seq_off, int32_t n_seqs,
                        const int32_t *pair_a, const int32_t *pair_b, int64_t n_pairs, int32_t match,
                        int32_t mismatch, int32_t gap_open, int32_t gap_ext, int32_t *score, int32_t *end_q,
                        int32_t *end_r, uint32_t *cigar_runs, int64_t cigar_cap, int64_t *cigar_off);
/* Device-pointer variant; `n_bases` = seq_off[n_seqs].  *n_runs_total is a HOST pointer. */
ISOF_API int isof_sg_align_pairs_dev(isof_ctx *ctx, const uint8_t *seq_cat, const int64_t *seq_off, int32_t n_seqs,
                            int64_t n_bases, const int32_t *pair_a, const int32_t *pair_b, int64_t n_pairs,
                            int32_t match, int32_t mismatch, int32_t gap_open, int32_t gap_ext, int32_t *score,
                            int32_t *end_q, int32_t *end_r, uint32_t *cigar_runs, int64_t cigar_cap,
                            int64_t *cigar_off, int32_t *features, int32_t delta_len, int64_t *n_runs_total);
/*
 * The SURVEY.md section 8b form: two concatenations, pair p = (s1[p], s2[p]).
 */
ISOF_API int isof_sg_align_batch(isof_ctx *ctx, const uint8_t *s1_cat, const int64_t *s1_off, const uint8_t *s2_cat,
                        const int64_t *s2_off, int32_t n_pairs, int32_t match, int32_t mismatch, int32_t gap_open,
                        int32_t gap_ext, int32_t *score, int32_t *end_q, int32_t *end_r, uint32_t *cigar_runs,
                        int64_t cigar_cap, int64_t *cigar_off);
/*
 * Fused epilogue: alignment + the integer features the reference derives from the CIGAR and the
 * aligned strings to decide bubble popping (parse_cigar_diversity, SimplifyGraph.py:175-196) and
 * isoform merging (find_first_significant_match / parse_cigar_diversity_isoform_level_new,
 * IsoformGeneration.py:251-320,370-397).  The float comparisons stay on the host (python), fed by
 * these integers.  features[p*ISOF_N_FEATURES + ...]:
 *   0 alen          alignment columns            1 n_runs
 *   2 nonmatch      sum of non-'=' run lengths   3 max_nonmatch_run
 *   4 start_match   first column of a 30-column window with >=22 equal columns, -1 if none
 *   5 end_match     the same on the reversed alignment, -1 if none
 *   6 structural    1 if a non-match run longer than delta_len lies between first and last match
 *   7 miss_runs     number of non-match runs (<= delta_len) between first and last match
 *   8 before_match  9 before_nomatch  10 after_match  11 after_nomatch   (IsoformGeneration.py:266-289)
 * cigar_runs/cigar_off may be NULL when only features are wanted.
 */
ISOF_API int isof_sg_features_pairs(isof_ctx *ctx, const uint8_t *seq_cat, const int64_t *seq_off, int32_t n_seqs,
                           const int32_t *pair_a, const int32_t *pair_b, int64_t n_pairs, int32_t match,
                           int32_t mismatch, int32_t gap_open, int32_t gap_ext, int32_t delta_len, int32_t *score,
                           int32_t *features, uint32_t *cigar_runs, int64_t cigar_cap, int64_t *cigar_off);

/* ------------------------------------------------------------------------------------------ */
/* (b') unit-cost global edit distance, Myers / Hyyro bit-parallel                            */
/* ------------------------------------------------------------------------------------------ */
/*
 * What edlib.align(query, target, mode="NW", task="distance" | "path", k=k) returns, for a batch of pairs drawn from a
 * sequence pool.  The reference declares edlib (setup.py:81, README.md:25) and has no call site for it (SURVEY.md section
 * 0.2): this entry point serves BASELINE config 5 (interval-pair sweep against a CPU Myers) and north_star's kernel (b);
 * the reference's decisions stay with isof_sg_* above.  query = pool[pair_a[p]] (rows), target = pool[pair_b[p]].
 * Bytes compare by value (case-sensitive, like edlib without additionalEqualities); empty sequences are allowed.
 *   k            < 0: no bound; else a distance above k is reported as -1 (and the pair gets no path), as edlib does.  In
 *                distance-only calls long pairs are then computed inside the band | i - j | <= k only (Ukkonen), exactly
 *   dist         per pair
 *   cigar_off    NULL: distances only (no trace is stored).  Else n_pairs+1 offsets into cigar_runs (filled even on
 *                ISOF_ERR_CAPACITY; cigar_runs may then be NULL to learn the size)
 *   cigar_runs   (len<<4|op) runs in alignment order with ops ISOF_CIGAR_EQ / _X / _I (consumes the query) / _D (consumes
 *                the target): edlib's extended CIGAR.  One optimal path; ties prefer the diagonal, then I, then D, walking
 *                back from the end (edlib documents no tie rule; the distance is the contract, the path is one witness).
 */
ISOF_API int isof_edit_distance_pairs(isof_ctx *ctx, const uint8_t *seq_cat, const int64_t *seq_off, int32_t n_seqs,
                                      const int32_t *pair_a, const int32_t *pair_b, int64_t n_pairs, int32_t k, int32_t *dist,
                                      uint32_t *cigar_runs, int64_t cigar_cap, int64_t *cigar_off);
/* Device-pointer variant; `n_bases` = seq_off[n_seqs]; *n_runs_total is a HOST pointer (may be NULL). */
ISOF_API int isof_edit_distance_pairs_dev(isof_ctx *ctx, const uint8_t *seq_cat, const int64_t *seq_off, int32_t n_seqs,
                                          int64_t n_bases, const int32_t *pair_a, const int32_t *pair_b, int64_t n_pairs,
                                          int32_t k, int32_t *dist, uint32_t *cigar_runs, int64_t cigar_cap,
                                          int64_t *cigar_off, int64_t *n_runs_total);

#ifdef __cplusplus
}
#endif
#endif /* ISONFORM_B200_H */
