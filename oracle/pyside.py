"""Python half of the CPU oracle (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

Restates, in our own words, the integer/host logic either side of the two kernels.  Each function
cites the reference lines it follows; tests/test_oracle_vs_reference.py pins them against golden
vectors produced by executing the reference itself (scripts/make_golden.py).
"""
from array import array
from collections import defaultdict
from itertools import groupby

from . import capi


# ----------------------------------------------------------------------------------------------
# M0  remove_read_polyA_ends                                                       main:47-67
# ----------------------------------------------------------------------------------------------
def remove_read_polyA_ends(seq, threshold_len, to_len):
    tail = min(len(seq) // 2, 100)
    if tail == 0:
        # seq[:-0] is '' and seq[-0:] is the whole string in Python: the reference therefore scans
        # the entire (0- or 1-character) read.
        head, rest = "", seq
    else:
        head, rest = seq[:-tail], seq[-tail:]
    parts = [head]
    for ch, grp in groupby(rest):
        run = sum(1 for _ in grp)
        if run > threshold_len and ch in ("A", "T"):
            parts.append(ch * to_len)
        else:
            parts.append(ch * run)
    return "".join(parts)


# ----------------------------------------------------------------------------------------------
# M1/M2  minimizers                                                         main:81-110,159-171
# ----------------------------------------------------------------------------------------------
def get_kmer_minimizers(seq, k_size, w_size):
    """[(kmer, pos)] exactly as main:81-110 under PYTHONHASHSEED=0 / CPython>=3.11."""
    pos = capi.minimizers(seq, k_size, w_size)
    return [(seq[p:p + k_size], int(p)) for p in pos]


def get_minimizers_and_positions(reads, w, k, hash_fcn="lex"):
    if hash_fcn != "lex":
        raise NotImplementedError("only the 'lex' order is reachable from the driver (main:405)")
    return {r_id: get_kmer_minimizers(reads[r_id][1], k, w) for r_id in reads}


# ----------------------------------------------------------------------------------------------
# M3  minimizers_comb_iterator                                                    main:215-224
# ----------------------------------------------------------------------------------------------
def minimizers_comb_iterator(minimizers, k, x_low, x_high):
    n = len(minimizers)
    for i in range(n - 1):
        m1, p1 = minimizers[i]
        partners = []
        for j in range(i + 1, n):
            m2, p2 = minimizers[j]
            d = p2 - p1
            if x_low < d <= x_high:
                partners.append((m2, p2))
            elif d > x_high:
                break
        partners.reverse()                      # farthest first
        yield (m1, p1), partners


# ----------------------------------------------------------------------------------------------
# M4  get_minimizer_combinations_database                                         main:174-212
# ----------------------------------------------------------------------------------------------
def get_minimizer_combinations_database(M, k, x_low, x_high, reads):
    db = defaultdict(lambda: defaultdict(lambda: array("I")))
    poly_a = "A" * k
    for r_id in M:
        for (m1, p1), partners in minimizers_comb_iterator(M[r_id], k, x_low, x_high):
            for (m2, p2) in partners:
                if m1 == poly_a and m2 == poly_a:
                    continue
                db[m1][m2].extend((r_id, p1, p2))
    limit = 3 * len(reads)
    for m1 in list(db.keys()):
        for m2 in list(db[m1].keys()):
            if len(db[m1][m2]) <= 3:            # single occurrence
                del db[m1][m2]
            # the reference re-reads db[m1][m2] here, which re-creates an empty array after a
            # delete (main:206); the empty entry behaves like "fewer than 3 occurrences" later.
            if len(db[m1][m2]) // 3 > limit:
                del db[m1][m2]
    return db


# ----------------------------------------------------------------------------------------------
# M5  find_most_supported_span                                                    main:308-338
# ----------------------------------------------------------------------------------------------
def find_most_supported_span(r_id, m1, p1, m1_curr_spans, db, all_intervals, k_size, delta_len):
    for (m2, p2) in m1_curr_spans:
        bucket = db[m1][m2]
        if len(bucket) // 3 < 3:
            continue
        inst = array("I", (r_id, p1, p2))
        span = p2 - p1
        for q in range(0, len(bucket) - 2, 3):
            other, q1, q2 = bucket[q], bucket[q + 1], bucket[q + 2]
            if other == r_id:
                continue
            if abs(span - (q2 - q1)) < delta_len:
                inst.extend((other, q1, q2))
        all_intervals.append((p1 + k_size, p2, len(inst) // 3, inst))


def read_intervals(r_id, minimizers, db, k, x_low, x_high, delta_len):
    """The per-read loop of main:417-480 with --exact (previously_corrected_regions always empty)."""
    out = []
    for (m1, p1), partners in minimizers_comb_iterator(minimizers, k, x_low, x_high):
        if partners:
            find_most_supported_span(r_id, m1, p1, partners, db, out, k, delta_len)
    return out


# ----------------------------------------------------------------------------------------------
# M6  weighted interval scheduling                                               main:227-263
# ----------------------------------------------------------------------------------------------
def solve_WIS(intervals_by_finish):
    n = len(intervals_by_finish)
    last_with_stop = {}
    for idx, (start, stop, _w, _inst) in enumerate(intervals_by_finish):
        if start < stop:
            last_with_stop[stop] = idx
    best_upto = []
    cur = 0
    for x in range(intervals_by_finish[-1][1] + 1):
        if x in last_with_stop:
            cur = last_with_stop[x]
        best_upto.append(cur)
    p = [None] + [best_upto[start] for (start, _s, _w, _i) in intervals_by_finish]
    eps = 0.0001
    v = [None] + [(w - 1) * (stop - start + eps) for (start, stop, w, _i) in intervals_by_finish]
    opt = [0]
    for j in range(1, n + 1):
        opt.append(max(v[j] + opt[p[j]], opt[j - 1]))
    chosen = []
    j = n
    while j > 0:
        if v[j] + opt[p[j]] > opt[j - 1]:
            chosen.append(j - 1)
            j = p[j]
        else:
            j -= 1
    return chosen


# ----------------------------------------------------------------------------------------------
# P1/P2  parasail_alignment + cigar_to_seq                          modules/consensus.py:13-67
# ----------------------------------------------------------------------------------------------
def cigar_to_seq(cigar_tuples, query, ref):
    qi = ri = 0
    qa, ra = [], []
    for ln, op in cigar_tuples:
        if op in ("=", "X"):
            qa.append(query[qi:qi + ln]); ra.append(ref[ri:ri + ln]); qi += ln; ri += ln
        elif op == "I":
            qa.append(query[qi:qi + ln]); ra.append("-" * ln); qi += ln
        elif op == "D":
            qa.append("-" * ln); ra.append(ref[ri:ri + ln]); ri += ln
        else:
            raise SystemExit()
    return "".join(qa), "".join(ra)


def parasail_alignment(s1, s2, match_score=2, mismatch_penalty=-2, opening_penalty=12, gap_ext=1,
                       tiebreak=None):
    score, _eq, _er, runs = capi.sg_align(s1, s2, match_score, mismatch_penalty, opening_penalty,
                                          gap_ext, tiebreak)
    tuples = capi.cigar_runs_to_tuples(runs)
    a1, a2 = cigar_to_seq(tuples, s1, s2)
    return a1, a2, capi.cigar_runs_to_string(runs), tuples, score


# ----------------------------------------------------------------------------------------------
# P3  bubble decision                              modules/SimplifyGraph.py:175-196, 630-664
# ----------------------------------------------------------------------------------------------
def parse_cigar_diversity(cigar_tuples, delta_perc, delta_len):
    miss = 0
    alen = 0
    for ln, op in cigar_tuples:
        alen += ln
        if op != "=" and op != "M":
            miss += ln
            if ln > delta_len:
                return False
    return (miss / alen) <= (max(delta_perc * alen, delta_len) / alen)


def align_bubble_decision(consensus1, consensus2, delta_len, tiebreak=None):
    """Length gates + alignment + decision of align_bubble_nodes (SimplifyGraph.py:630-664).
    Returns (good_to_pop, aligned?)."""
    l1, l2 = len(consensus1), len(consensus2)
    longer, shorter = (l1, l2) if l1 > l2 else (l2, l1)
    delta = 0.20
    if shorter > delta_len and longer > delta_len:
        if (longer - shorter) / longer < delta:
            *_, tuples, _score = parasail_alignment(consensus1, consensus2, 2, -8, 12, 1, tiebreak)
            return parse_cigar_diversity(tuples, delta, delta_len), True
        return False, False
    if shorter < delta_len and longer < delta_len:
        return True, False
    return (longer - shorter) < delta_len, False


# ----------------------------------------------------------------------------------------------
# P4  isoform-merge decision                        modules/IsoformGeneration.py:251-320,363-397
# ----------------------------------------------------------------------------------------------
def find_first_significant_match(a1, a2, windowsize, threshold):
    eq = [1 if x == y else 0 for x, y in zip(a1, a2)]
    for i in range(0, len(eq) - windowsize + 1):
        if sum(eq[i:i + windowsize]) / windowsize > threshold:
            return i
    return -1


def isoform_features(cigar_tuples, delta_len, first_match, last_match):
    """Integer features of parse_cigar_diversity_isoform_level_new, as the GPU epilogue emits them:
    (structural_break, miss_runs, before_match, before_nomatch, after_match, after_nomatch, alen)
    where the reference's miss_match_length == miss_runs * delta_len.  After a structural break
    the reference has already returned; the remaining features keep accumulating here so that
    they are defined for every input."""
    pos = 0
    brk = 0
    miss_runs = 0
    bm = bn = am = an = 0
    for ln, op in cigar_tuples:
        start = pos
        pos += ln
        if op != "=" and op != "M":
            if start < first_match:
                if op != "D":
                    bn += ln
            elif start >= last_match or start + ln >= last_match:
                if op != "D":
                    an += ln
            else:
                if ln > delta_len:
                    brk = 1                      # the reference returns False right here
                else:
                    miss_runs += 1
        else:
            if start < first_match:
                bm += ln
            elif start >= last_match:
                am += ln
    return brk, miss_runs, bm, bn, am, an, pos


def parse_cigar_diversity_isoform_level_new(cigar_tuples, delta, delta_len, delta_iso_len_3,
                                            delta_iso_len_5, overall_len, first_match, last_match):
    brk, miss_runs, bm, bn, am, an, _ = isoform_features(cigar_tuples, delta_len, first_match, last_match)
    if brk:
        return False
    if not (bm + bn < delta_iso_len_5) or not (am + an < delta_iso_len_3):
        return False
    similar = last_match - first_match
    if similar < 100:
        return False
    miss = miss_runs * delta_len
    if overall_len < 200:
        delta = 2 * delta
    return (miss / similar) <= (max(delta * similar, delta_len) / similar)


def align_to_merge(consensus1, consensus2, delta, delta_len, delta_iso_len_3, delta_iso_len_5,
                   tiebreak=None):
    a1, a2, _cig, tuples, _score = parasail_alignment(consensus1, consensus2, 2, -2, 12, 1, tiebreak)
    overall = sum(ln for ln, _ in tuples)
    start_match = find_first_significant_match(a1, a2, 30, 0.7)
    end_match = find_first_significant_match(a1[::-1], a2[::-1], 30, 0.7)
    if start_match < 0 or end_match < 0:
        return False
    return parse_cigar_diversity_isoform_level_new(tuples, delta, delta_len, delta_iso_len_3,
                                                   delta_iso_len_5, overall, start_match,
                                                   overall - end_match)
