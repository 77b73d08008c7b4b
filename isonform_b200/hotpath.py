"""Host-side mirror of the reference's hot-path seams, backed by the CUDA library.

Same names, argument meaning and error behaviour as the reference functions they replace
(SURVEY.md section 8b):

    get_kmer_minimizers / get_minimizers_and_positions        main:81-110, 159-171
    parasail_alignment / cigar_to_seq                         modules/consensus.py:13-67
    parse_cigar_diversity                                     modules/SimplifyGraph.py:175-196
    find_first_significant_match, align_to_merge,
    parse_cigar_diversity_isoform_level_new                   modules/IsoformGeneration.py:251-397

plus the batched forms the drivers use (`align_pairs`, `align_to_merge_pairs`,
`bubble_decisions`) and `edit_distance_pairs` / `edlib_align` (edlib's interface: setup.py:81, no call site).  Every numeric result comes from the GPU; the host only evaluates the
reference's float comparisons on the integer features the kernels emit.
"""
import re
import sys

import numpy as np

from ._lib import default_context, pack_sequences, runs_to_string, runs_to_tuples  # noqa: F401 (re-exported)

_CIGAR_SPLIT = re.compile(r"[=DXSMI]+")


# ----------------------------------------------------------------------------------------------
# (a) minimizers
# ----------------------------------------------------------------------------------------------
def minimizer_positions_batch(seqs, k_size, w_size, ctx=None):
    """list of int32 arrays: minimizer start positions per read (GPU)."""
    ctx = ctx or default_context()
    cat, off = pack_sequences(seqs)
    pos, poff = ctx.minimizers_batch(cat, off, k_size, w_size)
    return [pos[poff[i]:poff[i + 1]] for i in range(len(seqs))]


def get_kmer_minimizers(seq, k_size, w_size, ctx=None):
    """[(kmer, pos)] -- drop-in for main:81-110 (k-mer order = CPython hash under PYTHONHASHSEED=0)."""
    pos = minimizer_positions_batch([seq], k_size, w_size, ctx)[0]
    return [(seq[p:p + k_size], int(p)) for p in pos]


def get_minimizers_and_positions(reads, w, k, hash_fcn="lex", ctx=None):
    """{r_id: [(kmer, pos)]} for reads = {r_id: (acc, seq, qual)} -- drop-in for main:159-171.
    One kernel launch for the whole batch."""
    if hash_fcn != "lex":
        # main:405 hard-codes "lex"; get_kmer_maximizers is unreachable from the drivers
        raise NotImplementedError("hash_fcn %r is not reachable from the reference drivers" % (hash_fcn,))
    ids = list(reads)
    seqs = [reads[r][1] for r in ids]
    per_read = minimizer_positions_batch(seqs, k, w, ctx)
    return {r: [(s[p:p + k], int(p)) for p in pos] for r, s, pos in zip(ids, seqs, per_read)}


# ----------------------------------------------------------------------------------------------
# (b) alignment
# ----------------------------------------------------------------------------------------------
def cigar_to_seq(cigar, query, ref):
    """modules/consensus.py:13-52: (query_aln, ref_aln, cigar_tuples); unknown op -> sys.exit()."""
    tuples = []
    pos = 0
    for length in _CIGAR_SPLIT.split(cigar)[:-1]:
        pos += len(length)
        tuples.append((int(length), cigar[pos]))
        pos += 1
    qi = ri = 0
    qa, ra = [], []
    for ln, op in tuples:
        if op == "=" or op == "X":
            qa.append(query[qi:qi + ln]); ra.append(ref[ri:ri + ln]); qi += ln; ri += ln
        elif op == "I":
            ra.append("-" * ln); qa.append(query[qi:qi + ln]); qi += ln
        elif op == "D":
            ra.append(ref[ri:ri + ln]); qa.append("-" * ln); ri += ln
        else:
            sys.exit()
    return "".join(qa), "".join(ra), tuples


def align_pairs(seqs, pair_a, pair_b, match_score=2, mismatch_penalty=-2, opening_penalty=12, gap_ext=1, ctx=None,
                want_cigar=True, features_delta_len=None):
    """Batched parasail_alignment over index pairs into `seqs` (one library call)."""
    ctx = ctx or default_context()
    cat, off = pack_sequences(seqs)         # an empty sequence is an error only if a pair references it (device check)
    return ctx.sg_align_pairs(cat, off, pair_a, pair_b, match_score, mismatch_penalty, opening_penalty, gap_ext,
                              want_cigar=want_cigar, features_delta_len=features_delta_len)


def parasail_alignment(s1, s2, match_score=2, mismatch_penalty=-2, opening_penalty=12, gap_ext=1, ctx=None):
    """Drop-in for consensus.parasail_alignment (modules/consensus.py:55-67):
    (s1_alignment, s2_alignment, cigar_string, cigar_tuples, score)."""
    if len(s1) == 0 or len(s2) == 0:
        raise ValueError("parasail rejects empty sequences")
    ctx = ctx or default_context()
    one = getattr(ctx, "sg_align_one", None)
    if one is not None:                       # lean single-pair call (pre-allocated buffers, fused tiny-call kernel)
        score, _eq, _er, runs = one(s1, s2, match_score, mismatch_penalty, opening_penalty, gap_ext)
    else:
        res = align_pairs([s1, s2], [0], [1], match_score, mismatch_penalty, opening_penalty, gap_ext, ctx)
        score, runs = int(res["score"][0]), res["cigar"]
    cigar_string = runs_to_string(runs)
    a1, a2, tuples = cigar_to_seq(cigar_string, s1, s2)
    return a1, a2, cigar_string, tuples, score


# ----------------------------------------------------------------------------------------------
# (b') unit-cost edit distance -- edlib's shape (declared by the reference's setup.py:81, never called)
# ----------------------------------------------------------------------------------------------
def edit_distance_pairs(seqs, pair_a, pair_b, k=-1, task="distance", ctx=None):
    """Batched edlib.align(seqs[a], seqs[b], mode="NW", task=task, k=k): a list of dicts with edlib's keys
    {"editDistance": int (-1 above k), "cigar": extended CIGAR string or None}.  One library call (Myers/Hyyro
    bit-parallel kernel, csrc/myers.cu)."""
    if task not in ("distance", "path"):
        raise ValueError("task must be 'distance' or 'path'")
    ctx = ctx or default_context()
    cat, off = pack_sequences(seqs)
    res = ctx.edit_distance_pairs(cat, off, pair_a, pair_b, k=k, want_cigar=(task == "path"))
    out = []
    for p, d in enumerate(res["dist"]):
        cig = None
        if task == "path" and d >= 0:
            cig = runs_to_string(res["cigar"][res["cigar_off"][p]:res["cigar_off"][p + 1]])
        out.append({"editDistance": int(d), "cigar": cig})
    return out


def edlib_align(query, target, mode="NW", task="distance", k=-1, ctx=None):
    """One pair with edlib.align's signature (global mode only: the kernel computes NW)."""
    if mode != "NW":
        raise ValueError("only mode='NW' (global) is implemented")
    return edit_distance_pairs([query, target], [0], [1], k=k, task=task, ctx=ctx)[0]


# ----------------------------------------------------------------------------------------------
# decisions: float comparisons of the reference on integer features from the GPU
# ----------------------------------------------------------------------------------------------
def parse_cigar_diversity(cigar_tuples, delta_perc, delta_len):
    """modules/SimplifyGraph.py:175-196 (host form, for callers holding cigar tuples)."""
    miss = alen = 0
    for ln, op in cigar_tuples:
        alen += ln
        if op != "=" and op != "M":
            miss += ln
            if ln > delta_len:
                return False
    return (miss / alen) <= (max(delta_perc * alen, delta_len) / alen)


def _pop_from_features(f, delta_perc, delta_len):
    alen = f[:, 0].astype(np.float64)
    miss = f[:, 2].astype(np.float64)
    ok = f[:, 3] <= delta_len
    with np.errstate(divide="ignore", invalid="ignore"):
        div = miss / alen
        mod = np.maximum(delta_perc * alen, float(delta_len)) / alen
    return ok & (div <= mod)


def _merge_from_features(f, delta, delta_len, delta_iso_len_3, delta_iso_len_5):
    """IsoformGeneration.py:380-397 + 251-320 evaluated with numpy float64 (== Python float)."""
    alen = f[:, 0].astype(np.int64)
    start, end = f[:, 4].astype(np.int64), f[:, 5].astype(np.int64)
    found = (start >= 0) & (end >= 0)
    last = alen - end
    similar = last - start
    ok = found & (f[:, 6] == 0)
    ok &= (f[:, 8].astype(np.int64) + f[:, 9]) < delta_iso_len_5
    ok &= (f[:, 10].astype(np.int64) + f[:, 11]) < delta_iso_len_3
    ok &= similar >= 100
    sim = np.where(ok, similar, 1).astype(np.float64)
    miss = (f[:, 7].astype(np.int64) * delta_len).astype(np.float64)
    d = np.where(alen < 200, 2 * delta, delta)
    div = miss / sim
    mod = np.maximum(d * sim, float(delta_len)) / sim
    return ok & (div <= mod)


def bubble_decisions(seqs, pair_a, pair_b, delta_len, ctx=None):
    """align_bubble_nodes' gates + alignment + parse_cigar_diversity (SimplifyGraph.py:630-664) for
    many consensus pairs at once.  Returns (good_to_pop bool[P], aligned bool[P])."""
    pa = np.asarray(pair_a, dtype=np.int64)
    pb = np.asarray(pair_b, dtype=np.int64)
    lens = np.array([len(s) for s in seqs], dtype=np.int64)
    l1, l2 = lens[pa], lens[pb]
    longer, shorter = np.maximum(l1, l2), np.minimum(l1, l2)
    delta = 0.20
    both_long = (shorter > delta_len) & (longer > delta_len)
    with np.errstate(divide="ignore", invalid="ignore"):
        close = ((longer - shorter) / longer.astype(np.float64)) < delta
    aligned = both_long & close
    both_short = (shorter < delta_len) & (longer < delta_len)
    out = np.where(both_long, False, np.where(both_short, True, (longer - shorter) < delta_len))
    idx = np.nonzero(aligned)[0]
    if len(idx):
        res = align_pairs(seqs, pa[idx], pb[idx], 2, -8, 12, 1, ctx, want_cigar=False, features_delta_len=delta_len)
        out[idx] = _pop_from_features(res["features"], delta, delta_len)
    return out.astype(bool), aligned


def align_to_merge_pairs(seqs, pair_a, pair_b, delta, delta_len, delta_iso_len_3, delta_iso_len_5, ctx=None):
    """align_to_merge (IsoformGeneration.py:380-397) for many (consensus1, consensus2) index pairs."""
    res = align_pairs(seqs, pair_a, pair_b, 2, -2, 12, 1, ctx, want_cigar=False, features_delta_len=delta_len)
    return _merge_from_features(res["features"], delta, delta_len, delta_iso_len_3, delta_iso_len_5)


def align_to_merge(consensus1, consensus2, delta, delta_len, delta_iso_len_3, delta_iso_len_5, ctx=None):
    """Drop-in for IsoformGeneration.align_to_merge (single pair)."""
    return bool(align_to_merge_pairs([consensus1, consensus2], [0], [1], delta, delta_len, delta_iso_len_3,
                                     delta_iso_len_5, ctx)[0])


def find_first_significant_match(s1_alignment, s2_alignment, windowsize, alignment_threshold):
    """IsoformGeneration.py:370-377 (host form; the batched path gets it from the GPU features)."""
    eq = np.frombuffer(s1_alignment.encode("latin-1"), np.uint8) == np.frombuffer(s2_alignment.encode("latin-1"), np.uint8)
    n = len(eq)
    if n - windowsize + 1 <= 0:
        return -1
    c = np.concatenate([[0], np.cumsum(eq)])
    sums = c[windowsize:] - c[:-windowsize]
    hit = np.nonzero(sums / windowsize > alignment_threshold)[0]
    return int(hit[0]) if len(hit) else -1
