"""In-process hot path of one batch, in the order reference `main()` runs it (main:341-514):
polyA trim -> minimizers (GPU) -> pair database + span search (GPU) -> WIS (host) -> the
`all_intervals_for_graph` / `new_all_reads` structures GraphGeneration.generateGraphfromIntervals takes
(main:531).  The graph, bubble bookkeeping and spoa stay in the reference (out of scope);
their alignment calls come back through `merge.AlignmentMemo`.
"""
from . import hotpath, polya, spans


def read_batch(records):
    """[(acc, seq, qual)] -> {r_id: (acc, polyA-trimmed seq, qual)} with r_id = 1..N (main:351)."""
    return {i + 1: (acc, polya.remove_read_polyA_ends(seq, 12, 1), qual) for i, (acc, seq, qual) in enumerate(records)}


def batch_front_half(reads, k=20, w=31, xmin=40, xmax=80, delta_len=5, ctx=None):
    """Returns (all_intervals_for_graph, new_all_reads, skipped_read_ids, SpanIndex)."""
    xmin = max(xmin, 2 * k)                                      # main:623-625
    M = hotpath.get_minimizers_and_positions(reads, w, k, "lex", ctx)
    index = spans.SpanIndex(reads, M, k, xmin, xmax, delta_len, ctx)
    all_intervals_for_graph, new_all_reads, skipped = spans.batch_intervals_for_graph(index, reads)
    return all_intervals_for_graph, new_all_reads, skipped, index


def batches_front_half(batches, k=20, w=31, xmin=40, xmax=80, delta_len=5, ctx=None):
    """Super-batched front half for many independent batches (clusters): ONE minimizer launch and ONE
    pair-database/span-search launch sequence for all of them, then WIS per batch on the host.
    batches: list of {r_id: (acc, seq, qual)}.  Returns one (all_intervals_for_graph, new_all_reads, skipped,
    SpanIndex) tuple per batch -- identical to calling batch_front_half on each."""
    xmin = max(xmin, 2 * k)
    merged, owner = {}, []
    for b, reads in enumerate(batches):
        for r in reads:
            merged[(b, r)] = reads[r]
            owner.append((b, r))
    M_all = hotpath.get_minimizers_and_positions(merged, w, k, "lex", ctx) if merged else {}
    Ms = [{} for _ in batches]
    for (b, r) in owner:
        Ms[b][r] = M_all[(b, r)]
    indexes = spans.span_indexes_multi(batches, Ms, k, xmin, xmax, delta_len, ctx)
    return [spans.batch_intervals_for_graph(idx, reads) + (idx,) for idx, reads in zip(indexes, batches)]
