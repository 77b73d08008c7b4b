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
    """Returns (all_intervals_for_graph, new_all_reads, skipped_read_ids, SpanIndex).  The minimizer positions go from
    kernel (a) to kernel (c) as arrays: the `{r_id: [(kmer, pos)]}` dict of the reference seam is not materialised."""
    from ._lib import default_context, pack_sequences
    ctx = ctx or default_context()
    xmin = max(xmin, 2 * k)                                      # main:623-625
    ids = list(reads)
    cat, off = pack_sequences([reads[r][1] for r in ids])
    min_pos, min_off = ctx.minimizers_batch(cat, off, k, w)
    arrays = ctx.spans_batch(cat, off, min_pos, min_off, k, xmin, xmax, delta_len)
    index = spans.SpanIndex(reads, None, k, xmin, xmax, delta_len, _arrays=arrays, ids=ids)
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


def singleton_groups(reads):
    """`curr_best_seqs` of a batch in which every read is its own path group -- what a 5 %-error batch degenerates to
    (bubble popping stops early, IsoformGeneration.compute_equal_reads then finds no two reads on
    the same s->t path), so `merge_consensuses` runs its all-pairs loop over the raw reads (IsoformGeneration.py:421-427)."""
    return {r: [r] for r in reads}


def _singleton_consensuses(all_consensuses, alternative_consensuses, curr_best_seqs, reads, work_dir, max_seqs_to_spoa):
    # IsoformGeneration.generate_all_consensuses (:419-441) for groups of one read: the consensus is the read
    for id_, members in curr_best_seqs.items():
        seq = reads[members[0]][1]
        all_consensuses[id_] = seq
        alternative_consensuses.append((id_, seq))


def _longest_member(id1, id2, reads, curr_best_seqs, work_dir, max_seqs_to_spoa):
    # stand-in for generate_new_full_consensus (:400-416): `spoa` is an external binary (absent on the GPU box)
    members = (curr_best_seqs[id1] + curr_best_seqs[id2])[:max_seqs_to_spoa]
    return max((reads[q][1] for q in members), key=len)


def batch_hot_path(reads, k=20, w=31, xmin=40, xmax=80, delta_len=5, delta=0.15, delta_iso_len_3=30, delta_iso_len_5=50,
                   max_seqs_to_spoa=200, ctx=None, stats=None, curr_best_seqs=None):
    """Everything the GPU does for one batch, through the Python seams, in the order `main()` runs it (main:341-580):
    minimizers -> pair database + span search -> WIS picks (`batch_front_half`), then the all-pairs merge of the batch's
    isoform consensuses (`merge.merge_consensuses`, IsoformGeneration.py:464-507).  The graph / bubble / spoa stages in
    between belong to the reference (host, out of scope); without them the path groups are `singleton_groups(reads)`
    unless the caller passes its own.  Returns (all_intervals_for_graph, new_all_reads, skipped, new_consensuses)."""
    import time
    from . import merge
    t0 = time.perf_counter()
    graph_iv, new_reads, skipped, _index = batch_front_half(reads, k, w, xmin, xmax, delta_len, ctx)
    t1 = time.perf_counter()
    groups = dict(curr_best_seqs) if curr_best_seqs is not None else singleton_groups(reads)
    mstats = {}
    new_consensuses = merge.merge_consensuses(groups, None, delta, delta_len, reads, max_seqs_to_spoa, delta_iso_len_3,
                                              delta_iso_len_5, generate_all_consensuses=_singleton_consensuses,
                                              generate_new_full_consensus=_longest_member, ctx=ctx, stats=mstats)
    t2 = time.perf_counter()
    if stats is not None:
        stats["front_half_s"] = stats.get("front_half_s", 0.0) + (t1 - t0)
        stats["merge_s"] = stats.get("merge_s", 0.0) + (t2 - t1)
        stats["pairs_aligned"] = stats.get("pairs_aligned", 0) + mstats.get("pairs_aligned", 0)
        stats["pairs_used"] = stats.get("pairs_used", 0) + mstats.get("pairs_used", 0)
        stats["reads"] = stats.get("reads", 0) + len(reads)
        stats["isoforms_out"] = stats.get("isoforms_out", 0) + len(groups)
    return graph_iv, new_reads, skipped, new_consensuses
