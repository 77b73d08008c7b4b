"""Drivers of the all-pairs alignment loops.

`consensus.parasail_alignment` is a pure function of (s1, s2, scores); everything around it in the
reference is sequential (first success breaks a row, merged targets may get a regenerated
consensus).  The drivers here keep that sequential semantics bit-for-bit and move the alignments
off the critical path by *speculation*: decisions for every pair the loop can ask about are
computed in a few large GPU batches, then the reference's loop is replayed on the host over the
decision table; only pairs whose strings changed under the replay are re-aligned.

    merge_consensuses          modules/IsoformGeneration.py:464-507 (all-pairs within a batch)
    actual_merging_process     modules/batch_merging_parallel.py:160-212 (across batches)
    AlignmentMemo              the memo + prefetch behind the per-pair seams
                               (consensus.parasail_alignment, IsoformGeneration.align_to_merge)
"""
import numpy as np

from . import hotpath
from ._lib import default_context

# pairs per speculative launch: large enough to fill the GPU several times over (148 SMs x 16 warps x 2 pairs = one
# wave of 4736 pairs; a 64 k-pair launch takes ~0.15 s at C2 read lengths), which is also the most an early `break`
# can waste -- cheaper than the extra launches and host round trips of smaller blocks
_MIN_PAIRS_PER_CALL = 65536
_MAX_PAIRS_PER_CALL = 1 << 19


class AlignmentMemo:
    """Memo of alignment results keyed by the exact argument tuple; `prefetch_*` fill it in batches.

    Argument order is part of the key: the alignment is not symmetric under the tie-break rules
    (batch_merging_parallel.py:191-195 passes (long, short), IsoformGeneration.py:486-491
    (short, long))."""

    def __init__(self, ctx=None):
        self.ctx = ctx
        self.cigars = {}          # (s1, s2, match, mismatch, open, ext) -> (cigar_string, score)
        self.merge_feats = {}     # (s1, s2, delta_len) -> int32[12] (own copy, not a view of a batch result)
        self.hits = self.misses = 0

    def clear(self):
        """Drop the memoised results (e.g. between batches: keys are whole consensus strings)."""
        self.cigars.clear()
        self.merge_feats.clear()

    def _ctx(self):
        if self.ctx is None:
            self.ctx = default_context()
        return self.ctx

    # -- consensus.parasail_alignment ---------------------------------------------------------
    def prefetch_alignments(self, pairs, match_score=2, mismatch_penalty=-2, opening_penalty=12, gap_ext=1):
        params = (match_score, mismatch_penalty, opening_penalty, gap_ext)
        todo = [(a, b) for a, b in dict.fromkeys(pairs) if (a, b) + params not in self.cigars and a and b]
        if not todo:
            return 0
        one = getattr(self._ctx(), "sg_align_one", None)
        if len(todo) == 1 and one is not None:             # the memo-miss seam: lean single-pair call
            a, b = todo[0]
            score, _eq, _er, runs = one(a, b, *params)
            self.cigars[(a, b) + params] = (hotpath.runs_to_string(runs), score)
            return 1
        pool, pa, pb = _pool(todo)
        res = hotpath.align_pairs(pool, pa, pb, *params, ctx=self._ctx())
        off = res["cigar_off"]
        for p, (a, b) in enumerate(todo):
            self.cigars[(a, b) + params] = (hotpath.runs_to_string(res["cigar"][off[p]:off[p + 1]]), int(res["score"][p]))
        return len(todo)

    def parasail_alignment(self, s1, s2, match_score=2, mismatch_penalty=-2, opening_penalty=12, gap_ext=1):
        """Drop-in for consensus.parasail_alignment (modules/consensus.py:55-67)."""
        key = (s1, s2, match_score, mismatch_penalty, opening_penalty, gap_ext)
        hit = self.cigars.get(key)
        if hit is None:
            self.misses += 1
            if len(s1) == 0 or len(s2) == 0:
                raise ValueError("parasail rejects empty sequences")
            self.prefetch_alignments([(s1, s2)], match_score, mismatch_penalty, opening_penalty, gap_ext)
            hit = self.cigars[key]
        else:
            self.hits += 1
        cigar_string, score = hit
        a1, a2, tuples = hotpath.cigar_to_seq(cigar_string, s1, s2)
        return a1, a2, cigar_string, tuples, score

    # -- IsoformGeneration.align_to_merge -----------------------------------------------------
    def prefetch_merge(self, pairs, delta_len):
        todo = [(a, b) for a, b in dict.fromkeys(pairs) if (a, b, delta_len) not in self.merge_feats and a and b]
        if not todo:
            return 0
        pool, pa, pb = _pool(todo)
        res = hotpath.align_pairs(pool, pa, pb, 2, -2, 12, 1, ctx=self._ctx(), want_cigar=False, features_delta_len=delta_len)
        for p, (a, b) in enumerate(todo):
            self.merge_feats[(a, b, delta_len)] = res["features"][p].copy()
        return len(todo)

    def align_to_merge(self, consensus1, consensus2, delta, delta_len, delta_iso_len_3, delta_iso_len_5):
        """Drop-in for IsoformGeneration.align_to_merge (IsoformGeneration.py:380-397)."""
        key = (consensus1, consensus2, delta_len)
        f = self.merge_feats.get(key)
        if f is None:
            self.misses += 1
            if len(consensus1) == 0 or len(consensus2) == 0:
                raise ValueError("parasail rejects empty sequences")
            self.prefetch_merge([(consensus1, consensus2)], delta_len)
            f = self.merge_feats[key]
        else:
            self.hits += 1
        return bool(hotpath._merge_from_features(f[None, :], delta, delta_len, delta_iso_len_3, delta_iso_len_5)[0])


def _pool(pairs):
    """Deduplicated string pool + index pairs."""
    index = {}
    pool = []
    pa = np.empty(len(pairs), dtype=np.int32)
    pb = np.empty(len(pairs), dtype=np.int32)
    for p, (a, b) in enumerate(pairs):
        ia = index.get(a)
        if ia is None:
            ia = index[a] = len(pool)
            pool.append(a)
        ib = index.get(b)
        if ib is None:
            ib = index[b] = len(pool)
            pool.append(b)
        pa[p], pb[p] = ia, ib
    return pool, pa, pb


# ----------------------------------------------------------------------------------------------
# C1: merge_consensuses  (IsoformGeneration.py:464-507)
# ----------------------------------------------------------------------------------------------
def _row_decisions(pool, rows, n, seq1_override, delta, delta_len, d3, d5, ctx):
    """Decisions for rows `rows` (ascending list of i) against every j > i.  The left string of row i is
    seq1_override.get(i) or pool[i]; the right string is always the original pool[j]
    (IsoformGeneration.py:476-491: shorter string first, equal lengths keep (seq1, seq2))."""
    out = {i: np.zeros(n - 1 - i, dtype=bool) for i in rows}
    if not rows or min(rows) + 1 >= n:
        return out
    seqs = list(pool)
    left = np.arange(n, dtype=np.int64)                    # index of row i's left string in `seqs`
    for i in rows:
        if i in seq1_override:
            left[i] = len(seqs)
            seqs.append(seq1_override[i])
    lens = np.fromiter((len(x) for x in seqs), dtype=np.int64, count=len(seqs))
    # pairs are issued grouped by the right-hand string j (column-major over the block: j outer, i inner), so that
    # consecutive alignments share their reference sequence -- the DP kernel then reads substitution scores from a
    # per-lane query profile instead of recomputing them (csrc/sg_align.cu, PROF path)
    r = np.asarray(rows, dtype=np.int64)
    js = np.arange(int(r[0]) + 1, n, dtype=np.int64)
    I, J = np.meshgrid(r, js, indexing="xy")               # shape (len(js), len(rows))
    keep = I < J
    i_flat, j_flat = I[keep], J[keep]
    li = left[i_flat]
    swap = lens[j_flat] < lens[li]
    pa = np.where(swap, j_flat, li)
    pb = np.where(swap, li, j_flat)
    dec = np.asarray(hotpath.align_to_merge_pairs(seqs, pa, pb, delta, delta_len, d3, d5, ctx), dtype=bool)
    order = np.argsort(i_flat, kind="stable")              # back to row-major: row i, then j ascending
    dec = dec[order]
    pos = 0
    for i in rows:                                          # rows ascending == order of i_flat[order]
        cnt = n - 1 - i
        out[i] = dec[pos:pos + cnt].copy()
        pos += cnt
    return out


def merge_consensuses(curr_best_seqs, work_dir, delta, delta_len, reads, max_seqs_to_spoa, delta_iso_len_3,
                      delta_iso_len_5, *, generate_all_consensuses, generate_new_full_consensus, add_merged_reads=None,
                      ctx=None, stats=None):
    """Same arguments and result (`new_consensuses`, `curr_best_seqs` mutated in place) as
    IsoformGeneration.merge_consensuses.  The three consensus helpers are the caller's (the
    reference's own when patched in: they run spoa on the host)."""
    if add_merged_reads is None:
        def add_merged_reads(cbs, id2, id1):
            cbs[id2] = cbs[id2] + cbs[id1]
            cbs.pop(id1)
    new_consensuses, all_consensuses, alternative = {}, {}, []
    generate_all_consensuses(all_consensuses, alternative, curr_best_seqs, reads, work_dir, max_seqs_to_spoa)
    alternative.sort(key=lambda x: len(x[1]))
    n = len(alternative)
    ids = [a[0] for a in alternative]
    pool = [a[1] for a in alternative]
    row_of = {id_: i for i, id_ in enumerate(ids)}
    n_aligned = n_used = 0
    i = 0
    block_pairs = _MIN_PAIRS_PER_CALL          # grows while every speculated pair is used (rows that never merge)
    while i < n - 1:
        # ---- speculate a block of rows
        rows, budget = [], 0
        r = i
        while r < n - 1 and (budget < block_pairs or not rows):
            rows.append(r)
            budget += n - 1 - r
            r += 1
        merged_in_block = False
        override = {row: new_consensuses[ids[row]] for row in rows if ids[row] in new_consensuses}
        table = _row_decisions(pool, rows, n, override, delta, delta_len, delta_iso_len_3, delta_iso_len_5, ctx)
        n_aligned += budget
        # ---- replay the reference loop over the block
        for row in rows:
            id1 = ids[row]
            if id1 in new_consensuses and override.get(row, pool[row]) != new_consensuses[id1]:
                # this row's left string changed while replaying an earlier row of the block
                override2 = {row: new_consensuses[id1]}
                table[row] = _row_decisions(pool, [row], n, override2, delta, delta_len, delta_iso_len_3, delta_iso_len_5,
                                            ctx)[row]
                n_aligned += n - 1 - row
            dec = table[row]
            hit = np.nonzero(dec)[0]
            n_used += (int(hit[0]) + 1) if len(hit) else len(dec)
            if len(hit):
                merged_in_block = True
                id2 = ids[row + 1 + int(hit[0])]
                if len(curr_best_seqs[id2]) > 50:
                    add_merged_reads(curr_best_seqs, id2, id1)
                    new_consensuses[id2] = pool[row_of[id2]]
                else:
                    new_consensuses[id2] = generate_new_full_consensus(id1, id2, reads, curr_best_seqs, work_dir, max_seqs_to_spoa)
                    add_merged_reads(curr_best_seqs, id2, id1)
        i = rows[-1] + 1
        # all speculated pairs consumed (no row broke early): the batch behaves like an all-pairs job, so issue it in
        # launches that fill the GPU several times over; an early break shrinks the block again
        if not merged_in_block:
            block_pairs = min(block_pairs * 4, _MAX_PAIRS_PER_CALL)
        else:
            block_pairs = _MIN_PAIRS_PER_CALL
    for id_ in curr_best_seqs:
        if id_ not in new_consensuses:
            new_consensuses[id_] = all_consensuses[id_]
    if stats is not None:
        stats["pairs_aligned"] = n_aligned
        stats["pairs_used"] = n_used
    return new_consensuses


# ----------------------------------------------------------------------------------------------
# C2: actual_merging_process  (batch_merging_parallel.py:160-212)
# ----------------------------------------------------------------------------------------------
def actual_merging_process(all_infos_dict, delta, delta_len, delta_iso_len_3, delta_iso_len_5, max_seqs_to_spoa,
                           all_batch_sequences, work_dir, *, generate_consensus_path, memo=None, ctx=None):
    """Same arguments/side effects as batch_merging_parallel.actual_merging_process; `infos` objects
    need .sequence, .reads, .merged.  Every (long, short) pair of the initial sequences is
    prefetched per batch pair; the loop is then replayed through the memo."""
    memo = memo or AlignmentMemo(ctx)
    all_infos_list = sorted(all_infos_dict.items(), key=lambda x: x[0], reverse=True)
    for b_i, (batchid, id_dict) in enumerate(all_infos_list[:len(all_infos_list) - 1]):
        batch_id_list = sorted(id_dict.items(), key=lambda x: len(x[1].sequence))
        for b_j, (batchid2, id_dict2) in enumerate(all_infos_list[b_i + 1:]):
            batch_id_list2 = sorted(id_dict2.items(), key=lambda x: len(x[1].sequence))
            if batchid2 <= batchid:
                continue
            spec = []
            for _id, infos in batch_id_list:
                if infos.merged:
                    continue
                for _id2, infos2 in batch_id_list2:
                    if infos2.merged:
                        continue
                    if len(infos2.sequence) >= len(infos.sequence):
                        spec.append((infos2.sequence, infos.sequence))
                    else:
                        spec.append((infos.sequence, infos2.sequence))
            memo.prefetch_merge(spec, delta_len)
            for _t1, (id_, infos) in enumerate(batch_id_list):
                if infos.merged:
                    continue
                for _t2, (id2, infos2) in enumerate(batch_id_list2):
                    if len(infos2.sequence) >= len(infos.sequence):
                        batch_id_long, batch_id_short, id_long, infos_long, id_short, infos_short = \
                            batchid2, batchid, id2, infos2, id_, infos
                    else:
                        batch_id_long, batch_id_short, id_long, infos_long, id_short, infos_short = \
                            batchid, batchid2, id_, infos, id2, infos2
                    if infos_long.merged:
                        continue
                    if not infos_short.merged:
                        if memo.align_to_merge(infos_long.sequence, infos_short.sequence, delta, delta_len, delta_iso_len_3,
                                               delta_iso_len_5):
                            if len(infos_long.reads) > 50:
                                all_infos_dict[batch_id_long][id_long].reads = infos_long.reads + infos_short.reads
                                infos_short.merged = True
                            else:
                                new_consensus = generate_consensus_path(work_dir, infos_long.reads, infos_short.reads,
                                                                        all_batch_sequences, max_seqs_to_spoa)
                                all_infos_dict[batch_id_long][id_long].sequence = new_consensus
                                all_infos_dict[batch_id_long][id_long].reads = infos_long.reads + infos_short.reads
                                all_infos_dict[batch_id_short][id_short].merged = True
    return memo
