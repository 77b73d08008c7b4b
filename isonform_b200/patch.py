"""Rebinds the reference's hot-path seams to the B200 implementations.

    import isonform_b200.patch as b200
    b200.install(globals())        # at the top of the reference's `main`, after its imports

`install` needs the reference's `modules` package importable (it patches module attributes in
place) and a CUDA device (no CPU fallback).  `ctx` may be any object with the `_lib.Context`
interface.
"""
import functools

from . import bubbles, hotpath, merge, spans


def install(main_globals=None, ctx=None, memo=None, in_process_poa=False):
    """Returns the AlignmentMemo shared by all patched alignment seams (`memo.bubbles` is the
    bubble-popping prefetcher with its per-iteration statistics).
    in_process_poa=True also replaces the `spoa` subprocess (IsoformGeneration.run_spoa, :143-156) by the library's host
    POA: an independent consensus implementation, so outputs then match a reference run only if that run uses the same
    code as its `spoa` (scripts/isof_spoa)."""
    from modules import IsoformGeneration, batch_merging_parallel, consensus   # the reference tree

    memo = memo or merge.AlignmentMemo(ctx)
    if in_process_poa:
        from . import poa
        IsoformGeneration.run_spoa = poa.run_spoa          # before the prefetcher wraps it with its memo
    # modules/consensus.py:55  (callers SimplifyGraph.py:644, IsoformGeneration.py:381)
    consensus.parasail_alignment = memo.parasail_alignment
    # modules/SimplifyGraph.py:883-1151: one batched alignment call per bubble-popping iteration instead of one
    # GPU call per candidate pair (hooks SimplifyGraph.generate_combinations, :938, and memoises run_spoa)
    memo.bubbles = bubbles.BubblePrefetcher(memo).install()
    # modules/IsoformGeneration.py:380
    IsoformGeneration.align_to_merge = memo.align_to_merge
    # modules/IsoformGeneration.py:464 -- all-pairs loop with batched speculation
    IsoformGeneration.merge_consensuses = functools.partial(
        merge.merge_consensuses,
        generate_all_consensuses=IsoformGeneration.generate_all_consensuses,
        generate_new_full_consensus=IsoformGeneration.generate_new_full_consensus,
        add_merged_reads=IsoformGeneration.add_merged_reads, ctx=ctx)
    # modules/batch_merging_parallel.py:160
    batch_merging_parallel.actual_merging_process = functools.partial(
        merge.actual_merging_process, generate_consensus_path=batch_merging_parallel.generate_consensus_path,
        memo=memo, ctx=ctx)
    if main_globals is not None:
        # main:159, main:174, main:308 (module-level functions of the `main` script)
        main_globals["get_minimizers_and_positions"] = functools.partial(hotpath.get_minimizers_and_positions, ctx=ctx)
        main_globals["get_kmer_minimizers"] = functools.partial(hotpath.get_kmer_minimizers, ctx=ctx)
        main_globals["get_minimizer_combinations_database"] = functools.partial(
            spans.get_minimizer_combinations_database, ctx=ctx)
        main_globals["find_most_supported_span"] = spans.find_most_supported_span
    return memo
