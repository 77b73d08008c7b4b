"""CPU oracle for the isONform hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  Nothing under ``isonform_b200/`` does.

The C part (``isof_oracle.c``) restates SipHash/minimizers (pinned by executing the reference) and
parasail's semi-global affine alignment + cigar (PARITY UNPINNED: parasail is a third-party
dependency whose source is not in the reference tree and which is not installable here).
The Python part restates the integer pair/span path (main:174-224, 308-338) and the CIGAR
decision functions (SimplifyGraph.py:175-196, IsoformGeneration.py:251-397).
"""
from .capi import (  # noqa: F401
    pyhash, minimizers, minimizers_batch, sg_align, sg_align_pairs, num_threads, simd_lanes, simd_isa, TieBreak, all_tiebreaks,
    cigar_runs_to_string, cigar_runs_to_tuples, OPS, edit_distance, edit_distance_pairs, edit_path,
)
from .pyside import (  # noqa: F401
    get_kmer_minimizers, get_minimizers_and_positions, minimizers_comb_iterator,
    get_minimizer_combinations_database, find_most_supported_span, solve_WIS,
    remove_read_polyA_ends, cigar_to_seq, parasail_alignment, parse_cigar_diversity,
    find_first_significant_match, parse_cigar_diversity_isoform_level_new, align_to_merge,
    align_bubble_decision, isoform_features,
)
