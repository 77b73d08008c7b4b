#!/usr/bin/env python
"""How much does the UNPINNED part of the oracle matter?  (CPU only.)

parasail is absent, so three tie-breaks of the restated aligner rest on recollection (DESIGN.md section 3:
H source E-vs-F order, gap open-vs-extend on ties, end-cell scan order).  This script re-runs the
reference's two decisions -- bubble popping (parse_cigar_diversity, SimplifyGraph.py:175-196) and isoform
merging (align_to_merge, IsoformGeneration.py:380-397) -- under all 8 settings of the tie-break table on
seeded synthetic pairs in both regimes and counts (a) CIGARs that differ from the default table and
(b) DECISIONS that differ.  Output: one JSON document (profiles/rNN_tiebreak_sensitivity.json)."""
import itertools
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from isonform_b200 import synth  # noqa: E402


def pairs_bubble(rng, n):
    out = []
    for i in range(n):
        ln = int(rng.integers(20, 400))
        x = synth.random_dna(rng, ln)
        y = synth.mutate(rng, x, [0.01, 0.03, 0.05, 0.1][i % 4])
        if i % 5 == 0 and ln > 40:
            c = int(rng.integers(5, ln - 20)); y = y[:c] + y[c + int(rng.integers(1, 12)):]
        if i % 7 == 0:
            y = y + synth.random_dna(rng, int(rng.integers(1, 8)))
        out.append((x, y or "A"))
    return out


def pairs_isoform(rng, n):
    out = []
    for i in range(n):
        ln = int(rng.integers(300, 1600))
        x = synth.random_dna(rng, ln)
        y = synth.mutate(rng, x, [0.002, 0.01, 0.03, 0.05][i % 4])
        if i % 3 == 0:
            y = y[int(rng.integers(0, 70)):len(y) - int(rng.integers(0, 50))]
        if i % 6 == 1:
            c = len(y) // 2; y = y[:c] + y[c + int(rng.integers(3, 40)):]
        out.append((y, x) if len(y) <= len(x) else (x, y))
    return out


def main():
    rng = np.random.default_rng(20261019)
    nb, ni = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3000, 400)
    tables = list(itertools.product((0, 1), repeat=3))
    res = {"tables": ["h_e_before_f=%d gap_open_on_tie=%d end_col_first=%d" % t for t in tables], "regimes": {}}
    for name, pairs, decide in (
            ("bubble (2,-8,12,1) delta_len 5", pairs_bubble(rng, nb),
             lambda a, b, tb: oracle.align_bubble_decision(a, b, 5, tb)[0]),
            ("isoform (2,-2,12,1) delta 0.15 delta_len 5", pairs_isoform(rng, ni),
             lambda a, b, tb: oracle.align_to_merge(a, b, 0.15, 5, 30, 50, tb))):
        base_dec = [decide(a, b, None) for a, b in pairs]
        base_cig = [oracle.sg_align(a, b, 2, -8 if name.startswith("bubble") else -2, 12, 1)[3].tolist() for a, b in pairs]
        rows = []
        for t in tables:
            tb = oracle.TieBreak(*t)
            dec = [decide(a, b, tb) for a, b in pairs]
            cig = [oracle.sg_align(a, b, 2, -8 if name.startswith("bubble") else -2, 12, 1, tiebreak=tb)[3].tolist()
                   for a, b in pairs]
            rows.append({"table": list(t), "cigars_differing": int(sum(x != y for x, y in zip(cig, base_cig))),
                         "decisions_differing": int(sum(x != y for x, y in zip(dec, base_dec)))})
        res["regimes"][name] = {"pairs": len(pairs), "positive_decisions_default": int(sum(base_dec)), "by_table": rows}
        print(name, json.dumps(rows), flush=True)
    json.dump(res, sys.stdout if len(sys.argv) <= 3 else open(sys.argv[3], "w"), indent=1)


if __name__ == "__main__":
    main()
