"""Development aid: random long pairs (re-based 16-bit lanes, banded trace, warp traceback) against the 32-bit path of the
same library -- GPU vs GPU, so thousands of 4-9 kb pairs are cheap; the 32-bit path itself is pinned against the oracle
by the parity tests.  Exits non-zero on the first difference."""
import sys
sys.path.insert(0, ".")
import numpy as np
from isonform_b200 import _lib, hotpath, synth

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ctx = _lib.Context(0)
bad = 0
for rd in range(rounds):
    rng = np.random.default_rng(90000 + rd)
    seqs, pa, pb = [], [], []
    gene = synth.random_dna(rng, int(rng.integers(5000, 9500)))
    cuts = sorted(rng.integers(0, len(gene), size=12).tolist())
    for _ in range(int(rng.integers(8, 40))):
        def variant():
            v = gene
            for _k in range(int(rng.integers(0, 4))):                 # skip / duplicate random segments
                a, b = sorted(rng.choice(cuts, 2, replace=False).tolist())
                if b - a > 1500 or a == b:
                    continue
                v = v[:a] + v[b:] if rng.random() < 0.7 else v[:b] + v[a:b] + v[b:]
            lo = int(rng.integers(0, 400)) if rng.random() < 0.3 else 0
            hi = len(v) - (int(rng.integers(0, 400)) if rng.random() < 0.3 else 0)
            return synth.mutate(rng, v[lo:hi], float(rng.choice([0.0, 0.01, 0.05, 0.12])))
        x, y = variant(), variant()
        if min(len(x), len(y)) < 50:
            continue
        seqs += [x, y]; pa.append(len(seqs) - 2); pb.append(len(seqs) - 1)
    params = [(2, -2, 12, 1), (2, -8, 12, 1), (3, -4, 6, 2), (1, -1, 1, 1)][int(rng.integers(0, 4))]
    tb = [int(v) for v in rng.integers(0, 2, size=4)]
    band, tbm = int(rng.choice([0, 16, 128, 1024])), int(rng.choice([0, 1, 2]))
    budget = int(rng.choice([0, 0, 600 << 20]))
    ctx.set_tiebreak(*tb); ctx.set_band(band); ctx.set_traceback_warp(tbm); ctx.set_trace_budget(budget)
    ctx.set_packed(True)
    a = hotpath.align_pairs(seqs, pa, pb, *params, ctx=ctx, features_delta_len=5)
    af = hotpath.align_pairs(seqs, pa, pb, *params, ctx=ctx, want_cigar=False, features_delta_len=5)
    ctx.set_packed(False); ctx.set_band(0); ctx.set_traceback_warp(0)
    b = hotpath.align_pairs(seqs, pa, pb, *params, ctx=ctx, features_delta_len=5)
    ok = all(np.array_equal(a[k], b[k]) for k in ("score", "end_q", "end_r", "cigar_off", "cigar", "features")) and \
        all(np.array_equal(af[k], b[k]) for k in ("score", "cigar_off", "features"))
    print("round %d: %d pairs, params %s, table %s, band %d, traceback %d, budget %d: %s" % (rd, len(pa), params, tb, band, tbm, budget, "ok" if ok else "DIFFERENT"), flush=True)
    bad += not ok
ctx.set_tiebreak(0, 0, 0, 0)
sys.exit(1 if bad else 0)
