"""Development aid: host calls of 1-8 pairs of up to 4096 x 4096 through the fused one-launch kernel (shared-memory or
device trace, chase mode, warp-cooperative traceback) against the regular path of the same library.  Exits non-zero on the
first difference."""
import sys
sys.path.insert(0, ".")
import numpy as np
from isonform_b200 import _lib, hotpath, synth

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 200
ctx = _lib.Context(0)
bad = 0
for rd in range(rounds):
    rng = np.random.default_rng(555000 + rd)
    seqs, pa, pb = [], [], []
    for _ in range(int(rng.choice([1, 1, 1, 2, 5, 8]))):
        n = int(rng.choice([rng.integers(1, 300), rng.integers(250, 1100), rng.integers(1000, 4097)]))
        x = synth.random_dna(rng, n)
        kind = int(rng.integers(0, 5))
        if kind == 0:
            y = synth.random_dna(rng, int(rng.integers(1, 4097)))
        elif kind == 1:
            y = synth.mutate(rng, x, float(rng.choice([0.0, 0.03, 0.1]))) or "A"
        elif kind == 2:
            c = int(rng.integers(0, n)); y = (x[:c] + x[c + int(rng.integers(0, 600)):]) or "C"
        elif kind == 3:
            y = x[int(rng.integers(0, n)):] + synth.random_dna(rng, int(rng.integers(0, 200))) or "G"
        else:
            y = synth.mutate(rng, x, 0.05)
            if rng.random() < 0.5:
                y = y[:len(y) // 2] + "N" * 3 + y[len(y) // 2:]
        y = y[:4096]
        if rng.random() < 0.5:
            x, y = y, x
        seqs += [x, y]; pa.append(len(seqs) - 2); pb.append(len(seqs) - 1)
    params = [(2, -2, 12, 1), (2, -8, 12, 1), (1, -1, 1, 1)][int(rng.integers(0, 3))]
    tb = [int(v) for v in rng.integers(0, 2, size=4)] if rng.random() < 0.3 else [0, 0, 0, 0]
    ctx.set_tiebreak(*tb)
    ctx.set_fused_small(True)
    a = hotpath.align_pairs(seqs, pa, pb, *params, ctx=ctx, features_delta_len=5)
    a2 = hotpath.align_pairs(seqs, pa, pb, *params, ctx=ctx)
    ctx.set_fused_small(False)
    b = hotpath.align_pairs(seqs, pa, pb, *params, ctx=ctx, features_delta_len=5)
    b2 = hotpath.align_pairs(seqs, pa, pb, *params, ctx=ctx)
    ok = all(np.array_equal(a[k], b[k]) for k in ("score", "cigar_off", "cigar", "features")) and \
        all(np.array_equal(a2[k], b2[k]) for k in ("score", "end_q", "end_r", "cigar_off", "cigar"))
    if not ok or rd % 50 == 0:
        print("round %d: shapes %s params %s table %s: %s" % (rd, [(len(seqs[i]), len(seqs[j])) for i, j in zip(pa, pb)], params, tb,
                                                              "ok" if ok else "DIFFERENT"), flush=True)
    bad += not ok
ctx.set_tiebreak(0, 0, 0, 0); ctx.set_fused_small(True)
print("rounds %d, different %d" % (rounds, bad))
sys.exit(1 if bad else 0)
