"""Randomised check of kernel (b') under a bound k (the banded ring classes and their neighbours) against the oracle:
    python scripts/fuzz_myers_band.py [rounds]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from isonform_b200 import _lib, hotpath, synth  # noqa: E402

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 200
ctx = _lib.Context(0)
tot = 0
for rd in range(rounds):
    rng = np.random.default_rng(5000 + rd)
    alpha = "ACGTNRYK"[: int(rng.integers(2, 9))]
    seqs = []
    for _ in range(int(rng.integers(3, 14))):
        n = int(np.exp(rng.uniform(np.log(150), np.log(12000))))
        s = "".join(alpha[int(x)] for x in rng.integers(0, len(alpha), n))
        e = float(rng.random()) * 0.05
        cut = int(rng.integers(0, 40))
        seqs += [s, synth.mutate(rng, s, e) or "A", s[cut:], s[: n - cut] + "ACGT"[: int(rng.integers(0, 4))]]
    cat, off = hotpath.pack_sequences(seqs)
    S = len(seqs)
    base = (np.arange(S) // 4 * 4).astype(np.int32)
    pa = np.concatenate([base, np.arange(S, dtype=np.int32), rng.integers(0, S, 10).astype(np.int32)])
    pb = np.concatenate([np.arange(S, dtype=np.int32), base, rng.integers(0, S, 10).astype(np.int32)])
    full = oracle.edit_distance_pairs(cat, off, pa, pb, threads=os.cpu_count())
    for k in (int(rng.integers(0, 40)), int(rng.integers(0, 520)), int(rng.integers(0, 1200))):
        got = ctx.edit_distance_pairs(cat, off, pa, pb, k=k)["dist"]
        want = np.where(full <= k, full, -1)
        if not (got == want).all():
            p = int(np.nonzero(got != want)[0][0])
            raise SystemExit("MISMATCH round %d k %d pair %d: n %d m %d got %d want %d" % (rd, k, p, len(seqs[pa[p]]), len(seqs[pb[p]]), got[p], want[p]))
        tot += len(pa)
print("fuzz ok: %d rounds, %d bounded distances" % (rounds, tot))
