"""Latency of the fused tiny-call path (development aid): wall time per call; kernel time comes from ncu over this script."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from isonform_b200 import _lib, synth
ctx = _lib.Context(0)
rng = np.random.default_rng(1)
for L in (30, 100, 250, 500):
    x = synth.random_dna(rng, L); y = synth.mutate(rng, x, 0.05)
    for _ in range(20):
        ctx.sg_align_one(x, y, 2, -8, 12, 1)
    t0 = time.perf_counter()
    for _ in range(300):
        ctx.sg_align_one(x, y, 2, -8, 12, 1)
    print("L=%d  %.1f us per sg_align_one" % (L, (time.perf_counter() - t0) / 300 * 1e6), flush=True)
