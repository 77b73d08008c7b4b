"""Latency of the per-pair seam (development aid)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from isonform_b200 import _lib, hotpath, synth
ctx = _lib.Context(0)
rng = np.random.default_rng(1)
for L in (100, 500, 2000):
    x = synth.random_dna(rng, L); y = synth.mutate(rng, x, 0.05)
    hotpath.parasail_alignment(x, y, ctx=ctx)
    t0 = time.perf_counter()
    for _ in range(200):
        hotpath.parasail_alignment(x, y, ctx=ctx)
    dt = (time.perf_counter() - t0) / 200
    cat, off = hotpath.pack_sequences([x, y])
    t0 = time.perf_counter()
    for _ in range(200):
        ctx.sg_align_pairs(cat, off, [0], [1])
    dt2 = (time.perf_counter() - t0) / 200
    print("L=%d  parasail_alignment %.1f us/call   raw C-ABI call %.1f us" % (L, dt * 1e6, dt2 * 1e6))
