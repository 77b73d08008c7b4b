"""Latency of the fused tiny-call path (development aid): wall time per `sg_align_one` call, per bare C-ABI call (arguments
prepared once), and the kernel's own phase clocks (isof_debug_fused_phases, SM cycles -> us at the SM clock)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from isonform_b200 import _lib, synth
ctx = _lib.Context(0)
rng = np.random.default_rng(1)
MHZ = 1965.0
for L in (30, 100, 250, 500, 1000, 2000, 4000):
    x = synth.random_dna(rng, L); y = synth.mutate(rng, x, 0.05)
    for _ in range(20):
        ctx.sg_align_one(x, y, 2, -8, 12, 1)
    t0 = time.perf_counter()
    for _ in range(300):
        ctx.sg_align_one(x, y, 2, -8, 12, 1)
    one = (time.perf_counter() - t0) / 300 * 1e6
    ph = ctx.debug_fused_phases()
    # bare C-ABI call
    b = (x + y).encode()
    off = np.array([0, len(x), len(x) + len(y)], np.int64); pa = np.zeros(1, np.int32); pb = np.ones(1, np.int32)
    out = np.zeros(3, np.int32); coff = np.zeros(2, np.int64); runs = np.empty(4 * L + 64, np.uint32)
    args = (ctx._h, b, off.ctypes.data, 2, pa.ctypes.data, pb.ctypes.data, 1, 2, -8, 12, 1, out.ctypes.data, out.ctypes.data + 4,
            out.ctypes.data + 8, runs.ctypes.data, len(runs), coff.ctypes.data)
    f = ctx._L.isof_sg_align_pairs
    t0 = time.perf_counter()
    for _ in range(300):
        f(*args)
    bare = (time.perf_counter() - t0) / 300 * 1e6
    print("L=%d  sg_align_one %.1f us  bare C call %.1f us  kernel phases us: load %.1f dp %.1f traceback %.1f finish %.1f (sum %.1f)"
          % (L, one, bare, ph[0] / MHZ, ph[1] / MHZ, ph[2] / MHZ, ph[3] / MHZ, sum(ph[:4]) / MHZ), flush=True)
    st = [x for x in ph[4] if x[1] > 0]
    if L >= 500 and st:
        t0s = min(x[0] for x in st)
        print("      stripes (begin us, end us, batch re-reads, us spent on them): " +
              " ".join("(%.0f %.0f %d %.0f)" % ((b - t0s) / 1e3, (e - t0s) / 1e3, sp, cyc / MHZ) for b, e, sp, cyc in st), flush=True)
