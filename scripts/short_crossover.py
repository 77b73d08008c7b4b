"""Device time of all-short calls on the short-pair kernel vs the wavefront kernel, by pair count (development aid)."""
import sys, json
sys.path.insert(0, ".")
import numpy as np
from isonform_b200 import _lib, hotpath, synth
ctx = _lib.Context(0)
rng = np.random.default_rng(5)
for L in (50, 100, 200, 300):
    nb = 2000
    base = [synth.random_dna(rng, L) for _ in range(nb)]
    muts = [(synth.mutate(rng, x, 0.05) or "A")[:384] for x in base]
    cat, off = hotpath.pack_sequences(base + muts)
    for P in (64, 512, 2048, 8192, 32768, 131072, 400000):
        pa = (np.arange(P) % nb).astype(np.int32); pb = (nb + (np.arange(P) % nb)).astype(np.int32)
        row = {"L": L, "P": P}
        for name, on in (("short", 2), ("wave", 0)):
            ctx.set_short(on)
            ctx.sg_align_pairs(cat, off, pa, pb, 2, -8, 12, 1, want_cigar=False, features_delta_len=5)
            ctx.profile_enable(True); ctx.profile_reset()
            for _ in range(3):
                ctx.sg_align_pairs(cat, off, pa, pb, 2, -8, 12, 1, want_cigar=False, features_delta_len=5)
            pr = ctx.profile(); ctx.profile_enable(False)
            row[name + "_ms"] = round((pr["dp_ms"] + pr["traceback_ms"] + pr["other_ms"]) / 3, 4)
            row[name + "_dp_ms"] = round(pr["dp_ms"] / 3, 4)
        ctx.set_short(1)
        print(json.dumps(row), flush=True)
