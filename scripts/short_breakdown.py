"""Per-kernel device time of short-pair batches (development aid for the short-pair regime)."""
import sys, json
sys.path.insert(0, ".")
import numpy as np
from isonform_b200 import _lib, hotpath, synth
ctx = _lib.Context(0)
rng = np.random.default_rng(5)
for L in (30, 50, 100, 200, 400):
    nb = 2000
    n_pairs = 200000
    base = [synth.random_dna(rng, L) for _ in range(nb)]
    muts = [synth.mutate(rng, x, 0.05) or "A" for x in base]
    seqs = base + muts
    pa = (np.arange(n_pairs) % nb).astype(np.int32)
    pb = (nb + (np.arange(n_pairs) % nb)).astype(np.int32)
    cat, off = hotpath.pack_sequences(seqs)
    lens = np.diff(off)
    cells = float((lens[pa].astype(np.int64) * lens[pb]).sum())
    ctx.sg_align_pairs(cat, off, pa, pb, 2, -8, 12, 1)
    ctx.profile_enable(True); ctx.profile_reset()
    ctx.sg_align_pairs(cat, off, pa, pb, 2, -8, 12, 1)
    pr = ctx.profile(); ctx.profile_enable(False)
    print(json.dumps({"L": L, "pairs": n_pairs, "dp_ms": pr["dp_ms"], "tb_ms": pr["traceback_ms"], "other_ms": pr["other_ms"],
                      "dp_gcups": cells / pr["dp_ms"] / 1e6, "all_gcups": cells / (pr["dp_ms"] + pr["traceback_ms"] + pr["other_ms"]) / 1e6}), flush=True)
