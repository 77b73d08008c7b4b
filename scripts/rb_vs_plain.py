"""Kernel rate of the re-based 16-bit lanes vs the plain 16-bit lanes at nearly the same pair size (development aid)."""
import sys, json
sys.path.insert(0, ".")
import numpy as np
from isonform_b200 import _lib, hotpath, synth
ctx = _lib.Context(0)
rng = np.random.default_rng(3)
for L in (3900, 4300):
    base = synth.random_dna(rng, L)
    seqs = sorted((synth.mutate(rng, base, 0.05) for _ in range(260)), key=len)
    if L == 3900:
        seqs = [s[:4090] for s in seqs]
    ib, ia = np.tril_indices(len(seqs), -1)
    cat, off = hotpath.pack_sequences(seqs)
    lens = np.diff(off); cells = float((lens[ia].astype(np.int64) * lens[ib]).sum())
    for rep in range(2):
        ctx.set_overlap(False)
        ctx.profile_enable(True); ctx.profile_reset()
        ctx.sg_align_pairs(cat, off, ia.astype(np.int32), ib.astype(np.int32), 2, -2, 12, 1, want_cigar=False, features_delta_len=5)
        pr = ctx.profile(); ctx.profile_enable(False)
    print(json.dumps({"L": L, "max_len": int(lens.max()), "pairs": len(ia), "dp_ms": pr["dp_ms"], "dp_launches": pr["dp_launches"],
                      "gcups_kernel": cells / pr["dp_ms"] / 1e6, "packed_frac": pr["dp_cells_packed"] / pr["dp_cells"]}), flush=True)
