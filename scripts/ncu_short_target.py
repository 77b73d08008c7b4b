"""Program profiled by `ncu --set full -k regex:sg_dp_short_kernel`: one all-short call (400 k pairs of ~100 bp)."""
import sys
sys.path.insert(0, ".")
import numpy as np
from isonform_b200 import _lib, hotpath, synth
ctx = _lib.Context(0)
rng = np.random.default_rng(5)
L, nb, P = 100, 2000, 400000
base = [synth.random_dna(rng, L) for _ in range(nb)]
muts = [synth.mutate(rng, x, 0.05) or "A" for x in base]
cat, off = hotpath.pack_sequences(base + muts)
pa = (np.arange(P) % nb).astype(np.int32); pb = (nb + (np.arange(P) % nb)).astype(np.int32)
ctx.set_short(2)
ctx.sg_align_pairs(cat, off, pa, pb, 2, -8, 12, 1, want_cigar=False, features_delta_len=5)
lens = np.diff(off)
print("cells", int((lens[pa].astype(np.int64) * lens[pb]).sum()))
