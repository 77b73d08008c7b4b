"""One isof_edit_distance_pairs call (for ncu): python scripts/myers_one.py LEN PAIRS RELATED path|distance"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isonform_b200 import _lib, hotpath, synth  # noqa: E402

L, n_pairs, related, mode = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]), sys.argv[4]
rng = np.random.default_rng(5)
nb = 2000 if L <= 1000 else 400
base = [synth.random_dna(rng, L) for _ in range(nb)]
muts = [synth.mutate(rng, x, 0.05) or "A" for x in base]
pa = (np.arange(n_pairs) % nb).astype(np.int32)
pb = (nb + (np.arange(n_pairs) * 7 % nb)).astype(np.int32)
rel = rng.random(n_pairs) < related
pb[rel] = nb + pa[rel]
cat, off = hotpath.pack_sequences(base + muts)
lens = np.diff(off)
print("cells", float((lens[pa].astype(np.int64) * lens[pb]).sum()))
ctx = _lib.Context(0)
for _ in range(int(sys.argv[5]) if len(sys.argv) > 5 else 2):
    r = ctx.edit_distance_pairs(cat, off, pa, pb, want_cigar=(mode == "path"))
print("dist sum", int(r["dist"].sum()), "runs", 0 if r["cigar"] is None else len(r["cigar"]))
