"""Quick GPU timing of the DP path (development aid, not the bench)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from isonform_b200 import _lib, hotpath, synth

ctx = _lib.Context(0)
prof_on = not (len(sys.argv) > 3 and sys.argv[3] == "noprof")
ctx.profile_enable(prof_on)
n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 200
if len(sys.argv) > 2: ctx.set_packed(sys.argv[2] != "0")
reads, _, _ = synth.make_config("C2", n_reads=n_reads)
seqs = [s for _, s in reads]
cat, off = hotpath.pack_sequences(seqs)
ia, ib = np.triu_indices(len(seqs), 1)
for it in range(3):
    ctx.profile_reset()
    t0 = time.time()
    res = ctx.sg_align_pairs(cat, off, ia, ib, 2, -2, 12, 1, want_cigar=True, features_delta_len=5)
    dt = time.time() - t0
    pr = ctx.profile()
    if not prof_on:
        lens = np.diff(off); cells = float((lens[ia].astype(np.int64) * lens[ib]).sum())
        print("pairs", len(ia), "wall %.3fs" % dt, "GCUPS(wall) %.1f" % (cells / dt / 1e9)); continue
    print("pairs", len(ia), "wall %.3fs" % dt, "dp_ms %.1f tb_ms %.1f other %.2f" % (pr["dp_ms"], pr["traceback_ms"], pr["other_ms"]),
          "GCUPS(dp) %.1f" % (pr["dp_cells"] / pr["dp_ms"] / 1e6), "GCUPS(wall) %.1f" % (pr["dp_cells"] / dt / 1e9),
          "launches", pr["launches"], "dp_launches", pr["dp_launches"], "runs", len(res["cigar"]))
t0 = time.time()
pos = hotpath.minimizer_positions_batch(seqs, 20, 31, ctx)
print("minimizers wall %.4fs" % (time.time() - t0), sum(len(p) for p in pos), ctx.profile()["minimizer_ms"])
