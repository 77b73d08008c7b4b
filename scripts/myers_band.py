"""Kernel (b') under a bound k: device time of isof_edit_distance_pairs(k) on noisy copies (1 % error) by length and k --
the banded ring classes against the full computation (k = -1).   python scripts/myers_band.py [--out file]"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/myers_band.jsonl")
    args = ap.parse_args()
    import oracle
    from isonform_b200 import _lib, hotpath, synth
    ctx = _lib.Context(0)
    rng = np.random.default_rng(7)
    with open(args.out, "w") as fh:
        for L, n_pairs in ((1000, 200000), (2000, 50000), (5000, 16000), (10000, 4000)):
            nb = 400
            base = [synth.random_dna(rng, L) for _ in range(nb)]
            muts = [synth.mutate(rng, x, 0.01) or "A" for x in base]
            pa = (np.arange(n_pairs) % nb).astype(np.int32)
            pb = (nb + pa).astype(np.int32)
            cat, off = hotpath.pack_sequences(base + muts)
            lens = np.diff(off)
            cells = float((lens[pa].astype(np.int64) * lens[pb]).sum())
            full = ctx.edit_distance_pairs(cat, off, pa, pb)["dist"]
            want = oracle.edit_distance_pairs(cat, off, pa[:200], pb[:200], threads=os.cpu_count())
            assert (full[:200] == want).all()
            row = {"len": L, "pairs": n_pairs, "cells": cells, "dist_median": float(np.median(full)), "dist_max": int(full.max())}
            for k in (-1, 16, 64, 128, 256, 480):
                ctx.edit_distance_pairs(cat, off, pa, pb, k=k)
                ctx.profile_enable(True); ctx.profile_reset()
                r = ctx.edit_distance_pairs(cat, off, pa, pb, k=k)
                pr = ctx.profile(); ctx.profile_enable(False)
                exp = full if k < 0 else np.where(full <= k, full, -1)
                assert (r["dist"] == exp).all(), (L, k)
                row["k=%d" % k] = {"dp_ms": pr["dp_ms"], "other_ms": pr["other_ms"], "equivalent_gcups": cells / pr["dp_ms"] / 1e6,
                                   "within_k": float((r["dist"] >= 0).mean())}
            fh.write(json.dumps(row) + "\n"); fh.flush()
            print(L, n_pairs, "median d", row["dist_median"], " | ".join("k=%d %.2f ms (%.0f eq-GCUPS)" % (k, row["k=%d" % k]["dp_ms"], row["k=%d" % k]["equivalent_gcups"]) for k in (-1, 16, 64, 128, 256, 480)))


if __name__ == "__main__":
    main()
