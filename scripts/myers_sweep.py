"""Kernel (b') sweep (BASELINE config 5: interval pairs 50 bp - 5 kb against a CPU Myers): device time of
isof_edit_distance_pairs (distance only and with the path) per length, next to the oracle's Myers/Hyyro restatement
(the "edlib CPU" column; edlib itself is absent and has no call site in the reference) on all host cores.
    python scripts/myers_sweep.py [--out gpurun_out/myers_sweep.jsonl]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/myers_sweep.jsonl")
    ap.add_argument("--lengths", default="50,100,200,500,1000,2000,5000")
    ap.add_argument("--cells", type=float, default=2e11)
    ap.add_argument("--err", type=float, default=0.05)
    ap.add_argument("--tbwarp", type=int, default=2, help="isof_set_traceback_warp mode for the path walk")
    ap.add_argument("--related", type=float, default=0.25, help="fraction of pairs (x, noisy copy of x); the rest are unrelated")
    args = ap.parse_args()
    import oracle
    from isonform_b200 import _lib, hotpath, synth
    ctx = _lib.Context(0)
    ctx.set_traceback_warp(args.tbwarp)
    rng = np.random.default_rng(5)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as fh:
        for L in [int(x) for x in args.lengths.split(",")]:
            nb = 2000 if L <= 1000 else 400
            n_pairs = int(max(2000, min(2_000_000, args.cells / (L * L))))
            base = [synth.random_dna(rng, L) for _ in range(nb)]
            muts = [synth.mutate(rng, x, args.err) or "A" for x in base]
            pa = (np.arange(n_pairs) % nb).astype(np.int32)
            pb = (nb + (np.arange(n_pairs) * 7 % nb)).astype(np.int32)        # mostly unrelated pairs, 1/nb related
            rel = rng.random(n_pairs) < args.related                            # a sequence and its noisy copy
            pb[rel] = nb + pa[rel]
            cat, off = hotpath.pack_sequences(base + muts)
            lens = np.diff(off)
            cells = float((lens[pa].astype(np.int64) * lens[pb]).sum())
            row = {"len": L, "pairs": n_pairs, "cells": cells, "related": args.related}
            for name, want in (("distance", False), ("path", True)):
                ctx.edit_distance_pairs(cat, off, pa, pb, want_cigar=want)
                ctx.profile_enable(True); ctx.profile_reset()
                t0 = time.perf_counter()
                r = ctx.edit_distance_pairs(cat, off, pa, pb, want_cigar=want)
                wall = time.perf_counter() - t0
                pr = ctx.profile()
                ctx.profile_enable(False)
                dev_ms = pr["dp_ms"] + pr["traceback_ms"] + pr["other_ms"]
                row[name] = {"dp_ms": pr["dp_ms"], "traceback_ms": pr["traceback_ms"], "other_ms": pr["other_ms"],
                             "dp_gcups": cells / pr["dp_ms"] / 1e6, "device_gcups": cells / dev_ms / 1e6,
                             "device_pairs_per_s": n_pairs / dev_ms * 1e3, "host_pairs_per_s": n_pairs / wall,
                             "trace_bytes": pr["trace_bytes"]}
            ns = int(max(200, min(n_pairs, 3e9 / (L * L))))
            th = os.cpu_count()
            oracle.edit_distance_pairs(cat, off, pa[:ns], pb[:ns], threads=th)
            t0 = time.perf_counter()
            want = oracle.edit_distance_pairs(cat, off, pa[:ns], pb[:ns], threads=th)
            cw = time.perf_counter() - t0
            assert (r["dist"][:ns] == want).all()
            ccells = float((lens[pa[:ns]].astype(np.int64) * lens[pb[:ns]]).sum())
            row["myers_cpu"] = {"pairs": ns, "threads": th, "pairs_per_s": ns / cw, "gcups": ccells / cw / 1e9}
            fh.write(json.dumps(row) + "\n"); fh.flush()
            print(json.dumps(row))


if __name__ == "__main__":
    main()
