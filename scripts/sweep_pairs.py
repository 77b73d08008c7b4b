#!/usr/bin/env python
"""BASELINE config 5: interval-pair alignment sweep 50 bp - 5 kb (pairs (x, mutate(x, eps))).

    python scripts/sweep_pairs.py [--out profiles/rNN_sweep.json] [--cpu]

Reports pairs/s and GCUPS through the host-buffer C ABI (isof_sg_align_pairs incl. H2D/D2H) per
(length, eps, score set).  With --cpu the oracle port (all host cores) runs on a bounded sample, and so does
the Myers/Hyyro edit-distance restatement standing in for config 5's edlib CPU column."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isonform_b200 import _lib, hotpath, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--lengths", default="50,100,200,500,1000,2000,5000")
    args = ap.parse_args()
    # under torchrun (WORLD_SIZE > 1): weak scaling, every rank aligns its own copy of the pair list on its own GPU,
    # no data-path collective; pairs/s = world * pairs / max-over-ranks time
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    ctx = _lib.Context(int(os.environ.get("LOCAL_RANK", "0")))
    rows = []
    rng = np.random.default_rng(5)
    for L in [int(x) for x in args.lengths.split(",")]:
        n_pairs = int(max(1600, min(400000, 4e10 / (L * L))))
        for eps in (0.01, 0.05, 0.15):
            nb = min(n_pairs, 2000)
            base = [synth.random_dna(rng, L) for _ in range(nb)]
            muts = [synth.mutate(rng, x, eps) or "A" for x in base]
            seqs = base + muts                  # pair p = (orig[p % nb], mutated copy of it)
            pa = (np.arange(n_pairs) % nb).astype(np.int32)
            pb = (nb + (np.arange(n_pairs) % nb)).astype(np.int32)
            cat, off = hotpath.pack_sequences(seqs)
            lens = np.diff(off)
            cells = float((lens[pa].astype(np.int64) * lens[pb]).sum())
            for name, sc in (("bubble", (2, -8, 12, 1)), ("isoform", (2, -2, 12, 1))):
                ctx.sg_align_pairs(cat, off, pa[:1000], pb[:1000], *sc)            # warm-up
                if dist is not None:
                    dist.barrier()
                t0 = time.perf_counter()
                res = ctx.sg_align_pairs(cat, off, pa, pb, *sc)
                dt = time.perf_counter() - t0
                if dist is not None:
                    tmax = torch.tensor([dt], dtype=torch.float64, device="cuda")
                    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
                    dt = float(tmax[0])
                row = {"len": L, "eps": eps, "scores": name, "n_gpus": world, "pairs": n_pairs * world,
                       "pairs_per_s": world * n_pairs / dt, "gcups": world * cells / dt / 1e9, "cigar_runs": int(len(res["cigar"]))}
                # device time of the same call (all kernels, CUDA events): what is left when the pageable-memory
                # numpy wrapper (allocation, H2D/D2H of unpinned buffers) is taken out
                ctx.profile_enable(True); ctx.profile_reset()
                ctx.sg_align_pairs(cat, off, pa, pb, *sc)
                pr = ctx.profile(); ctx.profile_enable(False)
                dev_ms = pr["dp_ms"] + pr["traceback_ms"] + pr["other_ms"]
                if dist is not None:
                    tmax = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
                    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
                    dev_ms = float(tmax[0])
                row["device_ms"] = dev_ms
                row["device_pairs_per_s"] = world * n_pairs / (dev_ms * 1e-3)
                row["device_gcups"] = world * cells / (dev_ms * 1e-3) / 1e9
                if args.cpu:
                    import oracle
                    ns = int(max(16 * (os.cpu_count() or 1), min(n_pairs, 1e10 / (L * L))))
                    t0 = time.perf_counter()
                    oracle.sg_align_pairs(cat, off, pa[:ns], pb[:ns], *sc, threads=os.cpu_count(), simd=True)
                    dtc = time.perf_counter() - t0
                    row["cpu_pairs_per_s"] = ns / dtc
                    row["cpu_cores"] = oracle.num_threads()
                    row["speedup"] = row["pairs_per_s"] / row["cpu_pairs_per_s"]
                    # the "edlib CPU" column of config 5: unit-cost global edit DISTANCE only (no traceback, no
                    # affine gaps), Myers/Hyyro bit-vector restatement -- a different, much cheaper problem
                    nm = int(min(n_pairs, max(ns, 2e11 / (L * L))))
                    t0 = time.perf_counter()
                    oracle.edit_distance_pairs(cat, off, pa[:nm], pb[:nm], threads=os.cpu_count())
                    row["myers_cpu_pairs_per_s"] = nm / (time.perf_counter() - t0)
                rows.append(row)
                if rank == 0:
                    print(json.dumps(row), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    if args.out and rank == 0:
        json.dump({"config": "C5 pair sweep: wall clock of the numpy host API (isof_sg_align_pairs incl. H2D/D2H and buffer allocation); CPU = oracle SIMD port (same affine alignment with traceback), all host cores; myers_cpu = unit-cost edit distance only (Myers/Hyyro restatement standing in for edlib, which has no call site in the reference)", "rows": rows}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
