#!/usr/bin/env python
"""bench.py -- the isONform hot path on B200: interval alignments/s (+ GCUPS, reads/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload C2|C4|C1]

A *step* is one pass of the hot path over one synthetic isONclust batch (BASELINE.json configs[1]:
1000 ONT-like reads of a 10-isoform 3 kb gene at 5 % error, SURVEY.md section 8d "C2"):
  (a) minimizer extraction over the batch (k=20, w=31),
  (c) the minimizer-pair database + span search (xmin=40, xmax=80, delta_len=5), and
  (b) the all-pairs semi-global affine alignment (2,-2,12,1) of the batch's isoform consensuses with
      traceback and the integer merge-decision features -- the `merge_consensuses` loop of the
      reference (IsoformGeneration.py:473-501), where at 5 % error every read is its own group and no
      row merges, i.e. N(N-1)/2 = 499 500 alignments per batch.
`value` times the step with inputs resident in HBM (device pointers, C ABI `_dev` calls, CIGAR runs
written to a device buffer); `e2e` times the same step through the host-buffer C ABI with pinned host
memory: H2D of the reads and pair list, D2H of what the callers consume (minimizers, span records,
scores, decision features -- the merge decision never reads the CIGAR, so production passes
cigar_runs=NULL and so does this) inside the timed region.

Multi-GPU (torchrun, one rank per GPU).  The primary line is *weak* scaling: every rank runs the step on
its own cluster (seed = base + rank); clusters are independent units, there is no collective on the data
path, NCCL carries the barrier and the max-over-ranks of the timings.  `config3` is the *strong*-scaling
companion (BASELINE.json configs[2]): a fixed, seeded population of clusters of unequal size is sharded
over the ranks with `shard.assign_clusters` (LPT) and every cluster goes through the Python entry points
(`runner.batch_hot_path` = front half + `merge.merge_consensuses`) on the rank's own context.

--impl reference times the CPU restatement of the same step (oracle/, "port": the reference's aligner is
the un-vendored, un-installable parasail; its minimizer / span loops are pure Python) on all host cores
over a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

K_SIZE, W_SIZE = 20, 31
X_MIN, X_MAX = 40, 80            # main:623-625 raises xmin to 2k; xmax default 80
MATCH, MISMATCH, GAP_OPEN, GAP_EXT = 2, -2, 12, 1
DELTA_LEN = 5
DTYPE = "int16x2 (int32 fallback)"          # arithmetic of the DP kernel: two alignments per register in 16-bit halves
# Per-cell constants of sg_dp_kernel measured by ncu (`scripts/ncu_dp_constants.py` writes the file from one
# `ncu --set full` capture of the final kernel on a launch whose cell count the profiled program prints):
# ALU-pipe lane-operations per cell and DRAM bytes per cell.  Fallback = the round-1 hand count / capture.
NCU_CONSTANTS = os.path.join(ROOT, "profiles", "ncu_dp_constants.json")
FALLBACK_CONSTANTS = {"capture": "fallback: SASS hand count (DESIGN.md section 4b) and profiles/r01_ncu_v10_summary.txt",
                      "alu_lane_ops_per_cell": 5.0, "dram_bytes_per_cell": 0.5605}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2", choices=["C1", "C2", "C4"])
    ap.add_argument("--reads", type=int, default=None, help="override the batch size (development only)")
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("--cpu-sample-pairs", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip config3 / e2e_entry / C4 / C5 (development only)")
    ap.add_argument("--config3-clusters", type=int, default=64)
    ap.add_argument("--trace-budget-gb", type=float, default=None,
                    help="alignment trace scratch in GB (default: the library's 40-55 %% of free HBM)")
    return ap.parse_args()


def make_workload(workload, rank, n_reads=None, seed=None):
    from isonform_b200 import synth
    from isonform_b200._lib import pack_sequences
    from isonform_b200.polya import remove_read_polyA_ends
    base_seed = synth.CONFIGS[workload]["seed"] if seed is None else seed
    if n_reads is None and workload == "C4":
        n_reads = 1000                      # one 1000-read batch of the 2000-read cluster
    # rank 0 gets the canonical cluster; other ranks get clusters of the same gene family with re-drawn reads
    # (same size statistics, so that "weak scaling" compares like with like)
    reads, _iso, _truth = synth.make_config(workload, seed=base_seed, n_reads=n_reads,
                                            read_seed=None if rank == 0 else rank)
    seqs = [remove_read_polyA_ends(s, 12, 1) for _, s in reads]       # main:351
    seqs_sorted = sorted(seqs, key=len)                                # merge_consensuses sorts by length (:471)
    cat, off = pack_sequences(seqs_sorted)
    # all (shorter, longer) = (consensus1, consensus2) pairs, grouped by the longer sequence (the reference of
    # the alignment): consecutive pairs then share their reference, which the DP kernel exploits
    ib, ia = np.tril_indices(len(seqs_sorted), -1)
    cells = int((np.diff(off)[ia].astype(np.int64) * np.diff(off)[ib].astype(np.int64)).sum())
    cfg = synth.CONFIGS[workload]
    name = "%s: %d-read batch, %s-isoform %d bp gene, eps=%.2f; all-pairs merge_consensuses alignments" % (
        workload, len(seqs), cfg["n_isoforms"], cfg["gene_len"], cfg["eps"])
    return dict(seqs=seqs_sorted, cat=cat, off=off, ia=ia.astype(np.int32), ib=ib.astype(np.int32), cells=cells,
                n_reads=len(seqs), n_pairs=len(ia), n_bases=int(off[-1]), name=name, reads=reads)


def config_dict(wl):
    """The `config` both arms print (same keys, same values)."""
    return {"workload": wl["name"], "n_reads": wl["n_reads"], "pairs_per_batch": wl["n_pairs"],
            "mean_read_len": wl["n_bases"] / max(wl["n_reads"], 1), "k": K_SIZE, "w": W_SIZE, "xmin": X_MIN, "xmax": X_MAX,
            "scores": [MATCH, MISMATCH, GAP_OPEN, GAP_EXT], "delta_len": DELTA_LEN,
            "parallelism": "1 cluster per GPU, no data-path collective",
            "l2": "per-step trace working set (0.5 B/cell, ~1.9 TB at C2) >> 126 MB L2; no explicit flush"}


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------
# CPU arm (oracle port) -- also used for cpu_baseline
# ----------------------------------------------------------------------------------------------
_SPAN_STAGE_CACHE = {}


def cpu_span_stage_seconds(wl):
    """The reference's span stage (get_minimizer_combinations_database + find_most_supported_span for every read,
    main:174-212,308-338) on the whole batch, timed once: it is pure Python in the reference and runs on ONE core per
    batch (the reference's parallelism is one process per batch), so the oracle's Python restatement is the faithful
    cost.  Bounded: reads are sub-sampled to keep this under ~20 s and the time is scaled by (all reads / sampled reads)^2
    for the database + linearly for the per-read scan -- stated in `sample`."""
    key = id(wl)
    if key in _SPAN_STAGE_CACHE:
        return _SPAN_STAGE_CACHE[key]
    import oracle
    from isonform_b200 import runner
    n_all = wl["n_reads"]
    n_sub = min(n_all, 150)
    recs = [(acc, s, "I" * len(s)) for acc, s in wl["reads"][:n_sub]]
    reads = runner.read_batch(recs)
    t0 = time.perf_counter()
    M = oracle.get_minimizers_and_positions(reads, W_SIZE, K_SIZE)
    t_min = time.perf_counter() - t0
    t0 = time.perf_counter()
    db = oracle.get_minimizer_combinations_database(M, K_SIZE, X_MIN, X_MAX, reads)
    t_db = time.perf_counter() - t0
    t0 = time.perf_counter()
    for r_id in reads:
        oracle.pyside.read_intervals(r_id, M[r_id], db, K_SIZE, X_MIN, X_MAX, DELTA_LEN)
    t_scan = time.perf_counter() - t0
    scale = n_all / float(n_sub)
    # database: linear in reads; scan: occurrences per bucket grow with the read count -> reads^2 / reads_sub^2
    est = t_db * scale + t_scan * scale * scale
    out = {"seconds_full_batch_est": est, "reads_timed": n_sub, "db_s": t_db, "scan_s": t_scan, "python_minimizers_s": t_min}
    _SPAN_STAGE_CACHE[key] = out
    return out


def cpu_step(wl, sample_pairs, threads):
    import oracle
    t0 = time.perf_counter()
    oracle.minimizers_batch(wl["cat"], wl["off"], K_SIZE, W_SIZE, threads=threads)
    t_min = time.perf_counter() - t0
    sel = np.linspace(0, wl["n_pairs"] - 1, num=sample_pairs).astype(np.int64)
    t0 = time.perf_counter()
    oracle.sg_align_pairs(wl["cat"], wl["off"], wl["ia"][sel], wl["ib"][sel], MATCH, MISMATCH, GAP_OPEN, GAP_EXT,
                          threads=threads, want_cigar=True, simd=True)
    t_al = time.perf_counter() - t0
    lens = np.diff(wl["off"])
    cells = int((lens[wl["ia"][sel]].astype(np.int64) * lens[wl["ib"][sel]]).sum())
    # span stage, charged pro rata to the sampled share of the batch's alignments
    t_span = cpu_span_stage_seconds(wl)["seconds_full_batch_est"] * sample_pairs / max(wl["n_pairs"], 1)
    return t_min, t_al, cells, t_span


def pick_cpu_sample(wl, threads, target_s=12.0):
    """Size the pair sample so that one CPU step takes about target_s seconds."""
    import oracle
    unit = oracle.simd_lanes() * threads      # the SIMD port aligns one lane group per thread at a time
    probe = min(unit, wl["n_pairs"])
    _tm, t_al, _c, _ts = cpu_step(wl, probe, threads)
    per_pair = max(t_al / probe, 1e-6)
    want = int(max(probe, min(wl["n_pairs"], target_s / per_pair)))
    return min(wl["n_pairs"], max(unit, (want // unit) * unit))


def cpu_sample_text(sample, wl, span):
    import oracle
    return ("%d of %d pairs per step (evenly spaced) + minimizers of all %d reads (C, OpenMP) + the span stage charged pro "
            "rata (oracle Python restatement of main:174-338 timed on %d reads, 1 core, scaled to the batch: %.1f s); "
            "aligner = oracle/isof_oracle.c Gotoh+trace+cigar, %d pairs per thread in int16 SIMD lanes (%s), OpenMP"
            % (sample, wl["n_pairs"], wl["n_reads"], span["reads_timed"], span["seconds_full_batch_est"],
               oracle.simd_lanes(), oracle.simd_isa()))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    wl = make_workload(args.workload, 0, args.reads, args.seed)
    threads = os.cpu_count() or 1
    sample = args.cpu_sample_pairs or pick_cpu_sample(wl, threads, target_s=8.0)
    for _ in range(args.warmup):
        cpu_step(wl, max(threads, sample // 8), threads)
    times, cells, t_align = [], 0, 0.0
    for _ in range(args.steps):
        t_min, t_al, c, t_span = cpu_step(wl, sample, threads)
        times.append(t_min + t_al + t_span)
        t_align += t_al
        cells = c
    t = float(np.sum(times))
    value = sample * args.steps / t
    span = cpu_span_stage_seconds(wl)
    gcups = cells * args.steps / t_align / 1e9
    line = {"impl": "reference", "metric": "interval_alignments_per_s", "value": value, "unit": "alignments/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
            "config": config_dict(wl),
            "gcups": gcups,
            "cpu_baseline": {"value": value, "unit": "alignments/s", "cores": oracle.num_threads(), "kind": "port",
                             "aligner_gcups": gcups, "aligner_gcups_per_core": gcups / max(oracle.num_threads(), 1),
                             "sample": cpu_sample_text(sample, wl, span)},
            "e2e": {"value": value, "unit": "alignments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# GPU arm: the step on device-resident and on host buffers
# ----------------------------------------------------------------------------------------------
class DeviceStep:
    """One workload resident in HBM (torch owns the memory, the library gets raw pointers)."""

    def __init__(self, ctx, wl, dev, want_cigar=True):
        import torch
        from isonform_b200 import _lib
        self.ctx, self.wl, self._lib = ctx, wl, _lib
        P, R = wl["n_pairs"], wl["n_reads"]
        t = torch
        self.d_cat = t.from_numpy(np.array(wl["cat"])).to(dev)
        self.d_off = t.from_numpy(np.array(wl["off"])).to(dev)
        self.d_ia = t.from_numpy(np.array(wl["ia"])).to(dev)
        self.d_ib = t.from_numpy(np.array(wl["ib"])).to(dev)
        self.d_score = t.empty(P, dtype=t.int32, device=dev)
        self.d_endq = t.empty(P, dtype=t.int32, device=dev)
        self.d_endr = t.empty(P, dtype=t.int32, device=dev)
        self.d_coff = t.empty(P + 1, dtype=t.int64, device=dev)
        self.d_feat = t.empty((P, _lib.N_FEATURES), dtype=t.int32, device=dev)
        self.d_mpos = t.empty(wl["n_bases"] + R, dtype=t.int32, device=dev)
        self.d_moff = t.empty(R + 1, dtype=t.int64, device=dev)
        self.rec_cap = 8 * (wl["n_bases"] // 5) + 1024                 # ~0.15 minimizers/base, ~6 partners each
        self.d_rec = [t.empty(self.rec_cap, dtype=t.int32, device=dev) for _ in range(6)]
        self.d_boff = t.empty(self.rec_cap + 1, dtype=t.int32, device=dev)
        self.d_bdrop = t.empty(self.rec_cap, dtype=t.uint8, device=dev)
        self.want_cigar = want_cigar
        self.cig_cap = max(64 * P, 1 << 20) if want_cigar else 0
        self.d_cig = t.empty(max(self.cig_cap, 1), dtype=t.int32, device=dev)
        self.dev = dev
        self.n_rec = 0

    def run(self):
        import torch
        wl, ctx = self.wl, self.ctx
        P, R = wl["n_pairs"], wl["n_reads"]
        n_min = ctx.minimizers_batch_dev(self.d_cat.data_ptr(), self.d_off.data_ptr(), R, wl["n_bases"], K_SIZE, W_SIZE,
                                         self.d_mpos.data_ptr(), self.d_mpos.numel(), self.d_moff.data_ptr())
        self.n_rec, _n_b = ctx.spans_batch_dev(self.d_cat.data_ptr(), self.d_off.data_ptr(), R, self.d_mpos.data_ptr(),
                                               self.d_moff.data_ptr(), n_min, K_SIZE, X_MIN, X_MAX, DELTA_LEN, self.rec_cap,
                                               *[x.data_ptr() for x in self.d_rec], self.d_boff.data_ptr(),
                                               self.d_bdrop.data_ptr())
        while True:
            try:
                runs = ctx.sg_align_pairs_dev(self.d_cat.data_ptr(), self.d_off.data_ptr(), R, wl["n_bases"],
                                              self.d_ia.data_ptr(), self.d_ib.data_ptr(), P, MATCH, MISMATCH, GAP_OPEN, GAP_EXT,
                                              self.d_score.data_ptr(), self.d_endq.data_ptr(), self.d_endr.data_ptr(),
                                              self.d_cig.data_ptr() if self.want_cigar else None, self.cig_cap,
                                              self.d_coff.data_ptr(), self.d_feat.data_ptr(), DELTA_LEN)
                return n_min, runs
            except self._lib.IsofError as e:
                if e.code != self._lib.ERR_CAPACITY:
                    raise
                need = int(self.d_coff[P].item())
                self.cig_cap = need + need // 16
                self.d_cig = torch.empty(self.cig_cap, dtype=torch.int32, device=self.dev)


class HostStep:
    """The same step through the host-buffer C ABI: pinned inputs, H2D / D2H inside the call."""

    def __init__(self, ctx, wl, n_min_hint):
        import ctypes as C
        import torch
        from isonform_b200 import _lib
        self.C, self.ctx, self.wl = C, ctx, wl
        P, R = wl["n_pairs"], wl["n_reads"]
        t = torch
        pin = lambda a: t.from_numpy(np.array(a)).pin_memory()          # noqa: E731
        self.h_cat, self.h_off, self.h_ia, self.h_ib = pin(wl["cat"]), pin(wl["off"]), pin(wl["ia"]), pin(wl["ib"])
        self.h_score = t.empty(P, dtype=t.int32).pin_memory()
        self.h_feat = t.empty((P, _lib.N_FEATURES), dtype=t.int32).pin_memory()
        self.h_coff = t.empty(P + 1, dtype=t.int64).pin_memory()
        self.h_mpos = t.empty(int(n_min_hint) + 1024, dtype=t.int32).pin_memory()
        self.h_moff = t.empty(R + 1, dtype=t.int64).pin_memory()
        self.rec_cap = 8 * (wl["n_bases"] // 5) + 1024
        self.h_rec = [t.empty(self.rec_cap, dtype=t.int32).pin_memory() for _ in range(6)]
        self.h_boff = t.empty(self.rec_cap + 1, dtype=t.int32).pin_memory()
        self.h_bdrop = t.empty(self.rec_cap, dtype=t.uint8).pin_memory()
        self.n_out, self.n_rec_h, self.n_b_h = C.c_int64(0), C.c_int64(0), C.c_int64(0)

    def run(self):
        from isonform_b200.hotpath import _merge_from_features
        C, ctx, wl = self.C, self.ctx, self.wl
        L = ctx._L
        P, R = wl["n_pairs"], wl["n_reads"]
        ctx._check(L.isof_minimizers_batch(ctx._h, self.h_cat.data_ptr(), self.h_off.data_ptr(), R, K_SIZE, W_SIZE,
                                           self.h_mpos.data_ptr(), self.h_mpos.numel(), self.h_moff.data_ptr(),
                                           C.byref(self.n_out)))
        ctx._check(L.isof_spans_batch(ctx._h, self.h_cat.data_ptr(), self.h_off.data_ptr(), R, self.h_mpos.data_ptr(),
                                      self.h_moff.data_ptr(), K_SIZE, X_MIN, X_MAX, DELTA_LEN, self.rec_cap,
                                      *[x.data_ptr() for x in self.h_rec], self.h_boff.data_ptr(), self.h_bdrop.data_ptr(),
                                      C.byref(self.n_rec_h), C.byref(self.n_b_h)))
        # production form of the merge stage (hotpath.align_to_merge_pairs): features only, no CIGAR buffer
        ctx._check(L.isof_sg_features_pairs(ctx._h, self.h_cat.data_ptr(), self.h_off.data_ptr(), R, self.h_ia.data_ptr(),
                                            self.h_ib.data_ptr(), P, MATCH, MISMATCH, GAP_OPEN, GAP_EXT, DELTA_LEN,
                                            self.h_score.data_ptr(), self.h_feat.data_ptr(), None, 0, self.h_coff.data_ptr()))
        # the step's result as the caller consumes it: merge decisions from the features
        return int(_merge_from_features(self.h_feat.numpy(), 0.15, DELTA_LEN, 30, 50).sum())


def load_constants():
    try:
        with open(NCU_CONSTANTS) as fh:
            c = json.load(fh)
        if c.get("alu_lane_ops_per_cell") and c.get("dram_bytes_per_cell"):
            return c
    except Exception:
        pass
    return dict(FALLBACK_CONSTANTS)


# ----------------------------------------------------------------------------------------------
# companions of the primary line (rank 0 prints them inside the same JSON line)
# ----------------------------------------------------------------------------------------------
def config3_population(n_clusters, seed=33):
    """Fixed, seeded cluster population of unequal size (reads per cluster), the same for every N."""
    rng = np.random.default_rng(seed)
    sizes = rng.integers(40, 361, size=n_clusters)
    sizes[: max(1, n_clusters // 16)] = rng.integers(420, 561, size=max(1, n_clusters // 16))     # a few heavy clusters
    return {int(c): int(s) for c, s in enumerate(sizes)}


def run_config3(ctx, rank, world, n_clusters, barrier, allreduce):
    """BASELINE.json configs[2] at bench scale: clusters sharded by cluster id across the ranks (strong scaling)."""
    from isonform_b200 import runner, shard, synth
    pop = config3_population(n_clusters)
    plan = shard.assign_clusters(pop, world)
    mine = plan[rank]
    batches = []
    for c in mine:
        reads, _i, _t = synth.make_cluster(seed=1000 + c, gene_len=3000, n_isoforms=10, n_reads=pop[c], eps=0.05)
        batches.append(runner.read_batch([(a, s, "I" * len(s)) for a, s in reads]))
    # warm-up on the smallest cluster (allocations, first-use costs)
    if batches:
        runner.batch_hot_path(min(batches, key=len), K_SIZE, W_SIZE, X_MIN, X_MAX, DELTA_LEN, ctx=ctx)
    barrier()
    stats = {}
    t0 = time.perf_counter()
    for rd in batches:
        runner.batch_hot_path(rd, K_SIZE, W_SIZE, X_MIN, X_MAX, DELTA_LEN, ctx=ctx, stats=stats)
    busy = time.perf_counter() - t0
    barrier()
    wall = time.perf_counter() - t0
    tot = allreduce([float(stats.get("pairs_aligned", 0)), float(stats.get("reads", 0)), float(stats.get("front_half_s", 0.0)),
                     float(stats.get("merge_s", 0.0))], "sum")
    mx = allreduce([busy, wall], "max")
    mn = allreduce([-busy], "max")
    cost = [sum(shard.cluster_cost(pop[c]) for c in p) for p in plan]
    return {"workload": "config 3 at bench scale: %d clusters of 40-560 reads (seeded, 3 kb gene, 10 isoforms, eps 0.05), one batch each, "
                        "sharded by cluster id with shard.assign_clusters (LPT on reads^2); per cluster: runner.batch_hot_path = "
                        "minimizers + pair DB + span search + WIS picks, then merge.merge_consensuses over the reads as singleton groups"
                        % n_clusters,
            "scaling": "strong", "n_gpus": world, "clusters": n_clusters, "reads": int(tot[1]),
            "value": tot[0] / mx[1], "unit": "alignments/s", "reads_per_s": tot[1] / mx[1], "seconds": mx[1],
            "alignments": int(tot[0]), "busy_s_max": mx[0], "busy_s_min": -mn[0],
            "lpt_imbalance": max(cost) / (sum(cost) / float(world)) if sum(cost) else 1.0,
            "front_half_s_sum_over_ranks": tot[2], "merge_s_sum_over_ranks": tot[3],
            "note": "wall clock of the slowest rank after a barrier (host Python + H2D/D2H + kernels); no collective on the data path"}


def run_entry(ctx, wl):
    """reads/s through the Python entry points for the primary workload's batch (real ctx, host strings in,
    Python objects out): what a caller of the drop-in seams gets, host half included."""
    import torch
    from isonform_b200 import runner
    reads = runner.read_batch([(a, s, "I" * len(s)) for a, s in wl["reads"]])
    runner.batch_hot_path(reads, K_SIZE, W_SIZE, X_MIN, X_MAX, DELTA_LEN, ctx=ctx, stats={})       # warm-up pass (scratch growth)
    stats = {}
    ctx.profile_reset()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    runner.batch_hot_path(reads, K_SIZE, W_SIZE, X_MIN, X_MAX, DELTA_LEN, ctx=ctx, stats=stats)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    prof = ctx.profile()
    return {"reads_per_s": len(reads) / dt, "alignments_per_s": stats["pairs_aligned"] / dt, "seconds": dt,
            "front_half_s": stats["front_half_s"], "merge_s": stats["merge_s"], "alignments": stats["pairs_aligned"],
            "library_launches": int(prof["launches"]),
            "api": "runner.batch_hot_path(reads): get_minimizers_and_positions + SpanIndex + WIS picks (front half), "
                   "merge.merge_consensuses over singleton groups (4 speculative launches at C2)"}


def run_c4_sample(ctx, dev, n_reads=200):
    """BASELINE config 4 (10 kb reads) at a bounded size: 200 reads of the C4 cluster = 19 900 pairs."""
    import torch
    wl = make_workload("C4", 0, n_reads)
    st = DeviceStep(ctx, wl, dev, want_cigar=False)
    st.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream = torch.cuda.current_stream()
    ctx.profile_reset()
    e0.record(stream)
    st.run()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return {"workload": wl["name"], "pairs": wl["n_pairs"], "mean_read_len": wl["n_bases"] / wl["n_reads"], "ms_per_step": ms,
            "alignments_per_s": wl["n_pairs"] / ms * 1e3, "gcups": wl["cells"] / ms / 1e6,
            "call": "feature-only (as merge_consensuses issues it): banded trace, exact via second pass",
            "band_pairs_retried": ctx.profile()["band_pairs_retried"]}


def run_c5_points(ctx, rank=0, world=1, barrier=None, allreduce=None, lengths=(50, 100, 200, 500, 2000)):
    """BASELINE config 5 at three-plus points: pairs (x, mutate(x, 5 %)), bubble scores, all kernels of the call.
    N > 1: every rank aligns its own shard of the same size (pairs are independent: weak scaling, no collective); rates are
    the sum over ranks against the slowest rank's time."""
    from isonform_b200 import hotpath, synth
    rng = np.random.default_rng(5 + rank)
    rows = []
    for L in lengths:
        nb = 2000
        n_pairs = int(max(4000, min(400000, 4e10 / (L * L))))
        base = [synth.random_dna(rng, L) for _ in range(nb)]
        muts = [synth.mutate(rng, x, 0.05) or "A" for x in base]
        pa = (np.arange(n_pairs) % nb).astype(np.int32)
        pb = (nb + (np.arange(n_pairs) % nb)).astype(np.int32)
        cat, off = hotpath.pack_sequences(base + muts)
        lens = np.diff(off)
        cells = float((lens[pa].astype(np.int64) * lens[pb]).sum())
        ctx.sg_align_pairs(cat, off, pa, pb, 2, -8, 12, 1)
        ctx.profile_enable(True); ctx.profile_reset()
        if barrier:
            barrier()
        t0 = time.perf_counter()
        ctx.sg_align_pairs(cat, off, pa, pb, 2, -8, 12, 1)
        wall = time.perf_counter() - t0
        pr = ctx.profile(); ctx.profile_enable(False)
        dev_ms = pr["dp_ms"] + pr["traceback_ms"] + pr["other_ms"]
        dp_ms, tb_ms, pairs_all = pr["dp_ms"], pr["traceback_ms"], float(n_pairs)
        if allreduce and world > 1:
            dev_ms, dp_ms, tb_ms, wall = allreduce([dev_ms, dp_ms, tb_ms, wall], "max")
            cells, pairs_all = allreduce([cells, pairs_all], "sum")
        rows.append({"len": L, "pairs": int(pairs_all), "n_gpus": world, "device_ms": dev_ms, "device_pairs_per_s": pairs_all / dev_ms * 1e3,
                     "device_gcups": cells / dev_ms / 1e6, "host_pairs_per_s": pairs_all / wall, "dp_ms": dp_ms,
                     "traceback_ms": tb_ms})
    return rows


def run_c5_edit_distance(ctx, int32_peak, hbm_peak, rank=0, world=1, barrier=None, allreduce=None,
                         lengths=(50, 100, 200, 500, 1000, 2000, 5000)):
    """BASELINE config 5, the "vs edlib CPU" column: kernel (b'), unit-cost global edit distance (Myers/Hyyro bit-parallel,
    csrc/myers.cu) on pairs (x, mutate(x, 5 %)), distances only and with the path, next to the oracle's Myers restatement
    on all host cores over a bounded sample (edlib itself is absent and has no call site in the reference).
    N > 1 (config 5 asks for 1/2/4/8 GPUs): pairs are independent, every rank runs its own shard of the same size (weak
    scaling, no collective); the rates are the sum over ranks against the slowest rank's device time."""
    from isonform_b200 import hotpath, synth
    rng = np.random.default_rng(5 + rank)
    threads = os.cpu_count() or 1
    try:
        kc = json.load(open(os.path.join(ROOT, "profiles", "ncu_myers_constants.json")))["kernels"]
    except Exception:
        kc = None
    rows = []
    for L in lengths:
        nb = 2000 if L <= 1000 else 400
        n_pairs = int(max(4000, min(2000000, 1e11 / (L * L))))
        base = [synth.random_dna(rng, L) for _ in range(nb)]
        muts = [synth.mutate(rng, x, 0.05) or "A" for x in base]
        pa = (np.arange(n_pairs) % nb).astype(np.int32)
        pb = (nb + (np.arange(n_pairs) % nb)).astype(np.int32)
        cat, off = hotpath.pack_sequences(base + muts)
        lens = np.diff(off)
        cells = float((lens[pa].astype(np.int64) * lens[pb]).sum())
        row = {"len": L, "pairs": n_pairs * world, "n_gpus": world}
        for name, want in (("distance", False), ("path", True)):
            ctx.edit_distance_pairs(cat, off, pa, pb, want_cigar=want)
            ctx.profile_enable(True); ctx.profile_reset()
            if barrier:
                barrier()
            t0 = time.perf_counter()
            res = ctx.edit_distance_pairs(cat, off, pa, pb, want_cigar=want)
            wall = time.perf_counter() - t0
            pr = ctx.profile(); ctx.profile_enable(False)
            dev_ms = pr["dp_ms"] + pr["traceback_ms"] + pr["other_ms"]
            dp_ms, tb_ms = pr["dp_ms"], pr["traceback_ms"]
            cells_all, pairs_all = cells, float(n_pairs)
            if allreduce and world > 1:
                dev_ms, dp_ms, tb_ms, wall = allreduce([dev_ms, dp_ms, tb_ms, wall], "max")
                cells_all, pairs_all = allreduce([cells, float(n_pairs)], "sum")
            row[name] = {"device_ms": dev_ms, "dp_ms": dp_ms, "traceback_ms": tb_ms,
                         "device_gcups": cells_all / dev_ms / 1e6, "dp_gcups": cells_all / dp_ms / 1e6,
                         "device_pairs_per_s": pairs_all / dev_ms * 1e3, "host_pairs_per_s": pairs_all / wall}
        if L >= 2000:
            # the same distances under a bound k (edlib's k): long pairs then run inside the Ukkonen band | i - j | <= k
            kb = int(min(480, 0.12 * L))
            ctx.edit_distance_pairs(cat, off, pa, pb, k=kb)
            ctx.profile_enable(True); ctx.profile_reset()
            rk = ctx.edit_distance_pairs(cat, off, pa, pb, k=kb)
            prk = ctx.profile(); ctx.profile_enable(False)
            if not (rk["dist"] == np.where(res["dist"] <= kb, res["dist"], -1)).all():
                raise SystemExit("banded edit-distance parity failure at L=%d" % L)
            dpk = prk["dp_ms"]
            if allreduce and world > 1:
                dpk = allreduce([dpk], "max")[0]
            row["distance_bounded"] = {"k": kb, "dp_ms": dpk, "speedup_vs_unbounded": row["distance"]["dp_ms"] / dpk,
                                       "within_k": float((rk["dist"] >= 0).mean())}
        if kc and int32_peak:
            # distances only: INT32-ALU bound (ncu: ALU pipe 97 % of active cycles); with the path: bound by the trace write
            ops = kc["distance"]["alu_lane_ops_per_cell"]
            row["distance"]["roofline"] = {"bound": "int32_alu", "achieved": row["distance"]["dp_gcups"] * 1e9 * ops / 1e12 / world,
                                           "peak": int32_peak / 1e12, "unit": "Tlane-op/s per GPU",
                                           "frac": row["distance"]["dp_gcups"] * 1e9 * ops / int32_peak / world,
                                           "alu_lane_ops_per_cell": ops,
                                           "note": "ops per cell from the ncu capture at 1000 bp (eight-word lanes, 4 lanes per pair); "
                                                   "shorter pairs pad their words (50 bp: 50 of 64 rows) and execute more per useful cell, "
                                                   "so their frac understates the pipe's occupancy"}
            gbs = 0.25 * cells / (row["path"]["dp_ms"] * 1e-3) / 1e9
            row["path"]["roofline"] = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                                       "algorithmic_bytes_per_cell": 0.25, "traffic_bytes_per_cell": kc["path"]["dram_bytes_per_cell"]}
        if rank != 0:
            continue
        import oracle
        ns = int(max(200, min(n_pairs, 2e9 / (L * L))))
        oracle.edit_distance_pairs(cat, off, pa[:ns], pb[:ns], threads=threads)
        t0 = time.perf_counter()
        want_d = oracle.edit_distance_pairs(cat, off, pa[:ns], pb[:ns], threads=threads)
        cw = time.perf_counter() - t0
        if not (res["dist"][:ns] == want_d).all():
            raise SystemExit("edit-distance parity failure in the C5 sweep at L=%d" % L)
        ccells = float((lens[pa[:ns]].astype(np.int64) * lens[pb[:ns]]).sum())
        row["myers_cpu"] = {"pairs": ns, "threads": threads, "pairs_per_s": ns / cw, "gcups": ccells / cw / 1e9,
                            "kind": "port (Myers/Hyyro 64-bit blocks, oracle/isof_oracle.c), distances only"}
        rows.append(row)
    return rows


def run_seam_latency(ctx):
    """Per-call latency of the one-pair seam (consensus.parasail_alignment on a memo miss)."""
    from isonform_b200 import hotpath, synth
    rng = np.random.default_rng(1)
    out = {}
    for L in (100, 500, 2000):
        x = synth.random_dna(rng, L); y = synth.mutate(rng, x, 0.05)
        hotpath.parasail_alignment(x, y, 2, -8, 12, 1, ctx=ctx)
        n = 200 if L < 1000 else 50
        t0 = time.perf_counter()
        for _ in range(n):
            hotpath.parasail_alignment(x, y, 2, -8, 12, 1, ctx=ctx)
        out["us_per_call_%dbp" % L] = (time.perf_counter() - t0) / n * 1e6
        t0 = time.perf_counter()
        for _ in range(n):
            ctx.sg_align_one(x, y, 2, -8, 12, 1)
        out["us_per_c_abi_call_%dbp" % L] = (time.perf_counter() - t0) / n * 1e6
    out["note"] = ("hotpath.parasail_alignment = C ABI call + CIGAR string / aligned strings in Python; calls of <= 32 pairs up to "
                   "4096 x 4096 take the fused one-launch kernel (trace in shared memory up to ~500 x 500, warp-cooperative "
                   "traceback, one warp per 256-row stripe chasing its predecessor for multi-stripe pairs)")
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist
    from isonform_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allreduce(vals, op):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
        return [float(x) for x in t.tolist()]

    wl = make_workload(args.workload, rank, args.reads, args.seed)
    ctx = _lib.Context(local)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    if args.trace_budget_gb:
        ctx.set_trace_budget(int(args.trace_budget_gb * (1 << 30)))
    P, R = wl["n_pairs"], wl["n_reads"]
    step = DeviceStep(ctx, wl, dev, want_cigar=True)

    # ---- warm-up (also sizes the CIGAR buffer)
    n_min = n_runs = 0
    for _ in range(max(args.warmup, 1)):
        n_min, n_runs = step.run()
    int32_peak = ctx.int32_peak()

    # ---- timed: device-resident (production configuration: three-stream chunk pipeline, no per-launch events)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ctx.profile_enable(False)
    ctx.profile_reset()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        n_min, n_runs = step.run()
    e1.record(stream)
    barrier()
    ms_dev = e0.elapsed_time(e1)
    launches_timed = ctx.profile()["launches"]
    n_rec = step.n_rec

    # ---- kernel step: the same step once more with the chunks serialised on one stream and every launch bracketed by
    # CUDA events on that stream, so that per-launch durations are exact (in the pipelined configuration consecutive DP
    # launches overlap by design -- the tail of one chunk is filled by the next -- and individual durations are not
    # meaningful).  Not part of `value`; it is the lower bound of the kernel's rate (tails exposed), `frac_step` below
    # (whole timed step, every kernel included) the upper bound.
    ctx.set_overlap(False)
    ctx.profile_reset()
    ctx.profile_enable(True)
    barrier()
    step.run()
    barrier()
    prof = ctx.profile()
    ctx.profile_enable(False)
    ctx.set_overlap(True)

    # ---- timed: end to end through the host-buffer C ABI, pinned host memory
    host = HostStep(ctx, wl, n_min)
    host.run()
    ctx.profile_reset()
    barrier()
    t0 = time.perf_counter()
    merges = 0
    for _ in range(args.steps):
        merges = host.run()
    barrier()
    wall_e2e = time.perf_counter() - t0
    prof_e2e = ctx.profile()
    clk = clocks.stop() if rank == 0 else None

    # ---- max over ranks
    ms_dev_max, ms_e2e_max = allreduce([ms_dev, wall_e2e * 1e3], "max")
    pairs_all, cells_all, reads_all = allreduce([float(P), float(wl["cells"]), float(R)], "sum")

    extras = {}
    try:
        hbm_peak_early = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0))
    except Exception:
        hbm_peak_early = 6650.0
    if not args.no_extras:
        del step
        torch.cuda.empty_cache()
        extras["config3"] = run_config3(ctx, rank, world, args.config3_clusters, barrier, allreduce)
        ctx.release_scratch()          # the primary step's trace lanes (up to 80 % of HBM for long reads) go back before kernel (b') allocates its own
        c5e = run_c5_edit_distance(ctx, int32_peak, hbm_peak_early, rank, world, barrier, allreduce)
        ctx.release_scratch()
        c5a = run_c5_points(ctx, rank, world, barrier, allreduce)
        ctx.release_scratch()
        if rank == 0:
            extras["C5_edit_distance"] = c5e
            extras["C5"] = c5a
        if rank == 0 and world == 1:
            extras["e2e_entry"] = run_entry(ctx, wl)
            extras["seam_latency"] = run_seam_latency(ctx)
            extras["C4"] = run_c4_sample(ctx, dev)

    # ---- kernel (a) at super-batch scale (the C3 regime: many clusters' reads in one call); rank 0, untimed extra
    min_super = None
    if rank == 0:
        reps = max(1, min(32, int(2.0e8 // max(wl["n_bases"], 1))))
        offs = [wl["off"][:-1] + i * wl["n_bases"] for i in range(reps)] + [np.array([reps * wl["n_bases"]], np.int64)]
        d_cat_s = torch.from_numpy(np.array(wl["cat"])).to(dev).repeat(reps)
        d_off_s = torch.from_numpy(np.concatenate(offs).astype(np.int64)).to(dev)
        cap_s = reps * wl["n_bases"] // 4 + 1024
        d_pos_s = torch.empty(cap_s, dtype=torch.int32, device=dev)
        d_poff_s = torch.empty(reps * R + 1, dtype=torch.int64, device=dev)
        n_s = 0
        for _ in range(2):
            n_s = ctx.minimizers_batch_dev(d_cat_s.data_ptr(), d_off_s.data_ptr(), reps * R, reps * wl["n_bases"], K_SIZE, W_SIZE,
                                           d_pos_s.data_ptr(), cap_s, d_poff_s.data_ptr())
        torch.cuda.synchronize()
        es0, es1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        es0.record(stream)
        for _ in range(5):
            ctx.minimizers_batch_dev(d_cat_s.data_ptr(), d_off_s.data_ptr(), reps * R, reps * wl["n_bases"], K_SIZE, W_SIZE,
                                     d_pos_s.data_ptr(), cap_s, d_poff_s.data_ptr())
        es1.record(stream)
        torch.cuda.synchronize()
        ms_s = es0.elapsed_time(es1) / 5
        min_super = {"n_reads": reps * R, "mbases": reps * wl["n_bases"] / 1e6, "minimizers": int(n_s), "ms_per_call": ms_s,
                     "mbases_per_s": reps * wl["n_bases"] / ms_s / 1e3}
        del d_cat_s, d_off_s, d_pos_s, d_poff_s

    if rank == 0:
        K = args.steps
        value = pairs_all * K / (ms_dev_max * 1e-3)
        e2e_value = pairs_all * K / (ms_e2e_max * 1e-3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        consts = load_constants()
        n_launch = max(prof["dp_launches"], 1)
        dp_ms_launch = prof["dp_ms"] / n_launch
        cells_launch = prof["dp_cells"] / n_launch
        ops_cell = float(consts["alu_lane_ops_per_cell"])
        cups = prof["dp_cells"] / (prof["dp_ms"] * 1e-3)
        achieved_ops = cups * ops_cell
        # algorithmic bytes of the DP kernel: 0.5 B/cell trace write + the two sequences, per launch
        alg_bytes = 0.5 * cells_launch + 2.0 * wl["n_bases"] / n_launch
        achieved_gbs = alg_bytes / (dp_ms_launch * 1e-3) / 1e9
        step_cups = cells_all / world * K / (ms_dev_max * 1e-3)
        packed_frac = prof["dp_cells_packed"] / max(prof["dp_cells"], 1)
        profile_frac = prof["dp_cells_profile"] / max(prof["dp_cells"], 1)
        line = {
            "metric": "interval_alignments_per_s", "value": value, "unit": "alignments/s", "n_gpus": world,
            "steps": K, "warmup": args.warmup, "ms_per_step": ms_dev_max / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
            "config": config_dict(wl),
            "gcups": cells_all * K / (ms_dev_max * 1e-3) / 1e9,
            "reads_per_s": reads_all * K / (ms_dev_max * 1e-3),
            "minimizers_per_step": int(n_min), "minimizer_pairs_per_step": int(n_rec),
            "cigar_runs_per_step": int(n_runs), "merges_per_step": int(merges),
            "e2e": {"value": e2e_value, "unit": "alignments/s",
                    "h2d_bytes_per_step": int(prof_e2e["h2d_bytes"] // K), "d2h_bytes_per_step": int(prof_e2e["d2h_bytes"] // K),
                    "ms_per_step": ms_e2e_max / K, "gcups": cells_all * K / (ms_e2e_max * 1e-3) / 1e9,
                    "note": "host-buffer C ABI, pinned memory; D2H = minimizers, span records, scores, decision features "
                            "(feature-only call, cigar_runs=NULL, as merge.merge_consensuses issues it)"},
            "gpu_launches": int(launches_timed),
            "roofline": {"kernel": "sg_dp_kernel", "bound": "int32_alu", "achieved": achieved_ops / 1e12, "peak": int32_peak / 1e12,
                         "unit": "Tlane-op/s", "frac": achieved_ops / int32_peak if int32_peak else None,
                         "frac_step": step_cups * ops_cell / int32_peak if int32_peak else None,
                         "alu_lane_ops_per_cell": ops_cell, "cell_updates_per_s": cups, "gcups_kernel": cups / 1e9,
                         "launch_ms": dp_ms_launch, "launches": int(prof["dp_launches"]), "cells_per_launch": cells_launch,
                         "traffic": float(consts["dram_bytes_per_cell"]) * cells_launch,
                         "constants_from": consts.get("capture"),
                         "peak_source": "isof_int32_peak micro-kernel (independent VIADDMNMX chains), measured in this run",
                         "hbm": {"achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
                                 "algorithmic_bytes_per_launch": alg_bytes,
                                 "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650"},
                         "padded_cell_frac": prof["dp_cells_padded"] / max(prof["dp_cells"], 1),
                         "packed_path_cell_frac": packed_frac, "query_profile_cell_frac": profile_frac,
                         "note": "integer DP, ALU-pipe bound (tensor cores do not apply; HBM is the secondary resource): achieved = cells "
                                 "per launch x ALU lane-ops per cell (ncu, constants_from) / launch duration.  `frac` uses per-launch "
                                 "durations of one serialised step (isof_set_overlap(0), CUDA events on the launch stream, tails "
                                 "exposed: lower bound); `frac_step` charges the whole timed, pipelined step (all kernels) to the DP "
                                 "cells: upper bound"},
            "minimizer_kernel": {"mbases_per_s": wl["n_bases"] / max(prof["minimizer_ms"], 1e-9) / 1e3,
                                 "minimizers_per_s": n_min / max(prof["minimizer_ms"], 1e-9) * 1e3,
                                 "int32_ops_per_base": 180, "int32_frac": (wl["n_bases"] * 180.0 /
                                                                           max(prof["minimizer_ms"] * 1e-3, 1e-12)) / int32_peak,
                                 "note": "all four launches of isof_minimizers_batch_dev (hash+select, count scan, compaction); "
                                         "launch-latency bound at one 1000-read batch; super_batch = the same call over "
                                         "many batches' reads at once (C3 regime), CUDA events, int32_frac at 180 ops/base",
                                 "super_batch": dict(min_super, int32_frac=min_super["mbases_per_s"] * 1e6 * 180.0 / int32_peak) if min_super and int32_peak else min_super},
            "kernel_ms_per_step": {"dp": prof["dp_ms"], "traceback_features": prof["traceback_ms"],
                                   "minimizers": prof["minimizer_ms"], "plan_encode_spans": prof["other_ms"],
                                   "note": "serialised kernel step; in the timed (pipelined) steps traceback overlaps DP"},
            "clocks": clk,
        }
        line.update(extras)
        if not args.no_cpu_baseline and world == 1:
            import oracle
            threads = os.cpu_count() or 1
            sample = args.cpu_sample_pairs or pick_cpu_sample(wl, threads, target_s=12.0)
            t_min, t_al, c, t_span = cpu_step(wl, sample, threads)
            span = cpu_span_stage_seconds(wl)
            line["cpu_baseline"] = {"value": sample / (t_min + t_al + t_span), "unit": "alignments/s", "cores": oracle.num_threads(),
                                    "kind": "port", "aligner_gcups": c / t_al / 1e9,
                                    "aligner_gcups_per_core": c / t_al / 1e9 / max(oracle.num_threads(), 1),
                                    "minimizer_mbases_per_s": wl["n_bases"] / t_min / 1e6,
                                    "sample": cpu_sample_text(sample, wl, span)}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
