#!/usr/bin/env python
"""bench.py -- the isONform hot path on B200: interval alignments/s (+ GCUPS, reads/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A *step* is one pass of the hot path over one synthetic isONclust batch (BASELINE.json configs[1]:
1000 ONT-like reads of a 10-isoform 3 kb gene at 5 % error, SURVEY.md section 8d "C2"):
  (a) minimizer extraction over the batch (k=20, w=31),
  (c) the minimizer-pair database + span search (xmin=40, xmax=80, delta_len=5), and
  (b) the all-pairs semi-global affine alignment (2,-2,12,1) of the batch's isoform consensuses with
      traceback, CIGAR and the integer merge-decision features -- the `merge_consensuses` loop of
      the reference (IsoformGeneration.py:473-501), where at 5 % error every read is its own group
      and no row merges, i.e. N(N-1)/2 = 499 500 alignments per batch.
`value` times the step with inputs resident in HBM (device pointers, C ABI `_dev` calls); `e2e`
times the same step through the host-buffer C ABI with pinned host memory, H2D of the reads and
pair list and D2H of scores, CIGAR runs, offsets and features inside the timed region.

With N > 1 (torchrun) every rank processes its own cluster (seed = base + rank): clusters are
independent units, there is no collective on the data path ("weak" scaling); NCCL is used only
for the barrier and the max-over-ranks of the step time.

--impl reference times the CPU restatement of the same step (oracle/, "port": the reference's
aligner is the un-vendored, un-installable parasail; its minimizer loop is pure Python) on all
host cores over a bounded sample of the same workload.
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
# ALU-pipe instructions per DP cell in the SASS of sg_dp_kernel: the packed 16-bit
# path issues 12 per packed row (= 2 cells), 10 when the substitution scores come from the shared-memory
# query profile (tasks whose two alignments share the reference); the 32-bit path 10 per row (= 1 cell; 12 with
# wildcards).  Trace-bit gathering runs as IMADs on the FMA pipe and is not counted here.
OPS_PER_CELL_PACKED, OPS_PER_CELL_PROFILE, OPS_PER_CELL_32 = 6.0, 5.0, 10.0
# DRAM bytes per DP cell of sg_dp_kernel measured by ncu (dram__bytes_read.sum + dram__bytes_write.sum of one launch
# over its cell count; profiles/r01_ncu_v10_summary.txt)
NCU_DRAM_BYTES_PER_CELL = 0.5605


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
    ap.add_argument("--trace-budget-gb", type=float, default=None,
                    help="alignment trace scratch in GB (default: the library's 40-55 %% of free HBM); long-read batches "
                         "(C4) need more of it to keep every warp slot busy")
    return ap.parse_args()


def make_workload(args, rank):
    from isonform_b200 import synth
    from isonform_b200._lib import pack_sequences
    from isonform_b200.polya import remove_read_polyA_ends
    base_seed = synth.CONFIGS[args.workload]["seed"] if args.seed is None else args.seed
    n_reads = args.reads
    if n_reads is None and args.workload == "C4":
        n_reads = 1000                      # one 1000-read batch of the 2000-read cluster
    # rank 0 gets the canonical cluster; other ranks get clusters of the same gene family with re-drawn reads
    # (same size statistics, so that "weak scaling" compares like with like)
    reads, _iso, _truth = synth.make_config(args.workload, seed=base_seed, n_reads=n_reads,
                                            read_seed=None if rank == 0 else rank)
    seqs = [remove_read_polyA_ends(s, 12, 1) for _, s in reads]       # main:351
    seqs_sorted = sorted(seqs, key=len)                                # merge_consensuses sorts by length (:471)
    cat, off = pack_sequences(seqs_sorted)
    # all (shorter, longer) = (consensus1, consensus2) pairs, grouped by the longer sequence (the reference of
    # the alignment): consecutive pairs then share their reference, which the DP kernel exploits
    ib, ia = np.tril_indices(len(seqs_sorted), -1)
    cells = int((np.diff(off)[ia].astype(np.int64) * np.diff(off)[ib].astype(np.int64)).sum())
    name = "%s: %d-read batch, %s-isoform %d bp gene, eps=%.2f; all-pairs merge_consensuses alignments" % (
        args.workload, len(seqs), synth.CONFIGS[args.workload]["n_isoforms"], synth.CONFIGS[args.workload]["gene_len"],
        synth.CONFIGS[args.workload]["eps"])
    return dict(seqs=seqs_sorted, cat=cat, off=off, ia=ia.astype(np.int32), ib=ib.astype(np.int32), cells=cells,
                n_reads=len(seqs), n_pairs=len(ia), n_bases=int(off[-1]), name=name)


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
    return t_min, t_al, cells


def pick_cpu_sample(wl, threads, target_s=12.0):
    """Size the pair sample so that one CPU step takes about target_s seconds."""
    unit = 16 * threads                       # the SIMD port aligns 16 pairs per thread at a time
    probe = min(unit, wl["n_pairs"])
    _tm, t_al, _ = cpu_step(wl, probe, threads)
    per_pair = max(t_al / probe, 1e-6)
    want = int(max(probe, min(wl["n_pairs"], target_s / per_pair)))
    return min(wl["n_pairs"], max(unit, (want // unit) * unit))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    wl = make_workload(args, 0)
    threads = os.cpu_count() or 1
    sample = args.cpu_sample_pairs or pick_cpu_sample(wl, threads, target_s=10.0)
    for _ in range(args.warmup):
        cpu_step(wl, max(threads, sample // 8), threads)
    times, cells = [], 0
    for _ in range(args.steps):
        t_min, t_al, c = cpu_step(wl, sample, threads)
        times.append(t_min + t_al)
        cells = c
    t = float(np.sum(times))
    value = sample * args.steps / t
    line = {"impl": "reference", "metric": "interval_alignments_per_s", "value": value, "unit": "alignments/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": wl["name"], "n_reads": wl["n_reads"], "pairs_per_batch": wl["n_pairs"],
                       "k": K_SIZE, "w": W_SIZE, "scores": [MATCH, MISMATCH, GAP_OPEN, GAP_EXT]},
            "gcups": cells * args.steps / t / 1e9,
            "cpu_baseline": {"value": value, "unit": "alignments/s", "cores": oracle.num_threads(), "kind": "port",
                             "sample": "%d of %d pairs per step (evenly spaced) + minimizers of all %d reads; "
                                       "oracle/isof_oracle.c Gotoh+trace+cigar, 16 pairs per thread in int16 SIMD lanes (AVX2 where available), OpenMP"
                                       % (sample, wl["n_pairs"], wl["n_reads"])},
            "e2e": {"value": value, "unit": "alignments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from isonform_b200 import _lib
    import ctypes as C

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

    wl = make_workload(args, rank)
    ctx = _lib.Context(local)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    if args.trace_budget_gb:
        ctx.set_trace_budget(int(args.trace_budget_gb * (1 << 30)))
    P, R = wl["n_pairs"], wl["n_reads"]

    # ---- device-resident inputs / outputs (torch owns the memory, the library gets raw pointers)
    d_cat = torch.from_numpy(wl["cat"]).to(dev)
    d_off = torch.from_numpy(wl["off"]).to(dev)
    d_ia = torch.from_numpy(wl["ia"]).to(dev)
    d_ib = torch.from_numpy(wl["ib"]).to(dev)
    d_score = torch.empty(P, dtype=torch.int32, device=dev)
    d_endq = torch.empty(P, dtype=torch.int32, device=dev)
    d_endr = torch.empty(P, dtype=torch.int32, device=dev)
    d_coff = torch.empty(P + 1, dtype=torch.int64, device=dev)
    d_feat = torch.empty((P, _lib.N_FEATURES), dtype=torch.int32, device=dev)
    d_mpos = torch.empty(wl["n_bases"] + R, dtype=torch.int32, device=dev)
    d_moff = torch.empty(R + 1, dtype=torch.int64, device=dev)
    rec_cap = 8 * (wl["n_bases"] // 5) + 1024                 # ~0.15 minimizers/base, ~6 partners each
    d_rec = [torch.empty(rec_cap, dtype=torch.int32, device=dev) for _ in range(6)]
    d_boff = torch.empty(rec_cap + 1, dtype=torch.int32, device=dev)
    d_bdrop = torch.empty(rec_cap, dtype=torch.uint8, device=dev)
    cig_cap = [max(64 * P, 1 << 20)]
    d_cig = [torch.empty(cig_cap[0], dtype=torch.int32, device=dev)]

    stats = {}

    def step_dev():
        n_min = ctx.minimizers_batch_dev(d_cat.data_ptr(), d_off.data_ptr(), R, wl["n_bases"], K_SIZE, W_SIZE,
                                         d_mpos.data_ptr(), d_mpos.numel(), d_moff.data_ptr())
        n_rec, _n_b = ctx.spans_batch_dev(d_cat.data_ptr(), d_off.data_ptr(), R, d_mpos.data_ptr(), d_moff.data_ptr(), n_min,
                                          K_SIZE, X_MIN, X_MAX, DELTA_LEN, rec_cap, d_rec[0].data_ptr(), d_rec[1].data_ptr(),
                                          d_rec[2].data_ptr(), d_rec[3].data_ptr(), d_rec[4].data_ptr(), d_rec[5].data_ptr(),
                                          d_boff.data_ptr(), d_bdrop.data_ptr())
        stats["n_rec"] = n_rec
        while True:
            try:
                runs = ctx.sg_align_pairs_dev(d_cat.data_ptr(), d_off.data_ptr(), R, wl["n_bases"], d_ia.data_ptr(),
                                              d_ib.data_ptr(), P, MATCH, MISMATCH, GAP_OPEN, GAP_EXT, d_score.data_ptr(),
                                              d_endq.data_ptr(), d_endr.data_ptr(), d_cig[0].data_ptr(), cig_cap[0],
                                              d_coff.data_ptr(), d_feat.data_ptr(), DELTA_LEN)
                return n_min, runs
            except _lib.IsofError as e:
                if e.code != _lib.ERR_CAPACITY:
                    raise
                need = int(d_coff[P].item())
                cig_cap[0] = need + need // 16
                d_cig[0] = torch.empty(cig_cap[0], dtype=torch.int32, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also sizes the CIGAR buffer)
    for _ in range(max(args.warmup, 1) if args.warmup else 0):
        n_min, n_runs = step_dev()
    if args.warmup == 0:
        n_min, n_runs = step_dev()
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
        n_min, n_runs = step_dev()
    e1.record(stream)
    barrier()
    ms_dev = e0.elapsed_time(e1)
    launches_timed = ctx.profile()["launches"]

    # ---- roofline step: the same step once more with the chunks serialised on one stream and every launch
    # bracketed by CUDA events on that stream, so that per-kernel durations are exact (in the pipelined
    # configuration consecutive DP launches overlap by design and their individual durations are not
    # meaningful).  Not part of `value`.
    ctx.set_overlap(False)
    ctx.profile_reset()
    ctx.profile_enable(True)
    barrier()
    step_dev()
    barrier()
    prof = ctx.profile()
    ctx.profile_enable(False)
    ctx.set_overlap(True)
    PROF_STEPS = 1

    # ---- timed: end to end through the host-buffer C ABI, pinned host memory
    h_cat = torch.from_numpy(wl["cat"]).pin_memory()
    h_off = torch.from_numpy(wl["off"]).pin_memory()
    h_ia = torch.from_numpy(wl["ia"]).pin_memory()
    h_ib = torch.from_numpy(wl["ib"]).pin_memory()
    h_score = torch.empty(P, dtype=torch.int32).pin_memory()
    h_feat = torch.empty((P, _lib.N_FEATURES), dtype=torch.int32).pin_memory()
    h_coff = torch.empty(P + 1, dtype=torch.int64).pin_memory()
    h_cig = torch.empty(int(n_runs) + 1024, dtype=torch.int32).pin_memory()
    h_mpos = torch.empty(int(n_min) + 1024, dtype=torch.int32).pin_memory()
    h_moff = torch.empty(R + 1, dtype=torch.int64).pin_memory()
    L = ctx._L
    n_out = C.c_int64(0)
    h_rec = [torch.empty(rec_cap, dtype=torch.int32).pin_memory() for _ in range(6)]
    h_boff = torch.empty(rec_cap + 1, dtype=torch.int32).pin_memory()
    h_bdrop = torch.empty(rec_cap, dtype=torch.uint8).pin_memory()
    n_rec_h, n_b_h = C.c_int64(0), C.c_int64(0)

    def step_e2e():
        ctx._check(L.isof_minimizers_batch(ctx._h, h_cat.data_ptr(), h_off.data_ptr(), R, K_SIZE, W_SIZE, h_mpos.data_ptr(),
                                           h_mpos.numel(), h_moff.data_ptr(), C.byref(n_out)))
        ctx._check(L.isof_spans_batch(ctx._h, h_cat.data_ptr(), h_off.data_ptr(), R, h_mpos.data_ptr(), h_moff.data_ptr(), K_SIZE,
                                      X_MIN, X_MAX, DELTA_LEN, rec_cap, h_rec[0].data_ptr(), h_rec[1].data_ptr(),
                                      h_rec[2].data_ptr(), h_rec[3].data_ptr(), h_rec[4].data_ptr(), h_rec[5].data_ptr(),
                                      h_boff.data_ptr(), h_bdrop.data_ptr(), C.byref(n_rec_h), C.byref(n_b_h)))
        ctx._check(L.isof_sg_features_pairs(ctx._h, h_cat.data_ptr(), h_off.data_ptr(), R, h_ia.data_ptr(), h_ib.data_ptr(), P,
                                            MATCH, MISMATCH, GAP_OPEN, GAP_EXT, DELTA_LEN, h_score.data_ptr(),
                                            h_feat.data_ptr(), h_cig.data_ptr(), h_cig.numel(), h_coff.data_ptr()))
        # the step's result as the caller consumes it: merge decisions from the features
        from isonform_b200.hotpath import _merge_from_features
        return int(_merge_from_features(h_feat.numpy(), 0.15, DELTA_LEN, 30, 50).sum())

    step_e2e()
    ctx.profile_reset()
    barrier()
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        merges = step_e2e()
    e1.record(stream)
    barrier()
    wall_e2e = time.perf_counter() - t0
    prof_e2e = ctx.profile()
    clk = clocks.stop() if rank == 0 else None

    # ---- max over ranks
    t = torch.tensor([ms_dev, wall_e2e * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev_max, ms_e2e_max = float(t[0]), float(t[1])
    tot = torch.tensor([float(P), float(wl["cells"]), float(R)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    pairs_all, cells_all, reads_all = float(tot[0]), float(tot[1]), float(tot[2])

    # ---- kernel (a) at super-batch scale (the C3 regime: many clusters' reads in one call); rank 0, untimed extra
    min_super = None
    if rank == 0:
        reps = max(1, min(32, int(2.0e8 // max(wl["n_bases"], 1))))
        offs = [wl["off"][:-1] + i * wl["n_bases"] for i in range(reps)] + [np.array([reps * wl["n_bases"]], np.int64)]
        d_cat_s = d_cat.repeat(reps)
        d_off_s = torch.from_numpy(np.concatenate(offs).astype(np.int64)).to(dev)
        cap_s = reps * wl["n_bases"] // 4 + 1024
        d_pos_s = torch.empty(cap_s, dtype=torch.int32, device=dev)
        d_poff_s = torch.empty(reps * R + 1, dtype=torch.int64, device=dev)
        ctx.profile_enable(False)
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
        dp_ms_launch = prof["dp_ms"] / max(prof["dp_launches"], 1)
        # algorithmic bytes of the DP kernel: 0.5 B/cell trace write + the two sequences, per launch
        alg_bytes = (0.5 * prof["dp_cells"] + 2.0 * wl["n_bases"] * PROF_STEPS) / max(prof["dp_launches"], 1)
        achieved_gbs = alg_bytes / (dp_ms_launch * 1e-3) / 1e9
        cups = prof["dp_cells"] / (prof["dp_ms"] * 1e-3)
        packed_frac = prof["dp_cells_packed"] / max(prof["dp_cells"], 1)
        profile_frac = prof["dp_cells_profile"] / max(prof["dp_cells"], 1)
        OPS_PER_CELL = (profile_frac * OPS_PER_CELL_PROFILE + (packed_frac - profile_frac) * OPS_PER_CELL_PACKED +
                        (1.0 - packed_frac) * OPS_PER_CELL_32)
        line = {
            "metric": "interval_alignments_per_s", "value": value, "unit": "alignments/s", "n_gpus": world,
            "steps": K, "warmup": args.warmup, "ms_per_step": ms_dev_max / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": wl["name"], "n_reads": R, "pairs_per_batch": P, "mean_read_len": wl["n_bases"] / R,
                       "k": K_SIZE, "w": W_SIZE, "xmin": X_MIN, "xmax": X_MAX, "scores": [MATCH, MISMATCH, GAP_OPEN, GAP_EXT],
                       "delta_len": DELTA_LEN,
                       "parallelism": "1 cluster per GPU, no data-path collective",
                       "l2": "per-step trace working set %.0f GB >> 126 MB L2; no explicit flush" % (
                           prof["trace_bytes"] / PROF_STEPS / 1e9)},
            "gcups": cells_all * K / (ms_dev_max * 1e-3) / 1e9,
            "reads_per_s": reads_all * K / (ms_dev_max * 1e-3),
            "minimizers_per_step": int(n_min), "minimizer_pairs_per_step": int(stats.get("n_rec", 0)), "cigar_runs_per_step": int(n_runs), "merges_per_step": int(merges),
            "e2e": {"value": e2e_value, "unit": "alignments/s",
                    "h2d_bytes_per_step": int(prof_e2e["h2d_bytes"] // K), "d2h_bytes_per_step": int(prof_e2e["d2h_bytes"] // K),
                    "ms_per_step": ms_e2e_max / K, "gcups": cells_all * K / (ms_e2e_max * 1e-3) / 1e9},
            "gpu_launches": int(launches_timed),
            "roofline": {"kernel": "sg_dp_kernel", "bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak,
                         "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
                         "traffic": NCU_DRAM_BYTES_PER_CELL * prof["dp_cells"] / max(prof["dp_launches"], 1),
                         "traffic_note": "bytes per launch, scaled by cells from the ncu --set full capture of a 4950-pair launch "
                                         "of the same reads (profiles/r01_ncu_v10_summary.txt: dram write 20.379 GB + read 0.664 GB "
                                         "over 37.54 G cells = 0.5605 B/cell = 1.12x the 0.5 B/cell of trace: stripe/row padding, "
                                         "no re-reads); a bench-sized launch writes ~70 GB of trace, too large for ncu's "
                                         "save/restore replay",
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650",
                         "launch_ms": dp_ms_launch, "launches": int(prof["dp_launches"]),
                         "note": "integer DP: the kernel is INT32-ALU bound, see int_alu; algorithmic bytes = "
                                 "0.5 B/cell trace + sequences; kernel durations from a serialised roofline step "
                                 "(isof_set_overlap(0)) of the same workload, CUDA events on the launch stream"},
            "int_alu": {"cell_updates_per_s": cups, "gcups_kernel": cups / 1e9, "ops_per_cell": OPS_PER_CELL,
                        "achieved_lane_ops_per_s": cups * OPS_PER_CELL, "peak_lane_ops_per_s": int32_peak,
                        "frac": cups * OPS_PER_CELL / int32_peak if int32_peak else None,
                        "padded_cell_frac": prof["dp_cells_padded"] / max(prof["dp_cells"], 1), "packed_path_cell_frac": packed_frac, "query_profile_cell_frac": profile_frac,
                        "peak_source": "isof_int32_peak micro-kernel (independent VIADDMNMX chains), measured in this run"},
            "minimizer_kernel": {"mbases_per_s": wl["n_bases"] * PROF_STEPS / max(prof["minimizer_ms"], 1e-9) / 1e3,
                                 "minimizers_per_s": n_min * PROF_STEPS / max(prof["minimizer_ms"], 1e-9) * 1e3,
                                 "int32_ops_per_base": 180, "int32_frac": (wl["n_bases"] * PROF_STEPS * 180.0 /
                                                                           max(prof["minimizer_ms"] * 1e-3, 1e-12)) / int32_peak,
                                 "note": "all four launches of isof_minimizers_batch_dev (hash+select, count scan, compaction); "
                                         "launch-latency bound at one 1000-read batch; super_batch = the same call over "
                                         "many batches' reads at once (C3 regime), CUDA events, int32_frac at 180 ops/base",
                                 "super_batch": dict(min_super, int32_frac=min_super["mbases_per_s"] * 1e6 * 180.0 / int32_peak) if min_super and int32_peak else min_super},
            "kernel_ms_per_step": {"dp": prof["dp_ms"] / PROF_STEPS, "traceback_cigar_features": prof["traceback_ms"] / PROF_STEPS,
                                   "minimizers": prof["minimizer_ms"] / PROF_STEPS, "plan_encode_spans": prof["other_ms"] / PROF_STEPS,
                                   "note": "serialised roofline step; in the timed (pipelined) steps traceback overlaps DP"},
            "clocks": clk,
        }
        if not args.no_cpu_baseline and world == 1:
            import oracle
            threads = os.cpu_count() or 1
            sample = args.cpu_sample_pairs or pick_cpu_sample(wl, threads, target_s=12.0)
            t_min, t_al, c = cpu_step(wl, sample, threads)
            line["cpu_baseline"] = {"value": sample / (t_min + t_al), "unit": "alignments/s", "cores": oracle.num_threads(),
                                    "kind": "port", "gcups": c / t_al / 1e9, "minimizer_mbases_per_s": wl["n_bases"] / t_min / 1e6,
                                    "sample": "%d of %d pairs (evenly spaced) + minimizers of all %d reads; "
                                              "oracle/isof_oracle.c Gotoh+trace+cigar, 16 pairs per thread in int16 SIMD lanes (AVX2 where available), OpenMP"
                                              % (sample, P, R)}
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
