#!/usr/bin/env python
"""Kernel (a) at scale: minimizers of a super-batch (many clusters' reads in one call, the C3 regime),
device-resident, timed with CUDA events.  Prints one JSON line with Mbases/s and the INT32-ALU fraction
(180 INT32 ops per base) against the peak measured by isof_int32_peak in the same run."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isonform_b200 import _lib, hotpath  # noqa: E402


def main():
    n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 64000
    rng = np.random.default_rng(3)
    lens = rng.integers(1800, 3600, n_reads)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    cat = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, int(off[-1]))]
    ctx = _lib.Context(0)
    dev = torch.device("cuda:0")
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    d_cat = torch.from_numpy(cat).to(dev)
    d_off = torch.from_numpy(off).to(dev)
    cap = int(off[-1] // 4 + 1024)
    d_pos = torch.empty(cap, dtype=torch.int32, device=dev)
    d_poff = torch.empty(n_reads + 1, dtype=torch.int64, device=dev)
    peak = ctx.int32_peak()
    ctx.profile_enable(True)
    res = {}
    for (k, w) in [(20, 31), (9, 20)]:
        for _ in range(3):
            n_out = ctx.minimizers_batch_dev(d_cat.data_ptr(), d_off.data_ptr(), n_reads, int(off[-1]), k, w, d_pos.data_ptr(), cap,
                                             d_poff.data_ptr())
        ctx.profile_reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        steps = 5
        for _ in range(steps):
            ctx.minimizers_batch_dev(d_cat.data_ptr(), d_off.data_ptr(), n_reads, int(off[-1]), k, w, d_pos.data_ptr(), cap, d_poff.data_ptr())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        prof = ctx.profile()
        kms = prof["minimizer_ms"] / steps
        # check a sample against the per-read API result shape (full parity lives in tests/)
        res["k%d_w%d" % (k, w)] = {"n_reads": n_reads, "mbases": int(off[-1]) / 1e6, "minimizers": n_out, "call_ms": ms, "kernel_ms": kms,
                                  "mbases_per_s_call": int(off[-1]) / ms / 1e3, "mbases_per_s_kernels": int(off[-1]) / kms / 1e3,
                                  "int32_ops_per_base": 180 if k == 20 else None,
                                  "int32_frac": (int(off[-1]) * 180 / (kms * 1e-3) / peak) if k == 20 else None}
    res["int32_peak_lane_ops_per_s"] = peak
    print(json.dumps(res))


if __name__ == "__main__":
    main()
