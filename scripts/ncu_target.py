#!/usr/bin/env python
"""Program profiled by `ncu --set full -k regex:sg_dp_kernel`: ONE alignment call of a C2-like all-pairs job small
enough for ncu's save/restore replay (default: the first 100 reads of the C2 batch = 4950 pairs).  Prints (and writes to
--out) the exact cell count of the launch so that per-cell constants can be derived from the capture
(scripts/ncu_dp_constants.py).  Not a benchmark: nothing timed here is reported."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isonform_b200 import _lib, hotpath, synth  # noqa: E402
from isonform_b200.polya import remove_read_polyA_ends  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=100)
    ap.add_argument("--workload", default="C2")
    ap.add_argument("--out", default="gpurun_out/ncu_target.json")
    a = ap.parse_args()
    reads, _i, _t = synth.make_config(a.workload, n_reads=a.reads)
    seqs = sorted((remove_read_polyA_ends(s, 12, 1) for _, s in reads), key=len)
    cat, off = hotpath.pack_sequences(seqs)
    ib, ia = np.tril_indices(len(seqs), -1)
    lens = np.diff(off)
    cells = int((lens[ia].astype(np.int64) * lens[ib]).sum())
    ctx = _lib.Context(0)
    ctx.set_overlap(False)                    # one chunk on one stream: the launch the capture sees is the whole job
    ctx.profile_enable(True)
    ctx.profile_reset()
    ctx.sg_align_pairs(cat, off, ia.astype(np.int32), ib.astype(np.int32), 2, -2, 12, 1, want_cigar=False, features_delta_len=5)
    pr = ctx.profile()
    info = {"workload": "%s first %d reads, all pairs" % (a.workload, a.reads), "pairs": int(len(ia)), "cells": cells,
            "cells_padded": int(pr["dp_cells_padded"]), "cells_packed": int(pr["dp_cells_packed"]),
            "cells_profile": int(pr["dp_cells_profile"]), "dp_launches": int(pr["dp_launches"]), "dp_ms_unprofiled_or_replayed": pr["dp_ms"]}
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump(info, open(a.out, "w"), indent=1)
    print(json.dumps(info))


if __name__ == "__main__":
    main()
