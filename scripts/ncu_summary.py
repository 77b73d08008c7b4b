#!/usr/bin/env python
"""Summarise ncu outputs for profiles/: per-kernel shares from a `--metrics gpu__time_duration.sum`
launch list (CSV log) and selected raw metrics from a `--set full` .ncu-rep."""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_lsu.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum", "lts__t_bytes.sum", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0][:70]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e6 if u.startswith("n") else v / 1e3 if u.startswith("u") else v * 1e3 if u in ("s", "second") else v
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print("# launch list %s: per-kernel device time (cold-cache, serialised -- compare SHARES)" % path)
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-72s n=%5d %12.3f ms %6.2f%%" % (k, v[0], v[1], 100 * v[1] / tot))
    print("total %.3f ms over %d launches" % (tot, sum(v[0] for v in agg.values())))


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print("# raw metrics of %s (%d captured launches)" % (path, len(rows) - 2))
    ki = hdr.index("Kernel Name") if "Kernel Name" in hdr else None
    if ki is not None:
        print("kernels:", sorted(set(r[ki].split("(")[0] for r in rows[2:])))
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print("%-90s %-14s %s" % (k, units[i], " | ".join(r[i] for r in rows[2:])))


if __name__ == "__main__":
    for a in sys.argv[1:]:
        (raw if a.endswith(".ncu-rep") else launches)(a)
        print()
