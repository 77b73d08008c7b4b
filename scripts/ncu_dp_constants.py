#!/usr/bin/env python
"""Derive the per-cell constants bench.py uses for the roofline of sg_dp_kernel from one `ncu --set full` capture.

    ncu -i gpurun_out/prof_dp.ncu-rep --page raw --csv > profiles/rNN_ncu_dp_raw.csv
    python scripts/ncu_dp_constants.py profiles/rNN_ncu_dp_raw.csv gpurun_out/ncu_target.json [--out profiles/ncu_dp_constants.json]

alu_lane_ops_per_cell = ALU-pipe lane-operations of the launch / its cells, with the lane-operations taken from the
pipe's activity counter: sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed / 100 x sm__cycles_elapsed.avg x SMs x
64 lanes per SM and cycle (4 sub-partitions x 16 lanes; the same unit as the isof_int32_peak micro-kernel, whose result is
SMs x 64 x clock).  `--set full` does not collect the absolute sm__inst_executed_pipe_alu.sum, the activity counter is exact.
dram_bytes_per_cell = (dram__bytes_read.sum + dram__bytes_write.sum) / cells."""
import csv
import json
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed_pipe_alu.sum",
        "sm__inst_executed_pipe_alu.sum", "smsp__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fma.sum",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.avg", "smsp__cycles_active.avg", "device__attribute_multiprocessor_count",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.avg.per_second"]
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "byte": 1.0, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0,
        "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0}


def main():
    raw, target = sys.argv[1], sys.argv[2]
    out = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else "profiles/ncu_dp_constants.json"
    rows = list(csv.reader(l for l in open(raw) if not l.startswith("==")))
    head, units = rows[0], rows[1]
    data = [r for r in rows[2:] if any("sg_dp_kernel" in c for c in r)]
    if not data:
        raise SystemExit("no sg_dp_kernel row in %s" % raw)
    r = data[0]
    m = {}
    for name in WANT:
        if name in head:
            i = head.index(name)
            try:
                m[name] = float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)
            except ValueError:
                pass
    t = json.load(open(target))
    cells = float(t["cells"])
    sms = m.get("device__attribute_multiprocessor_count", 148.0)
    lane_cycles = m["sm__cycles_elapsed.avg"] * sms * 64.0
    alu_ops = m["sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed"] / 100.0 * lane_cycles
    fma_ops = m.get("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0) / 100.0 * lane_cycles
    c = {"capture": raw, "kernel": "sg_dp_kernel", "launch": t, "metrics": m,
         "alu_lane_ops_per_cell": alu_ops / cells,
         "alu_lane_ops_per_updated_cell": alu_ops / float(t.get("cells_padded", cells)),
         "fma_lane_ops_per_cell": fma_ops / cells,
         "alu_pipe_pct_active": m.get("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
         "alu_pipe_pct_elapsed": m.get("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed"),
         "dram_bytes_per_cell": (m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]) / cells,
         "dram_bytes": m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"],
         "gcups_under_ncu": cells / m["gpu__time_duration.sum"] / 1e9}
    json.dump(c, open(out, "w"), indent=1)
    print(json.dumps({k: v for k, v in c.items() if k not in ("metrics", "launch")}, indent=1))


if __name__ == "__main__":
    main()
