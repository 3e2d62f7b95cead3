#!/usr/bin/env python
"""Condenses an `ncu --page raw --csv` export into the JSON bench.py reads (profiles/<tag>_traffic.json) and a short
CSV of the metrics DESIGN.md quotes.
    ncu -i capture.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_summary.py raw.csv --reads 100000 --tag r1_lanes_v8 [--kernel lane_fill]"""
import argparse
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("raw")
    ap.add_argument("--reads", type=int, required=True)
    ap.add_argument("--tag", required=True)
    ap.add_argument("--kernel", default="lane_fill")
    ap.add_argument("--command", default="bench.py --steps 1 --warmup 1")
    args = ap.parse_args()
    rows = list(csv.reader(open(args.raw)))
    hdr, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(hdr)}
    launches, table = [], []
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        if args.kernel not in name:
            continue

        def val(metric):
            if metric not in col or r[col[metric]] == "":
                return None
            v = float(r[col[metric]].replace(",", ""))
            return v * SCALE.get(units[col[metric]], 1.0)

        short = name.split("(")[0].replace("void ", "").replace("pg2::", "")
        issue = val("smsp__issue_active.avg.pct_of_peak_sustained_active")
        fp64 = val("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")
        launches.append({
            "kernel": short, "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
            "time_ms_under_ncu": val("gpu__time_duration.sum"), "warp_instructions": val("smsp__inst_executed.sum"),
            # the kernel's bound is the issue port (one warp-instruction per cycle and sub-partition): the FP64 pipe
            # (2 cycles per instruction) and the ALU pipe run in its shadow (pg2_measure_dispatch_mix)
            "issue_active_pct": issue, "fp64_pipe_active_pct": fp64,
            "alu_pipe_active_pct": val("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        })
        table.append([short] + [r[col[m]] if m in col else "" for m in KEEP])
    out = {"source": "ncu --set full --clock-control none (%s, %d reads); tools/ncu_summary.py" % (args.command, args.reads),
           "reads_per_gpu": args.reads, "launches": launches}
    with open(os.path.join(ROOT, "profiles", args.tag + "_traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    with open(os.path.join(ROOT, "profiles", args.tag + "_ncu_summary.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + KEEP)
        w.writerow([""] + [units[col[m]] if m in col else "" for m in KEEP])
        w.writerows(table)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
