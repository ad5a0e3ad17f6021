#!/usr/bin/env python3
"""Summarises ncu output for profiles/: `launches <csv>` aggregates a gpu__time_duration launch list by kernel;
`metrics <ncu-rep>` prints the roofline-relevant raw metrics per profiled launch."""
import collections
import csv
import re
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__sass_inst_executed_op_global_ld.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg, tot = collections.defaultdict(lambda: [0, 0.0]), 0.0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[row["Metric Unit"]]
        k = re.sub(r"\(.*", "", row["Kernel Name"])[:90]
        agg[k][0] += 1
        agg[k][1] += v
        tot += v
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot:.3f} ms of kernel time (ncu-serialised, cold cache)")
    print(f"# {'ms':>10s} {'share':>6s} {'count':>6s}  kernel")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{t:12.3f} {100 * t / tot:5.1f}% {n:6d}  {k}")


def metrics(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    hdr, units, rows = r[0], r[1], r[2:]
    ki = hdr.index("Kernel Name")
    print(f"# {path}")
    for n, row in enumerate(rows):
        print(f"## launch {n}: {row[ki][:100]}")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"{w:85s} {row[i]:>18s} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "metrics": metrics}[sys.argv[1]](sys.argv[2])
