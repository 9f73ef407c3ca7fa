#!/usr/bin/env python
"""Summarise an ncu report (read here, without a GPU) into the handful of numbers DESIGN.md / profiles/ quote.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep            # per-kernel table of key metrics
    python tools/ncu_summary.py --launches gpurun_out/launches.csv # per-kernel share of device time from the launch list
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
        "smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio",
        "smsp__average_warp_latency_issue_stalled_wait.ratio", "smsp__average_warp_latency_issue_stalled_not_selected.ratio",
        "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio", "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio"]


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    return hdr, rows[2:] if len(rows) > 2 and not rows[1][0].isdigit() else rows[1:], (rows[1] if len(rows) > 1 else None)


def summarize(rep):
    hdr, rows, units = raw_rows(rep)
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows:
        name = r[col.get("Kernel Name", 4)]
        print("== %s  (id %s)" % (name[:100], r[0]))
        for k in KEYS:
            if k in col:
                u = units[col[k]] if units else ""
                print("   %-78s %s %s" % (k, r[col[k]], u))


def launches(path):
    tot = defaultdict(float)
    cnt = defaultdict(int)
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(io.StringIO("".join(lines)))
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1e-3)
        k = r["Kernel Name"].split("(")[0]
        tot[k] += v
        cnt[k] += 1
    s = sum(tot.values())
    print("%-60s %8s %12s %10s %7s" % ("kernel", "launches", "total us", "avg us", "share"))
    for k in sorted(tot, key=lambda x: -tot[x]):
        print("%-60s %8d %12.1f %10.2f %6.1f%%" % (k[:60], cnt[k], tot[k], tot[k] / cnt[k], 100 * tot[k] / s))


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2])
    else:
        summarize(sys.argv[1])
