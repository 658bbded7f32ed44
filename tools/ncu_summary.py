#!/usr/bin/env python
"""Summaries for profiles/: (1) launch list csv -> markdown table of kernel shares, (2) .ncu-rep -> selected raw metrics.

  python tools/ncu_summary.py launches gpurun_out/r01_v3_launches.csv
  python tools/ncu_summary.py raw gpurun_out/prof.ncu-rep
"""
import collections
import csv
import io
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
           "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "smsp__average_warp_latency_per_inst_issued.ratio",
           "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores", "sass__inst_executed_local_loads",
           "sass__inst_executed_local_stores"]


def launches(path):
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    tot = collections.Counter()
    cnt = collections.Counter()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"].split("(")[0][:100]
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
        tot[name] += v
        cnt[name] += 1
    s = sum(tot.values())
    print("| kernel | launches | total ns | share |\n|---|---|---|---|")
    for k, v in tot.most_common():
        print("| `%s` | %d | %.0f | %.2f %% |" % (k, cnt[k], v, 100 * v / s))


def raw(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(txt)))
    h, u, v = r[0], r[1], r[2]
    print("| metric | unit | value |\n|---|---|---|")
    print("| Kernel Name |  | %s |" % v[h.index("Kernel Name")])
    for m in METRICS:
        if m in h:
            print("| %s | %s | %s |" % (m, u[h.index(m)], v[h.index(m)]))


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
