#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of numbers the roofline discussion needs.
usage: python tools/ncu_summary.py file.ncu-rep [extra_metric_substring ...]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
extra = sys.argv[2:]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__average_t_sectors_per_request_pipe_lsu_mem_global_op_ld.ratio",
        "l1tex__average_t_sectors_per_request_pipe_lsu_mem_global_op_st.ratio",
        "l1tex__t_sector_hit_rate.pct", "sm__cycles_active.avg", "smsp__cycles_active.avg"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    print("==== %s  grid %s block %s" % (d["Kernel Name"][:90], d["Grid Size"], d["Block Size"]))
    for k in KEYS:
        if k in d:
            print("  %-75s %s %s" % (k, d[k], u[k]))
    stalls = [(float(v.replace(",", "")), k) for k, v in d.items()
              if "average_warps_issue_stalled" in k and k.endswith("per_issue_active.ratio") and v not in ("", "n/a")]
    for v, k in sorted(stalls, reverse=True)[:7]:
        print("  stall %-69s %.2f" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
    for e in extra:
        for k, v in d.items():
            if e in k:
                print("  + %-73s %s %s" % (k, v, u[k]))
