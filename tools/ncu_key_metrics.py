#!/usr/bin/env python3
"""Print the key metrics of an .ncu-rep (one kernel) as a short table.  Usage: ncu_key_metrics.py rep"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg",
        "sm__cycles_active.avg", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__inst_executed_pipe_fmaheavy.sum", "sm__inst_executed_pipe_fmaheavy.sum"]
# every per-pipe instruction / utilisation metric of the capture (FP64, IMAD / fmaheavy / fmalite, ALU, ...): the binding
# roofline of these kernels is an issue pipe, so the whole breakdown is wanted (north_star: IMAD-pipe utilisation)
PIPE_RE = ("sm__inst_executed_pipe_", "smsp__inst_executed_pipe_", "sm__pipe_", "smsp__pipe_", "smsp__issue_active", "sm__inst_issued",
           "smsp__inst_issued")
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for row in rows[2:]:
    print("kernel:", row[hdr.index("Kernel Name")][:80])
    for h, u, v in zip(hdr, units, row):
        pipe = h.startswith(PIPE_RE) and (h.endswith(".sum") or ".avg.pct_of_peak_sustained_active" in h) and v not in ("0", "")
        if h in WANT or pipe:
            print("  %-84s %-10s %s" % (h, u, v))
