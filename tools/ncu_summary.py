#!/usr/bin/env python
"""Text summary of one kernel of an ncu report for profiles/: key launch / throughput / stall metrics (the raw page) plus
the source-line hot spots (tools/ncu_hotspots.py).   python tools/ncu_summary.py <report.ncu-rep> <cubin> <kernel substring>"""
import csv
import io
import subprocess
import sys

rep, cubin, kernel = sys.argv[1:4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
head, units = rows[0], rows[1]
vals = next(r for r in rows[2:] if kernel in ",".join(r))
want = ("Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__shared_mem_per_block_dynamic",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum")
for i, k in enumerate(head):
    if k in want or (k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio")):
        print("%-86s %-16s %s" % (k, units[i], vals[i]))
print()
sys.stdout.flush()
subprocess.run([sys.executable, __file__.replace("ncu_summary.py", "ncu_hotspots.py"), rep, cubin, kernel, "30"])
