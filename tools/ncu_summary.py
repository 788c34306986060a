#!/usr/bin/env python3
"""Summarise an ncu report exported with `--page raw --csv` and `--page source --csv`:
headline metrics, opcode mix, and executed instructions / active lanes / stall samples per SASS chunk."""
import collections
import csv
import re
import sys

raw, src = sys.argv[1], sys.argv[2]
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 250
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__inst_executed.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct', 'smsp__average_warps_issue_stalled', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__inst_executed_pipe_alu.sum.pct',
        'sm__inst_executed_pipe_fma.sum.pct', 'sm__inst_executed_pipe_fp64.sum.pct', 'launch__occupancy_limit']
for h, u, v in zip(hdr, units, vals):
    if any(w in h for w in want) and 'per_second' not in h and 'pct_of_peak_sustained_elapsed' not in h:
        print("%-90s %-14s %s" % (h, u, v))
rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
ix = {h: i for i, h in enumerate(rows[hi])}
data = rows[hi + 1:]
f = lambda r, k: float(r[ix[k]] or 0)
tot = sum(f(r, "Instructions Executed") for r in data)
thr = sum(f(r, "Thread Instructions Executed") for r in data)
print("\nSASS rows %d, warp instructions %.4g, thread instructions %.4g, lanes/instr %.2f" % (len(data), tot, thr, thr / tot))
op = collections.Counter()
for r in data:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
    op[m.group(2).split('.')[0] if m else '?'] += f(r, "Instructions Executed")
print("opcode mix: " + ", ".join("%s %.1f%%" % (o, 100 * c / tot) for o, c in op.most_common(16)))
print("\nchunk        warp-inst   share  lanes  samples  no_inst  wait  short_sb  long_sb  math  branch")
for k in range(0, len(data), chunk):
    seg = data[k:k + chunk]
    ie = sum(f(r, "Instructions Executed") for r in seg)
    te = sum(f(r, "Thread Instructions Executed") for r in seg)
    g = lambda name: int(sum(f(r, name) for r in seg))
    print("%5d-%5d  %10.4g  %5.1f%%  %5.1f  %7d  %7d  %5d  %7d  %7d  %5d  %5d" % (
        k, k + chunk, ie, 100 * ie / tot, te / max(ie, 1), g("# Samples"), g("stall_no_inst"), g("stall_wait"),
        g("stall_short_sb"), g("stall_long_sb"), g("stall_math"), g("stall_branch_resolving")))
