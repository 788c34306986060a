#!/usr/bin/env python
"""Per-source-line hot spots of one kernel from an ncu report (read HERE, no GPU needed):
   python tools/ncu_hotspots.py <report.ncu-rep> <cubin built with -lineinfo> <kernel substring> [top N]
Joins `ncu --page source --print-source sass` (instructions executed, stall samples per SASS instruction) with the line
table of `nvdisasm -g` (the report's SASS carries no file names when the source tree is not where it was built).
Prints: share of warp instructions and of stall samples per source line (innermost inlined location), the average
number of active threads, and the dominant stall reasons."""
import collections
import csv
import io
import re
import subprocess
import sys


def line_table(cubin, kernel):
    out = subprocess.run(["nvdisasm", "-g", "-c", cubin], stdout=subprocess.PIPE, text=True).stdout
    table, cur, on = {}, None, False
    for ln in out.splitlines():
        if ln.startswith("//---") and ".text." in ln:
            on = kernel in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
        if m:
            table[int(m.group(1), 16)] = cur
    return table


def main():
    rep, cubin, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    table = line_table(cubin, kernel)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE, text=True).stdout
    lines = raw.splitlines()
    # the first line names the kernel; the header follows
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
    head = rows[0]
    ci = {k: head.index(k) for k in ("Address", "Source", "# Samples", "Instructions Executed", "Thread Instructions Executed")}
    stall_cols = [(k, i) for i, k in enumerate(head) if k.startswith("stall_") and "Not Issued" not in k]
    base = None
    per = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
    tot_i = tot_s = 0
    for r in rows[1:]:
        if len(r) < len(head):
            continue
        addr = int(r[ci["Address"]], 16)
        if base is None:
            base = addr
        key = table.get(addr - base, ("?", 0))
        inst, thr, smp = int(r[ci["Instructions Executed"]]), int(r[ci["Thread Instructions Executed"]]), int(r[ci["# Samples"]])
        e = per[key]
        e[0] += inst; e[1] += thr; e[2] += smp
        for k, i in stall_cols:
            v = int(r[i]) if r[i].isdigit() else 0
            if v:
                e[3][k] += v
        tot_i += inst; tot_s += smp
    print("kernel %s: %d SASS rows, %.4g warp instructions, %d samples" % (kernel, len(rows) - 1, tot_i, tot_s))
    print("%-34s %7s %7s %6s  %s" % ("source line", "inst%", "smpl%", "thr", "top stalls"))
    for key, e in sorted(per.items(), key=lambda kv: -kv[1][2])[:top]:
        stalls = ", ".join("%s %.0f%%" % (k[6:], 100.0 * v / max(e[2], 1)) for k, v in e[3].most_common(3))
        print("%-34s %6.2f%% %6.2f%% %6.1f  %s" % ("%s:%d" % key, 100.0 * e[0] / tot_i, 100.0 * e[2] / max(tot_s, 1),
                                                    e[1] / max(e[0], 1), stalls))


if __name__ == "__main__":
    main()
