#!/usr/bin/env python3
"""Aggregate an ncu source-page CSV by source line (via nvdisasm --print-line-info of the same cubin).
usage: ncu_by_line.py src.csv disasm.txt kernel_substring [bucket]"""
import collections, csv, re, sys
src, dis, kern = sys.argv[1], sys.argv[2], sys.argv[3]
bucket = int(sys.argv[4]) if len(sys.argv) > 4 else 10
rows = list(csv.reader(open(src)))
kname = [r for r in rows if r and r[0] == "Kernel Name"]
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
ix = {h: i for i, h in enumerate(rows[hi])}
data = rows[hi + 1:]
# functions in the disassembly, each a list of (offset, file:line)
funcs = {}; cur = "?"; name = None
for line in open(dis):
    if line.startswith(".text."):
        name = line.strip(); funcs[name] = []; continue
    if name is None: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/", line)
    if m: funcs[name].append((int(m.group(1), 16), cur))
cands = [k for k in funcs if kern in k]
name = max(cands, key=lambda k: len(funcs[k]))
a2l = dict(funcs[name])
base = int(data[0][ix["Address"]], 16)
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0.0]); tot = 0.0; tots = 0.0
for r in data:
    off = int(r[ix["Address"]], 16) - base
    f, l = a2l.get(off, ("?", 0))
    ie = float(r[ix["Instructions Executed"]] or 0); te = float(r[ix["Thread Instructions Executed"]] or 0)
    sm = float(r[ix["# Samples"]] or 0)
    k = (f, l // bucket * bucket)
    agg[k][0] += ie; agg[k][1] += te; agg[k][2] += sm; tot += ie; tots += sm
print("kernel rows %d, mapped function %s (%d instr)" % (len(data), name[:60], len(funcs[name])))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][2])[:40]:
    print("%-26s %5d  samples %5.1f%%  warp-inst %5.1f%%  lanes %4.1f" % (k[0], k[1], 100 * v[2] / max(tots, 1), 100 * v[0] / max(tot, 1), v[1] / max(v[0], 1)))
