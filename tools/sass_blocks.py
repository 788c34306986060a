#!/usr/bin/env python3
"""Map the SASS of a kernel to source lines (nvdisasm --print-line-info output) and print, per
contiguous run of the same source file, how many instructions it holds -- used to see how much code each
block of the lane state machine occupies in the instruction cache."""
import re
import sys

path, kernel = sys.argv[1], sys.argv[2]
step = int(sys.argv[3]) if len(sys.argv) > 3 else 100
infn = False
cur = "?"
idx = 0
marks = []
for line in open(path):
    if line.startswith(".text."):
        infn = kernel in line
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = "%s:%s" % (m.group(1).split("/")[-1], m.group(2))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,5}\*/", line):
        if idx % step == 0:
            marks.append((idx, cur))
        idx += 1
print("instructions:", idx)
for i, c in marks:
    print(i, c)
