#!/usr/bin/env python
"""Print the SASS of one kernel with source-line annotations collapsed: `sass_view.py <nvdisasm -g -c output> <kernel substr>`.
Each instruction is prefixed with file:line of its innermost location; no inlining chain."""
import re, sys
path, kname = sys.argv[1], sys.argv[2]
lines = open(path).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l)
cur = ""
for l in lines[start + 1:]:
    if l.startswith("//---------------------") or l.startswith("\t.section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = f"{m.group(1).split('/')[-1][4:14]}:{m.group(2)}"
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        print(f"{cur:18s} {m.group(1)} {m.group(2)}")
    elif re.match(r"^\.L_x_\d+:", l):
        print(l)
