#!/usr/bin/env python
"""Attribute ncu per-SASS-instruction counts to CUDA source lines.

usage: attribute_sass.py <ncu_source_page_sass.csv> <nvdisasm -g -c output> <mangled kernel name> [top_n]
The ncu CSV comes from `ncu -i rep --page source --csv --kernel-name ... --launch-count 1` (SASS view,
one row per instruction, in address order); nvdisasm -g gives `//## File "...", line N` markers in the
same instruction order.
"""
import csv
import re
import sys
from collections import defaultdict

ncu_csv, sass_path, kname = sys.argv[1:4]
top_n = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(ncu_csv)))
hdr_i = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hdr_i]
ie, smp, src = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
insts = []
for r in rows[hdr_i + 1:]:
    if len(r) != len(hdr) or not r[ie].isdigit():
        break  # next kernel / next launch section
    insts.append((int(r[ie]), int(r[smp]), r[src].strip()))

lines = open(sass_path).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.strip().startswith(".text." + kname + ":") or l.strip() == kname + ":")
cur = ("?", 0)
seq = []
inl = ""
for l in lines[start + 1:]:
    if l.startswith("\t.section") or l.strip().startswith("//--------------------- .text."):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        inl = m.group(3)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
    if m:
        seq.append((cur, m.group(1).strip()))
print(f"ncu instructions: {len(insts)}, nvdisasm instructions: {len(seq)}")
n = min(len(insts), len(seq))
by_line = defaultdict(lambda: [0, 0, 0])
tot = sum(i[0] for i in insts)
tot_s = sum(i[1] for i in insts)
for (cnt, s, text), (loc, stext) in zip(insts[:n], seq[:n]):
    b = by_line[loc]
    b[0] += cnt
    b[1] += s
    b[2] += 1
print(f"total warp-instructions executed: {tot}, stall samples: {tot_s}")
for loc, (cnt, s, k) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top_n]:
    print(f"{cnt:>13d} {100 * cnt / tot:5.1f}%  samples {100 * s / max(tot_s, 1):5.1f}%  sass {k:4d}  {loc[0]}:{loc[1]}")
