#!/usr/bin/env python
"""Per-source-line / per-opcode warp-instruction counts of one kernel from an .ncu-rep.

usage: line_profile.py <report.ncu-rep> <lib.so> <mangled kernel name> <n_groups> [min_per_group]
n_groups = number of 32-step groups the launch evaluated (series * ceil(T/32)), so the output reads
"warp-instructions per 32 eval steps".
"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

rep, lib, kname, groups = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
min_v = float(sys.argv[5]) if len(sys.argv) > 5 else 2.0
tmp = tempfile.mkdtemp()
subprocess.run(f"cd {tmp} && cuobjdump -xelf all {os.path.abspath(lib)} >/dev/null 2>&1", shell=True, check=True)
cubin = max((f for f in os.listdir(tmp) if f.endswith(".cubin")), key=lambda f: os.path.getsize(os.path.join(tmp, f)))
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
short = re.sub(r"^_ZN\d+[a-z0-9]+?(\d+)", "", kname)
m_ = re.match(r"_ZN3b2p(\d+)", kname)
fname = kname[len(m_.group(0)):len(m_.group(0)) + int(m_.group(1))] if m_ else kname
src_csv = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + fname],
                         capture_output=True, text=True).stdout
rows = list(csv.reader(src_csv.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hi]
ie, ssrc, smp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
insts = []
for r in rows[hi + 1:]:
    if len(r) != len(hdr) or not r[ie].isdigit():
        break
    insts.append((int(r[ie]), int(r[smp]), r[ssrc].strip()))
start = next(i for i, l in enumerate(sass) if l.strip().startswith(".text." + kname + ":"))
cur, seq = ("?", 0), []
for l in sass[start + 1:]:
    if l.startswith("\t.section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
    if m:
        seq.append((cur, m.group(1).strip()))
assert len(seq) == len(insts), (len(seq), len(insts))
by, byop, bysmp = defaultdict(float), defaultdict(float), defaultdict(float)
tot_s = sum(i[1] for i in insts) or 1
for (c, s, _), (loc, st) in zip(insts, seq):
    by[loc] += c / groups
    bysmp[loc] += 100.0 * s / tot_s
    tok = st.split()
    op = tok[1] if tok[0].startswith("@") else tok[0]
    byop[op.split(".")[0]] += c / groups
srcs = {}
def line_text(loc):
    f, n = loc
    if f not in srcs:
        for root in ("greptimedb_b200/csrc", "."):
            p = os.path.join(root, f)
            if os.path.exists(p):
                srcs[f] = open(p).read().split("\n")
                break
        else:
            srcs[f] = []
    return srcs[f][n - 1].strip()[:96] if 0 < n <= len(srcs[f]) else ""
total = sum(by.values())
print(f"kernel {kname}\nSASS instructions {len(seq)}; warp-instructions per 32-step group: {total:.1f}")
for loc in sorted(by):
    if by[loc] >= min_v:
        print(f"{by[loc]:7.1f}  stall {bysmp[loc]:4.1f}%  {loc[0][:16]:16s}:{loc[1]:4d}  {line_text(loc)}")
print("opcode mix:", ", ".join(f"{k} {v:.1f}" for k, v in sorted(byop.items(), key=lambda kv: -kv[1])[:24]))
