#!/usr/bin/env python
"""Per-instruction view of one kernel from an ncu report: address-ordered SASS with executed count per group,
sample share and the dominant stall reason.  usage: stall_view.py <rep> <kernel regex> <n_groups> [min_pct]"""
import csv, subprocess, sys
rep, kre, groups = sys.argv[1], sys.argv[2], float(sys.argv[3])
minp = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = 0
data = []
for r in rows[hi + 1:]:
    if len(r) != len(hdr) or not r[ix["Instructions Executed"]].isdigit():
        break
    smp = int(r[ix["# Samples"]])
    tot += smp
    data.append(r)
agg = {s: 0 for s in stalls}
for r in data:
    for s in stalls:
        agg[s] += int(r[ix[s]] or 0)
print("total samples", tot, " instructions/group %.1f" % (sum(int(r[ix["Instructions Executed"]]) for r in data) / groups))
print("stall mix: " + ", ".join(f"{k[6:]} {100*v/tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v * 200 > tot))
for n, r in enumerate(data):
    smp = int(r[ix["# Samples"]])
    ex = int(r[ix["Instructions Executed"]]) / groups
    if 100.0 * smp / tot < minp:
        continue
    top = max(stalls, key=lambda s: int(r[ix[s]] or 0))
    print(f"{n:5d} {ex:6.2f}/grp {100.0*smp/tot:5.2f}%  {top[6:]:14s} {r[ix['Source']].strip()}")
