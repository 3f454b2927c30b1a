#!/usr/bin/env python
"""Aggregates `ncu -i X.ncu-rep --page source --csv --kernel-name regex:K` (SASS level): stall
reasons, opcode histogram by executed warp instructions, hottest instructions."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break                      # first launch only
    if len(r) == len(hdr) and r[0] != hdr[0]:
        data.append(r)
iS, iI, iSrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
num = lambda v: int(float(v)) if v not in ("", "-") else 0
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
print("sass lines", len(data), "samples", sum(num(r[iS]) for r in data), "warp instructions", sum(num(r[iI]) for r in data))
agg = {}
for r in data:
    for i in stall_cols:
        agg[hdr[i]] = agg.get(hdr[i], 0) + num(r[i])
print("stalls:", sorted(agg.items(), key=lambda kv: -kv[1])[:10])
op = {}
for r in data:
    t = r[iSrc].split()
    o = t[1] if t[0].startswith("@") else t[0]
    op[o] = op.get(o, 0) + num(r[iI])
print("opcodes:", sorted(op.items(), key=lambda kv: -kv[1])[:45])
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print("top sampled:")
for r in sorted(data, key=lambda r: -num(r[iS]))[:n]:
    print(r[iS], r[iI], r[iSrc][:100])
