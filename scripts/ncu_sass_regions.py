#!/usr/bin/env python
"""Splits the SASS of each kernel in `ncu --page source --csv` output at BAR.SYNC instructions and prints, per region,
the share of stall samples, executed instructions, the opcode mix and the dominant stall reasons.
Usage: ncu -i REP --page source --csv > f.csv; scripts/ncu_sass_regions.py f.csv [kernel-index]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
# kernels are separated by "Kernel Name" rows
kernels, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        kernels.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] and len(r) == len(cur["hdr"]):
        cur["rows"].append(r)
k = kernels[which]
h = k["hdr"]
si, ii, src = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
stall_cols = [(i, x) for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
tot = sum(float(r[si]) for r in k["rows"])
toti = sum(float(r[ii]) for r in k["rows"])
print(k["name"][:100], "samples", tot, "warp-instructions", toti)
regions, cur = [], []
for r in k["rows"]:
    cur.append(r)
    if "BAR.SYNC" in r[src]:
        regions.append(cur)
        cur = []
if cur:
    regions.append(cur)
for n, reg in enumerate(regions):
    s = sum(float(r[si]) for r in reg)
    ins = sum(float(r[ii]) for r in reg)
    if s / tot < 0.004:
        continue
    ops = collections.Counter()
    for r in reg:
        op = r[src].split()
        op = [o for o in op if not o.startswith("@")][0].split(".")[0]
        ops[op] += float(r[ii])
    st = collections.Counter()
    for i, x in stall_cols:
        st[x] += sum(float(r[i]) for r in reg)
    top = ", ".join(f"{a}:{b / max(ins, 1) * 100:.0f}%" for a, b in ops.most_common(6))
    stt = ", ".join(f"{a[6:]}:{b / max(s, 1) * 100:.0f}%" for a, b in st.most_common(4))
    print(f"region {n:3d} [{len(reg):5d} sass] samples {s / tot * 100:5.1f}%  instr {ins / toti * 100:5.1f}%  | {top} | {stt}")
