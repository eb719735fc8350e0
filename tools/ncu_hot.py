#!/usr/bin/env python3
"""Hot SASS instructions of one kernel launch in an .ncu-rep: python tools/ncu_hot.py report.ncu-rep kernel_regex [launch_index] [min_pct]
Prints every instruction holding >= min_pct of the stall samples, in program order, with BAR/branch markers for orientation."""
import csv, subprocess, sys, io

rep, kern = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
minpct = float(sys.argv[4]) if len(sys.argv) > 4 else 0.7
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kern], capture_output=True, text=True).stdout
tables, cur = [], None
for row in csv.reader(io.StringIO(out)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": []}; tables.append(cur); continue
    if cur is not None:
        cur["rows"].append(row)
t = tables[which]
hdr = t["rows"][0]
ix = {h: i for i, h in enumerate(hdr)}
rows = [r for r in t["rows"][1:] if len(r) > 5]
samp = ix["# Samples"]
tot = sum(float(r[samp] or 0) for r in rows)
stall_cols = [(h, i) for h, i in ix.items() if h.startswith("stall_") and "Not Issued" not in h]
print(t["name"][:120], "| launches in report:", len(tables), "| total samples", tot, "| instructions", len(rows))
acc = 0.0
for k, r in enumerate(rows):
    v = float(r[samp] or 0)
    acc += v
    ins = r[ix["Source"]].strip()
    mark = any(m in ins for m in ("BAR.SYNC", "MATCH", "EXIT", "ATOM", "RED."))
    if 100 * v / tot >= minpct or (mark and v > 0):
        top = sorted(((float(r[i] or 0), h) for h, i in stall_cols), reverse=True)[:2]
        why = " ".join(f"{h[6:]}={int(x)}" for x, h in top if x > 0)
        print(f"{k:5d} {100*v/tot:5.1f}% cum {100*acc/tot:5.1f}%  {ins[:70]:70s} {why}")
