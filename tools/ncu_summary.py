#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (shares of the step)."""
import csv, collections, sys, json
src, dst = sys.argv[1], sys.argv[2]
lines = [l for l in open(src) if l.startswith('"')]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = row["Kernel Name"].split("(")[0].replace("void ", "")
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    ms = v / 1e6 if unit == "ns" else v / 1e3 if unit == "us" else v
    agg[name][0] += 1
    agg[name][1] += ms
tot = sum(v[1] for v in agg.values())
with open(dst, "w") as f:
    f.write(f"# ncu launch list summary ({src}); cold-cache, serialised: compare SHARES\n\n| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| {k} | {v[0]} | {v[1]:.3f} | {100 * v[1] / tot:.1f}% |\n")
    f.write(f"| **total** | {sum(v[0] for v in agg.values())} | {tot:.3f} | 100% |\n")
print(open(dst).read())
