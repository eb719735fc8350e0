#!/usr/bin/env python3
"""Summarise an ncu launch list (ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv --log-file X.csv ...)
per kernel: python tools/launch_summary.py X.csv [out.md] [traffic.json]"""
import csv, sys, json, collections, re

src = sys.argv[1]
rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
per = collections.defaultdict(lambda: collections.defaultdict(float))
launches = collections.defaultdict(set)
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").replace("bz::", "")
    name = re.sub(r"<[^>]*>", "", name)   # template variants together, as bz2b200_kernel_times reports them
    metric, unit, val = r[ix["Metric Name"]], r[ix["Metric Unit"]], float(r[ix["Metric Value"]].replace(",", ""))
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    per[name][metric] += val * scale
    launches[name].add(r[ix["ID"]])
tot = sum(v["gpu__time_duration.sum"] for v in per.values())
lines = [f"# ncu launch list summary ({src}); cold-cache, serialised launches: compare SHARES, not absolute times", "",
         "| kernel | launches | total ms | share | DRAM read MB/launch | DRAM write MB/launch |", "|---|---:|---:|---:|---:|---:|"]
traffic = {}
for name, v in sorted(per.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
    n = len(launches[name])
    rd, wr = v.get("dram__bytes_read.sum"), v.get("dram__bytes_write.sum")
    lines.append(f"| {name} | {n} | {v['gpu__time_duration.sum']:.3f} | {100*v['gpu__time_duration.sum']/tot:.1f}% | "
                 f"{'-' if rd is None else f'{rd/n/1e6:.1f}'} | {'-' if wr is None else f'{wr/n/1e6:.1f}'} |")
    if rd is not None:
        traffic[name] = {"launches": n, "dram_bytes_per_launch": (rd + wr) / n, "ms_per_launch": v["gpu__time_duration.sum"] / n}
lines.append(f"\ntotal {tot:.3f} ms over {sum(len(s) for s in launches.values())} launches")
text = "\n".join(lines)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text + "\n")
else:
    print(text)
if len(sys.argv) > 3 and traffic:
    json.dump({"source": f"ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum over the bench command ({src})",
               "kernels": traffic}, open(sys.argv[3], "w"), indent=1)
