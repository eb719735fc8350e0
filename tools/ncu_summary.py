#!/usr/bin/env python3
"""Summarise an .ncu-rep (ncu --set full) per kernel launch: python tools/ncu_summary.py report.ncu-rep [--md]"""
import csv, subprocess, sys, io

COLS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"), ("sm__inst_executed.avg.per_cycle_elapsed", "ipc"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("launch__registers_per_thread", "regs"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bankconf"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"), ("launch__grid_size", "grid")]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("| kernel | " + " | ".join(n for _, n in COLS) + " |")
    print("|---|" + "---|" * len(COLS))
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0]
        vals = []
        for c, n in COLS:
            if c not in idx:
                vals.append("-"); continue
            v = r[idx[c]]; u = units[idx[c]]
            try:
                f = float(v.replace(",", ""))
                if n == "us":
                    f = f * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
                if n in ("rdMB", "wrMB"):
                    f = f * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1)
                vals.append(f"{f:.1f}" if abs(f) < 1e6 else f"{f:.3g}")
            except ValueError:
                vals.append(v)
        print(f"| {name} | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main()
