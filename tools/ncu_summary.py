#!/usr/bin/env python
"""Per-kernel table (markdown) of the metrics that matter here from `ncu -i report.ncu-rep --page raw --csv`.
usage: ncu -i prof.ncu-rep --page raw --csv > raw.csv; python tools/ncu_summary.py raw.csv"""
import csv
import re
import sys

WANT = [("us", "gpu__time_duration.sum"), ("grid", "launch__grid_size"), ("block", "launch__block_size"),
        ("regs", "launch__registers_per_thread"), ("warps_active %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("warp instr (M)", "smsp__inst_executed.sum"), ("issue active %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("dram read MB", "dram__bytes_read.sum"), ("dram write MB", "dram__bytes_write.sum"),
        ("dram %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("lsu pipe %", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
        ("fma pipe %", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
        ("l2 %", "lts__t_sectors.avg.pct_of_peak_sustained_elapsed"),
        ("smem bank conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
        ("smem wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    kernels = rows[2:]
    names = [re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "").replace("b200cam::", "")[:40] for r in kernels]
    print("| metric | " + " | ".join(f"`{n}`" for n in names) + " |")
    print("|---|" + "---|" * len(names))
    for label, m in WANT:
        if m not in idx:
            continue
        vals = []
        for r in kernels:
            v, u = r[idx[m]], units[idx[m]]
            try:
                f = float(v.replace(",", ""))
                if label.endswith("(M)"):
                    f /= 1e6
                if u == "byte":
                    f /= 1e6
                if u == "Kbyte":
                    f /= 1e3
                if u == "Gbyte":
                    f *= 1e3
                vals.append(f"{f:.1f}" if abs(f) < 1e4 else f"{f:.0f}")
            except ValueError:
                vals.append(v)
        print(f"| {label} | " + " | ".join(vals) + " |")
    stall = [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
    line = []
    for r in kernels:
        top = sorted(((float(r[idx[h]]), h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")) for h in stall), reverse=True)[:5]
        line.append(", ".join(f"{n} {v:.2f}" for v, n in top))
    print("| top stalls (per issue) | " + " | ".join(line) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
