#!/usr/bin/env python
"""Turn an ncu CSV (`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`) of
`bench.py --quick` into (a) a per-kernel markdown table and (b) profiles/traffic.json: measured DRAM bytes per step,
which bench.py reports as roofline.traffic next to the algorithmic bytes.
usage: python tools/ncu_traffic.py launches.csv N B [step_marker_kernel]"""
import collections
import csv
import json
import re
import sys
from pathlib import Path


def main(path, N, B, marker="k_cols_accum"):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rows = list(csv.DictReader(lines))
    per = collections.OrderedDict()
    for r in rows:
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("b200cam::", "")[:70]
        val = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        d = per.setdefault(name, collections.defaultdict(float))
        m = r["Metric Name"]
        if m == "gpu__time_duration.sum":
            d["ns"] += val * {"ns": 1, "us": 1e3, "ms": 1e6}.get(unit, 1)
            d["n"] += 1
        else:
            d[m] += val * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    steps = max(1, int(sum(d["n"] for k, d in per.items() if marker in k)))
    tot_bytes = sum(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in per.values()) / steps
    print(f"steps covered: {steps}; measured DRAM traffic per step: {tot_bytes / 1e6:.1f} MB "
          f"(algorithmic: {int(B) * 48 * int(N) ** 2 / 1e6:.1f} MB)\n")
    print("| kernel | launches/step | avg us | DRAM read MB/launch | DRAM write MB/launch |")
    print("|---|---|---|---|---|")
    for k, d in sorted(per.items(), key=lambda kv: -kv[1]["ns"]):
        if d["n"] == 0:
            continue
        print(f"| `{k}` | {d['n'] / steps:.1f} | {d['ns'] / d['n'] / 1e3:.1f} | {d['dram__bytes_read.sum'] / d['n'] / 1e6:.2f} | "
              f"{d['dram__bytes_write.sum'] / d['n'] / 1e6:.2f} |")
    out = Path(__file__).resolve().parent.parent / "profiles" / "traffic.json"
    data = json.loads(out.read_text()) if out.exists() else {}
    data[f"N{N}_B{B}"] = tot_bytes
    out.write_text(json.dumps(data, indent=1))


if __name__ == "__main__":
    main(*sys.argv[1:])
