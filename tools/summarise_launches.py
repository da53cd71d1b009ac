#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown)."""
import collections
import csv
import re
import sys


def main(path: str, marker: str = "k_tie_term") -> None:
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    for row in rows:
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")[:80]
        t = float(row["Metric Value"].replace(",", ""))
        t = t / 1000 if row["Metric Unit"] == "ns" else t
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
    steps = max(1, sum(v[0] for k, v in agg.items() if marker in k))
    tot = sum(v[1] for v in agg.values())
    print(f"launches: {len(rows)}, steps covered: {steps}, sum of kernel time per step: {tot / steps:.1f} us "
          "(ncu: cold caches, serialised - compare shares, not absolutes)\n")
    print("| kernel | launches/step | avg us | us/step | share |")
    print("|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if v[1] / tot < 0.002:
            continue
        print(f"| `{k}` | {v[0] / steps:.1f} | {v[1] / v[0]:.1f} | {v[1] / steps:.1f} | {100 * v[1] / tot:.1f}% |")


if __name__ == "__main__":
    main(*sys.argv[1:])
