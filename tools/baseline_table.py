#!/usr/bin/env python
"""Rows of BASELINE.md section 6 from committed bench lines.  usage: python tools/baseline_table.py profiles/r02_bench*.json"""
import json
import sys


def main(paths):
    print("| Config | GPUs | B/GPU | N | t_step (us) | img/s | algorithmic GB/s | % of peak | e2e img/s | CPU ref img/s (cores) | torch-CUDA ref img/s | clocks MHz | source |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    for p in paths:
        d = json.loads([ln for ln in open(p) if ln.startswith("{")][-1])
        c, r = d["config"], d["roofline"]
        cpu = d.get("cpu_baseline") or {}
        tc = d.get("torch_cuda_baseline") or {}
        n = d["n_gpus"]
        print(f"| {c.get('baseline_config', '2')} | {n} | {c['global_batch'] // n} | {c.get('size', 256)} | {d['ms_per_step'] * 1e3:.1f} | "
              f"{d['value']:.0f} | {r['achieved']:.0f} | {100 * r['frac']:.1f} | {d['e2e']['value']:.0f} | "
              f"{cpu.get('value', float('nan')):.1f} ({cpu.get('cores', '-')}) | {tc.get('value', float('nan')):.0f} | "
              f"{d['clocks']['sm_mhz']:.0f} / {d['clocks']['sm_max_mhz']:.0f} {d['clocks']['reasons']} | `{p}` |")


if __name__ == "__main__":
    main(sys.argv[1:])
