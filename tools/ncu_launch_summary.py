#!/usr/bin/env python
"""Developer tool: per-kernel launch count / mean duration from an `ncu --metrics gpu__time_duration.sum --csv` launch list.

    python tools/ncu_launch_summary.py profiles/rNN/ncu_launches_bench_c2.csv [name-regex]
"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    lines = [ln for ln in open(path, newline="") if ln.startswith('"')]
    per = collections.defaultdict(list)
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r["Metric Unit"], 1.0)
        name = re.sub(r"^void ", "", r["Kernel Name"])
        if pat is None or pat.search(name):
            per[name[:110]].append(v)
    total = sum(sum(v) for v in per.values())
    print(f"| launches | mean us | share | kernel |\n|---|---|---|---|")
    for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        print(f"| {len(v)} | {sum(v) / len(v):.2f} | {100 * sum(v) / total:.1f} % | `{k}` |")


if __name__ == "__main__":
    main()
