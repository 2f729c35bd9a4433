#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total, share)."""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    tot = defaultdict(lambda: [0, 0.0])
    for r in rd:
        if "gpu__time_duration" not in r.get("Metric Name", ""):
            continue
        name = r["Kernel Name"]
        name = re.sub(r"\(.*", "", name)
        name = re.sub(r"^void\s+", "", name)
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        tot[name][0] += 1
        tot[name][1] += us
    total = sum(t for _, t in tot.values())
    print(f"| kernel | launches | total us | share |\n|---|---|---|---|")
    for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k[:110]}` | {n} | {t:.1f} | {100 * t / total:.1f}% |")
    print(f"| **all** | {sum(n for n, _ in tot.values())} | {total:.1f} | 100% |")


if __name__ == "__main__":
    main(sys.argv[1])
