#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.

usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/launches_rNN.md
"""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot, n = 0.0, 0
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        unit = row["Metric Unit"]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
        n += 1
    print(f"launches: {n}, summed device time: {tot / 1e3:.2f} ms (cold-cache, serialised under ncu)\n")
    print("| kernel | launches | total ms | avg us | share |")
    print("|---|---:|---:|---:|---:|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {c} | {t / 1e3:.2f} | {t / c:.1f} | {t / tot * 100:.1f}% |")


if __name__ == "__main__":
    main(sys.argv[1])
