#!/usr/bin/env python3
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total ms, share)."""
import collections
import csv
import re
import sys


def summarise(path):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    total = 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "s": 1e3}[row["Metric Unit"]]
        agg[name][0] += 1
        agg[name][1] += v
        total += v
    out = [f"{'kernel':58s} {'launches':>8s} {'total ms':>10s} {'share':>6s}"]
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append(f"{k[:58]:58s} {n:8d} {t:10.2f} {100 * t / total:5.1f}%")
    out.append(f"{'TOTAL':58s} {sum(v[0] for v in agg.values()):8d} {total:10.2f}")
    return "\n".join(out)


if __name__ == "__main__":
    print(summarise(sys.argv[1]))
