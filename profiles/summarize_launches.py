#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (share of the step)."""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = row["Kernel Name"].split("(")[0][:70]
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1000 if u in ("ns", "nsecond") else v * 1000 if u in ("ms", "msecond") else v
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"{'kernel':72s} {'n':>5s} {'total_us':>10s} {'avg_us':>8s} {'share':>6s}")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:72s} {n:5d} {t:10.1f} {t / n:8.2f} {t / tot * 100:5.1f}%")
    print(f"{'TOTAL':72s} {sum(a[0] for a in agg.values()):5d} {tot:10.1f}")


if __name__ == "__main__":
    main(sys.argv[1])
