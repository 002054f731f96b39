#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total
and mean duration, share of the total.  Usage: launch_summary.py launches.csv [skip_first_n]"""
import csv
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        rows.append((r["Kernel Name"].split("(")[0], v * scale))
    rows = rows[skip:]
    agg = OrderedDict()
    for k, us in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1; a[1] += us
    tot = sum(a[1] for a in agg.values())
    print("%-44s %6s %12s %10s %7s" % ("kernel", "count", "total_us", "mean_us", "share"))
    for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-44s %6d %12.1f %10.1f %6.1f%%" % (k[:44], c, us, us / c, 100 * us / tot))
    print("%-44s %6d %12.1f" % ("TOTAL", len(rows), tot))


if __name__ == "__main__":
    main()
