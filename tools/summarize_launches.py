#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (share of step)."""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        unit = row["Metric Unit"]
        v = {"ns": v / 1e3, "us": v, "usecond": v, "ms": v * 1e3, "s": v * 1e6}.get(unit, v / 1e3)
        agg[row["Kernel Name"].split("(")[0][:90]][0] += 1
        agg[row["Kernel Name"].split("(")[0][:90]][1] += v
    tot = sum(v[1] for v in agg.values())
    print("# %s : %d kernels, %.1f us total (cold-cache, serialised: compare SHARES)" % (path, sum(v[0] for v in agg.values()), tot))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:20]:
        print("%-92s n=%4d total=%10.1f us avg=%9.1f us share=%5.1f%%" % (k, v[0], v[1], v[1] / v[0], 100 * v[1] / tot))


if __name__ == "__main__":
    main(sys.argv[1])
