#!/usr/bin/env python3
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: count, total, average, share."""
import collections
import csv
import re
import sys


def main(path, top=30):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = [(r["Kernel Name"], float(r["Metric Value"].replace(",", ""))) for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, v in rows:
        k = re.sub(r"\(.*", "", k).replace("void ", "").replace("b200w::<unnamed>::", "").replace("unnamed>::", "")
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v for _, v in agg.values())
    print("%d launches, %.3f ms total (serialised, cold-cache: compare shares)" % (len(rows), tot / 1e6))
    for k, (n, v) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        print("%-46s n=%6d total=%10.1f us avg=%9.2f us share=%5.1f%%" % (k[:46], n, v / 1e3, v / n / 1e3, 100 * v / tot))


if __name__ == "__main__":
    main(sys.argv[1])
