#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (markdown)."""
import csv
import sys
from collections import OrderedDict


def main(path, skip=0):
    rows = []
    with open(path) as fh:
        lines = [l for l in fh if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((r["Kernel Name"].split("(")[0].replace("void ", ""), r["Grid Size"], r["Block Size"],
                         float(r["Metric Value"]) / 1e3))
    rows = rows[skip:]
    agg = OrderedDict()
    for k, g, b, us in rows:
        a = agg.setdefault((k, g, b), [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(v[1] for v in agg.values())
    print("| kernel | grid | block | launches | avg us | share |")
    print("|---|---|---|---|---|---|")
    for (k, g, b), (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %s | %s | %d | %.1f | %.1f%% |" % (k, g, b, n, us / n, 100 * us / tot))
    print("\ntotal %.1f us over %d launches" % (tot, len(rows)))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
