#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (markdown table)."""
import collections
import csv
import sys


def load(path):
    lines = [ln for ln in open(path) if not ln.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    for r in rows:
        name = r["Kernel Name"]
        v = float(r["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r["Metric Unit"], 1.0)
        c = agg.setdefault(name, [0, 0.0])
        c[0] += 1
        c[1] += v
    return rows, agg


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    rows, agg = load(path)
    tot = sum(v[1] for v in agg.values())
    print("| launches | total us | avg us | share | kernel |\n|---:|---:|---:|---:|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("| %d | %.1f | %.2f | %.1f%% | `%s` |" % (v[0], v[1], v[1] / v[0], 100 * v[1] / tot, k[:100]))
    ours = sum(v[1] for k, v in agg.items() if "dgn::" in k)
    print("\ntotal %.1f us over %d launches; dgn:: kernels %.1f us (%.1f%%)" % (tot, len(rows), ours, 100 * ours / tot))


if __name__ == "__main__":
    main()
