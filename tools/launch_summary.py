#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and the ordered timeline."""
import collections
import csv
import re
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = []
    for x in csv.DictReader(lines):
        if x.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v, u = float(x["Metric Value"].replace(",", "")), x["Metric Unit"]
        us = v / 1000 if u.startswith("n") else (v if u.startswith("u") else v * 1000)
        rows.append((int(x["ID"]), x["Kernel Name"], us, x.get("Grid Size"), x.get("Block Size")))
    return rows


def main():
    rows = load(sys.argv[1])
    thr = float(sys.argv[2]) if len(sys.argv) > 2 else 500.0
    tot = sum(r[2] for r in rows)
    print(f"{len(rows)} launches, {tot / 1e3:.3f} ms of kernel time")
    agg = collections.OrderedDict()
    for _, k, us, _, _ in rows:
        name = re.sub(r"\(.*", "", k)
        name = re.sub(r"void |es::|at::native::|\(anonymous namespace\)::", "", name)[:70]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        print(f"{v[1] / 1e3:9.3f} ms {100 * v[1] / tot:5.1f}% n={v[0]:3d}  {k}")
    print()
    for i, k, us, g, b in rows:
        if us > thr:
            print(i, f"{us / 1e3:8.3f} ms", g, b, re.sub(r"void |es::", "", k)[:80])


if __name__ == "__main__":
    main()
