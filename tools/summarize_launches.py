#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name:
launch count, total and share of device time.  Usage: summarize_launches.py launches.csv [out.md]"""
import csv
import re
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"^void\s+", "", name)
        name = re.sub(r"<.*", "", name).replace("pcrcg::", "")
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        rows.append((name, val * scale))
    tot = sum(v for _, v in rows) or 1.0
    agg = defaultdict(lambda: [0, 0.0])
    for n, v in rows:
        agg[n][0] += 1
        agg[n][1] += v
    out = [f"# launch list summary: {path}", "", f"launches: {len(rows)}  total device time: {tot / 1e3:.3f} ms", "",
           "| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
    for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| {n} | {c} | {v:.1f} | {100 * v / tot:.1f}% |")
    text = "\n".join(out) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
