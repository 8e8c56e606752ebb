#!/usr/bin/env python
"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`):
launch count, mean duration and share of the listed GPU time per kernel.

    python tools/ncu_launches.py gpurun_out/launches.csv ["command line that produced it"]
"""
import csv
import sys


def main():
    path = sys.argv[1]
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = {n: i for i, n in enumerate(rows[0])}
    agg = {}
    for r in rows[1:]:
        if r[hdr["Metric Name"]] != "gpu__time_duration.sum":
            continue
        ns = float(r[hdr["Metric Value"]].replace(",", ""))
        if r[hdr["Metric Unit"]] in ("us", "usecond"):
            ns *= 1e3
        elif r[hdr["Metric Unit"]] in ("ms", "msecond"):
            ns *= 1e6
        a = agg.setdefault(r[hdr["Kernel Name"]], [0, 0.0])
        a[0] += 1
        a[1] += ns
    total = sum(a[1] for a in agg.values()) or 1.0
    if len(sys.argv) > 2:
        print(sys.argv[2])
    print("(serialised, cold-cache launch times under ncu: compare shares, not absolutes)\n")
    for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n:4d} x {ns / n / 1e3:10.1f} us avg  {100 * ns / total:6.2f}%  {name}")


if __name__ == "__main__":
    main()
