#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, agg = None, collections.defaultdict(list)
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            try:
                v = float(d["Metric Value"].replace(",", ""))
            except ValueError:
                continue
            unit = d["Metric Unit"]
            v = v / 1000 if unit == "ns" else v * 1000 if unit == "ms" else v
            agg[d["Kernel Name"].split("(")[0]].append(v)
    tot = sum(sum(v) for v in agg.values())
    print("%-40s %5s %12s %8s" % ("kernel", "n", "mean_us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("%-40s %5d %12.2f %7.1f%%" % (k[:40], len(v), sum(v) / len(v), 100 * sum(v) / tot))
    if agg:
        print("total per frame (us): %.1f" % (tot / max(len(v) for v in agg.values())))


if __name__ == "__main__":
    main(sys.argv[1])
