#!/usr/bin/env python
"""Summarise an ncu launch-list CSV (gpu__time_duration [+ tensor pipe %]) -- helper for profiles/."""
import collections
import csv
import re
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    recs = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = row["Kernel Name"]
        short = re.sub(r"\(.*", "", name).replace("void ", "").replace("mmad::<unnamed>::", "").replace("unnamed>::", "")[:34]
        d = recs.setdefault(row["ID"], {"name": short, "grid": row["Grid Size"]})
        v = float(row["Metric Value"].replace(",", ""))
        if row["Metric Name"].startswith("gpu__time"):
            u = row["Metric Unit"]
            d["us"] = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        else:
            d["tens"] = v
    return list(recs.values())


if __name__ == "__main__":
    L = load(sys.argv[1])
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 22
    for d in L[-n:]:
        print(f"{d['name']:36s} {d['grid']:14s} {d.get('us', 0):9.1f} us  tensor {d.get('tens', 0):5.1f}%")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for d in L:
        agg[d["name"]][0] += 1
        agg[d["name"]][1] += d.get("us", 0)
    tot = sum(v[1] for v in agg.values())
    print("--- totals over the capture ---")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:36s} n={v[0]:4d} {v[1]:10.1f} us  share {v[1] / tot:.3f}")
