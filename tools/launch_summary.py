#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and the longest launches."""
import collections
import csv
import sys


def main(path, top=20):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    seq = []
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0]
        if "krb" not in name:
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v
        agg[name][0] += 1
        agg[name][1] += ms
        seq.append((name, row.get("Block Size"), row.get("Grid Size"), round(ms, 4)))
    tot = sum(v[1] for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:60s} n={v[0]:5d} {v[1]:10.3f} ms {100 * v[1] / tot:5.1f}%")
    print(f"total {tot:.3f} ms over {len(seq)} launches")
    for s in sorted(seq, key=lambda t: -t[3])[:top]:
        print(s)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 20)
