#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of device time)."""
import collections, csv, re, sys

def main(path, top=25):
    lines = open(path).readlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in csv.DictReader(lines[start:]):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        v = float(r["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r["Metric Unit"], 1.0)
        k = (name, r["Grid Size"])
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"total device time {tot:.1f} us over {sum(v[0] for v in agg.values())} launches")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{v[1]:10.1f} us {100 * v[1] / tot:5.1f}%  n={v[0]:3d}  avg {v[1] / v[0]:8.1f} us  {k[0][:56]} grid={k[1]}")

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
