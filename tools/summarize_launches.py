#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.

    python tools/summarize_launches.py gpurun_out/launches.csv > profiles/<name>.txt
"""
import collections
import csv
import re
import sys


def main(path, top=60):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    total, n = 0.0, 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        agg[name][0] += 1
        agg[name][1] += v
        total += v
        n += 1
    print(f"# {path}: {n} launches, {total:.1f} us of kernel time (cold-cache, serialised: compare shares)")
    print(f"# {'us':>10} {'share':>6} {'n':>5} {'avg us':>9}  kernel")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{t:12.1f} {100 * t / total:5.1f}% {c:5d} {t / c:9.1f}  {k[:110]}")


if __name__ == "__main__":
    main(sys.argv[1])
