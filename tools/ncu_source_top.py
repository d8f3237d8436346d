#!/usr/bin/env python
"""Top SASS instructions by stall samples from `ncu -i X.ncu-rep --page source --csv`."""
import csv
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    total = 0
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        try:
            s = int(r[idx["# Samples"]])
        except ValueError:
            continue
        total += s
        data.append((s, r))
    print("total samples", total)
    for lineno, (s, r) in enumerate(data):
        pass
    order = sorted(range(len(data)), key=lambda i: -data[i][0])[:top]
    for i in sorted(order):
        s, r = data[i]
        stalls = sorted(((int(r[idx[c]]), c[6:]) for c in stall_cols if r[idx[c]].isdigit() and int(r[idx[c]]) > 0), reverse=True)[:3]
        print(f"{i:5d} {100*s/total:5.1f}% exec {r[idx['Instructions Executed']]:>9s}  {r[idx['Source']].strip()[:70]:70s} {stalls}")


if __name__ == "__main__":
    main()
