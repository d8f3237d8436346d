#!/usr/bin/env python
"""Opcode histogram of libmm3d.so's SASS (cuobjdump -sass), per kernel for the opcodes that show which hardware paths
the code uses: UTCHMMA / UTCBAR (tcgen05.mma / commit), LDTM (tcgen05.ld), UBLKCP (cp.async.bulk), UTMALDG (TMA tensor
loads), LDGSTS (cp.async), SYNCS (mbarrier), REDG / ATOMG (global reductions / atomics).

    python tools/sass_histogram.py > profiles/<name>.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mm2d3d_b200", "libmm3d.so")
KEY = ("UTCHMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMAPF", "LDGSTS", "SYNCS", "REDG", "ATOMG", "RED", "ATOM",
       "ELECT", "ACQBULK", "FENCE", "UTCATOMSWS")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    total = collections.Counter()
    per = collections.defaultdict(collections.Counter)
    fn = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            sym = m.group(1)
            k = re.search(r"(k_[a-z0-9_]+)", sym)
            fn = k.group(1) if k else sym[:40]
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and fn:
            op = m.group(1)
            total[op] += 1
            if op in KEY:
                per[fn][op] += 1
    sha = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    print(f"# cuobjdump -sass mm2d3d_b200/libmm3d.so (built for sm_100a from the tree at {sha}); {sum(total.values())} instructions, "
          f"{len(total)} distinct opcodes")
    print("# per-kernel counts of the opcodes that identify the hardware path:")
    for fn_, c in sorted(per.items(), key=lambda kv: -sum(kv[1].values())):
        print(f"  {fn_:<28} " + "  ".join(f"{o} {n}" for o, n in sorted(c.items(), key=lambda kv: -kv[1])))
    print("# all opcodes:")
    print("  " + "  ".join(f"{o} {n}" for o, n in total.most_common()))


if __name__ == "__main__":
    main()
