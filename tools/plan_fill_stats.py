#!/usr/bin/env python
"""Fill statistics of the submanifold row plans of one batch-8 nuScenes-shaped input: non-empty (tile, offset) blocks
per tile and the share of present rows inside them -- for the plan as built, and for what-if orderings emulated with
torch sorts (offset subsets processed in separate passes)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mm2d3d_b200 import synth  # noqa: E402
from mm2d3d_b200.metadata import Metadata  # noqa: E402


def popc(x):
    x = x.to(torch.int64) & 0xFFFFFFFF
    c = torch.zeros_like(x)
    for b in range(27):
        c += (x >> b) & 1
    return c


def tiles_union(keys_sorted):
    n = keys_sorted.numel()
    pad = (-n) % 128
    k = torch.cat([keys_sorted, keys_sorted.new_zeros(pad)]).view(-1, 128)
    u = torch.zeros(k.shape[0], dtype=torch.int64, device=k.device)
    for b in range(27):
        u |= (((k >> b) & 1).amax(1)) << b
    return u


def main():
    dev = torch.device("cuda", 0)
    locs, _ = synth.make_batch(sys.argv[1] if len(sys.argv) > 1 else "nuscenes", batch=8)
    meta = Metadata(torch.from_numpy(locs).to(dev), 4096, 7)
    cls = torch.tensor([abs(k // 9 - 1) + abs((k // 3) % 3 - 1) + abs(k % 3 - 1) for k in range(27)], device=dev)
    sub = {"centre+faces": sum(1 << k for k in range(27) if cls[k] <= 1), "edges": sum(1 << k for k in range(27) if cls[k] == 2),
           "corners": sum(1 << k for k in range(27) if cls[k] == 3)}
    s = 4096
    for l in range(7):
        nbr = meta.nbr_table(s)  # [n, 27]
        n = nbr.shape[0]
        rowmask = torch.zeros(n, dtype=torch.int64, device=dev)
        for k in range(27):
            rowmask |= (nbr[:, k] >= 0).to(torch.int64) << k
        pairs = int(popc(rowmask).sum())
        perm, tmask, tbl, order = meta.plan_tensors("smc", s)
        T = (n + 127) // 128
        items = int(popc(tmask[:T]).sum())
        line = f"L{l} rows {n:7d} pairs/row {pairs / n:5.2f} | plan: items/tile {items / T:5.2f} fill {pairs / (items * 128):.2f}"
        # what-if: three passes over offset subsets, rows of a pass sorted by their sub-mask
        tot = 0
        for name, sm in sub.items():
            km = rowmask & sm
            km = km[km != 0]
            u = tiles_union(torch.sort(km).values)
            tot += int(popc(u).sum())
        line += f" | 3 passes: items {tot} ({tot / items:.2f}x) fill {pairs / (tot * 128):.2f}"
        # what-if: one pass, rows sorted globally by the full mask (the plan sorts within chunks of 8192 rows by a class-ordered key)
        u = tiles_union(torch.sort(rowmask).values)
        line += f" | global sort: items {int(popc(u).sum())}"
        # the plan's 19-bit key (bit 18 = any corner, 17..6 edges, 5..0 faces), sorted globally / two-level (a stable global
        # sort by the key's top bits, then a stable sort by the full key inside chunks of 8192 rows)
        key = torch.zeros(n, dtype=torch.int64, device=dev)
        edge, face = 17, 5
        for k in range(27):
            c = int(cls[k])
            if c == 0:
                continue
            if c == 3:
                pos = 18
            elif c == 2:
                pos = edge
                edge -= 1
            else:
                pos = face
                face -= 1
            key |= ((rowmask >> k) & 1) << pos
        def items_of(order):
            return int(popc(tiles_union(rowmask[order])).sum())
        line += f" | global by key: {items_of(torch.sort(key, stable=True).indices)}"
        for top in (4, 7, 10, 13):
            coarse = key >> (19 - top)
            o1 = torch.sort(coarse, stable=True).indices
            k1 = key[o1]
            chunks = []
            for c0 in range(0, n, 8192):
                kk = k1[c0:c0 + 8192]
                chunks.append(o1[c0:c0 + 8192][torch.sort(kk, stable=True).indices])
            line += f" | top{top}+chunks: {items_of(torch.cat(chunks))}"
        print(line)
        s //= 2


if __name__ == "__main__":
    main()
