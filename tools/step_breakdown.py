#!/usr/bin/env python
"""Where one training step spends its time: host enqueue time vs GPU time, per phase
(structure + plans, forward, backward).  python tools/step_breakdown.py --mode tf32"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from mm2d3d_b200 import _lib, synth  # noqa: E402
from mm2d3d_b200 import scn as scn_mod  # noqa: E402
from mm2d3d_b200.metadata import Metadata  # noqa: E402
from mm2d3d_b200.unet import UNetSCN  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="tf32")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--shape", default="nuscenes")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    scn_mod.set_conv_mode(a.mode)
    torch.manual_seed(0)
    net = UNetSCN(in_channels=3, m=16, num_planes=7, full_scale=4096).to(dev)
    batches = []
    for r in range(4):
        locs, feats = synth.make_batch(a.shape, batch=8, seed0=8 * r)
        g = np.random.default_rng(r).standard_normal((locs.shape[0], 16), dtype=np.float32)
        batches.append((torch.from_numpy(locs).to(dev), torch.from_numpy(feats).to(dev), torch.from_numpy(g).to(dev)))

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def step(i, marks=None):
        locs, feats, g = batches[i % 4]
        for p in net.parameters():
            p.grad = None
        x = feats.detach().requires_grad_(True)
        t0 = time.perf_counter()
        e0 = ev()
        out = net([locs, x])
        t1 = time.perf_counter()
        e1 = ev()
        out.backward(g)
        t2 = time.perf_counter()
        e2 = ev()
        if marks is not None:
            marks.append((t1 - t0, t2 - t1, e0, e1, e2))

    for i in range(5):
        step(i)
    torch.cuda.synchronize()
    marks = []
    t0 = time.perf_counter()
    e_begin = ev()
    for i in range(a.steps):
        step(i, marks)
    host_total = time.perf_counter() - t0
    e_end = ev()
    torch.cuda.synchronize()
    gpu_total = e_begin.elapsed_time(e_end)
    hf = sum(m[0] for m in marks) / len(marks) * 1e3
    hb = sum(m[1] for m in marks) / len(marks) * 1e3
    gf = sum(m[2].elapsed_time(m[3]) for m in marks) / len(marks)
    gb = sum(m[3].elapsed_time(m[4]) for m in marks) / len(marks)
    print(f"per step: host enqueue {host_total / a.steps * 1e3:.2f} ms (fwd {hf:.2f}, bwd {hb:.2f}); "
          f"GPU elapsed {gpu_total / a.steps:.2f} ms (fwd incl. structure {gf:.2f}, bwd {gb:.2f})")

    # phases in isolation (GPU time with the host far ahead is not guaranteed; each phase is synchronised)
    def timed(fn, n=10):
        fn()
        torch.cuda.synchronize()
        ts, hs = [], []
        for _ in range(n):
            t0 = time.perf_counter()
            e0 = ev()
            r = fn()
            h = time.perf_counter() - t0
            e1 = ev()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
            hs.append(h * 1e3)
        return sorted(ts)[n // 2], sorted(hs)[n // 2], r

    locs, feats, g = batches[0]
    t, h, meta = timed(lambda: Metadata(locs, 4096, 7))
    print(f"structure (hash, 7 levels, nbr27, one sync): GPU {t:.2f} ms, host {h:.2f} ms")

    def plans():
        m = Metadata(locs, 4096, 7)
        specs, s = [], 4096
        for l in range(7):
            specs.append(("smc", s))
            if l < 6:
                specs += [("down", s), ("up", s)]
            s //= 2
        m.build_plans(specs)
        return m
    t2, h2, _ = timed(plans)
    print(f"structure + 19 plans: GPU {t2:.2f} ms, host {h2:.2f} ms  -> plans {t2 - t:.2f} ms GPU")
    launches0 = _lib.lib.mm3d_kernel_launches()
    step(0)
    torch.cuda.synchronize()
    print("libmm3d launches per step:", _lib.lib.mm3d_kernel_launches() - launches0)


if __name__ == "__main__":
    main()
