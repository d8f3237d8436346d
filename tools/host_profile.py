#!/usr/bin/env python
"""Host-side cost of one step (cProfile of the enqueue path, GPU drained before each phase)."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from mm2d3d_b200 import synth  # noqa: E402
from mm2d3d_b200 import scn as scn_mod  # noqa: E402
from mm2d3d_b200.unet import UNetSCN  # noqa: E402

dev = torch.device("cuda", 0)
scn_mod.set_conv_mode("tf32")
net = UNetSCN(in_channels=3, m=16, num_planes=7, full_scale=4096).to(dev)
locs, feats = synth.make_batch("nuscenes", batch=8)
locs, feats = torch.from_numpy(locs).to(dev), torch.from_numpy(feats).to(dev)
g = torch.randn(locs.shape[0], 16, device=dev)


def step(timers=None):
    for p in net.parameters():
        p.grad = None
    x = feats.detach().requires_grad_(True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = net([locs, x])
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    out.backward(g)
    t3 = time.perf_counter()
    torch.cuda.synchronize()
    if timers is not None:
        timers.append((t1 - t0, t3 - t2))


for _ in range(5):
    step()
tm = []
for _ in range(20):
    step(tm)
print(f"host time with an idle GPU: forward call {1e3 * np.median([a for a, _ in tm]):.2f} ms (includes the structure "
      f"build's own GPU time + read-back), backward call {1e3 * np.median([b for _, b in tm]):.2f} ms")
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    step()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
