#!/usr/bin/env python
"""Host-side time of the pipelined training step, phase by phase (the bench's loop: structure of step i+1 enqueued on
a second stream, then forward and backward of step i, host kept one step ahead).  Tells whether the step is bound by
the GPU or by the enqueueing host thread."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from mm2d3d_b200 import synth  # noqa: E402
from mm2d3d_b200 import scn as scn_mod  # noqa: E402
from mm2d3d_b200.dp import FlatGradAllReduce  # noqa: E402
from mm2d3d_b200.unet import UNetSCN  # noqa: E402

dev = torch.device("cuda", 0)
scn_mod.set_conv_mode("tf32")
net = UNetSCN(in_channels=3, m=16, num_planes=7, full_scale=4096).to(dev)
flat = FlatGradAllReduce(net)
batches = []
for r in range(4):
    locs, feats = synth.make_batch("nuscenes", batch=8, seed0=8 * r)
    batches.append((torch.from_numpy(locs).to(dev), torch.from_numpy(feats).to(dev),
                    torch.randn(locs.shape[0], 16, device=dev)))
prep_stream = torch.cuda.Stream(device=dev)
prepared, in_flight = {}, []
T = {k: [] for k in ("zero", "prepare", "forward", "backward", "throttle", "total")}


def prepare(i):
    with torch.cuda.stream(prep_stream):
        prepared[i] = net.prepare(batches[i % 4][0])


def step(i, rec):
    t0 = time.perf_counter()
    flat.zero_()
    x = batches[i % 4][1].detach().requires_grad_(True)
    t1 = time.perf_counter()
    if i not in prepared:
        prepare(i)
    cur = prepared.pop(i)
    prepare(i + 1)
    t2 = time.perf_counter()
    out = net([cur, x])
    t3 = time.perf_counter()
    out.backward(batches[i % 4][2])
    flat.all_reduce_mean()
    t4 = time.perf_counter()
    ev = torch.cuda.Event()
    ev.record()
    in_flight.append(ev)
    if len(in_flight) > 1:
        in_flight.pop(0).synchronize()
    t5 = time.perf_counter()
    if rec:
        for k, v in zip(T, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t5 - t0)):
            T[k].append(v)


for i in range(10):
    step(i, False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(10, 110):
    step(i, True)
e1.record()
torch.cuda.synchronize()
print(f"GPU: {e0.elapsed_time(e1) / 100:.3f} ms per step")
for k, v in T.items():
    print(f"host {k:>9}: median {1e3 * np.median(v):.3f} ms   mean {1e3 * np.mean(v):.3f} ms")
print("(throttle = waiting for the GPU to finish the previous step; ~0 means the host is the bottleneck)")
if len(sys.argv) > 1 and sys.argv[1] == "profile":
    import cProfile
    import pstats
    pr = cProfile.Profile()
    pr.enable()
    for i in range(110, 160):
        step(i, False)
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(25)
