#!/usr/bin/env python
"""How much of a step is the sparse-structure build?  Times forward+backward of the bench workload with the structure of
every batch prepared ONCE and reused (no build in the timed loop) against the normal pipelined loop (one build per step
on the prepare stream).  A development measurement, not a benchmark line."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from mm2d3d_b200 import scn as scn_mod, synth  # noqa: E402
from mm2d3d_b200.unet import UNetSCN  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    scn_mod.set_conv_mode("tf32")
    torch.manual_seed(0)
    net = UNetSCN(in_channels=3, m=16, block_reps=1, residual_blocks=False, full_scale=4096, num_planes=7).to(dev)
    res = []
    for r in range(4):
        locs, feats = synth.make_batch("nuscenes", batch=8, seed0=r * 8)
        g = np.random.default_rng(r).standard_normal((locs.shape[0], 16), dtype=np.float32)
        res.append((torch.from_numpy(locs).to(dev), torch.from_numpy(feats).to(dev), torch.from_numpy(g).to(dev)))
    prep_stream = torch.cuda.Stream(device=dev)
    fixed = [net.prepare(r[0]) for r in res]
    torch.cuda.synchronize()

    def step(i, reuse, nxt):
        locs, feats, g = res[i % 4]
        x = feats.detach().requires_grad_(True)
        if reuse:
            cur = fixed[i % 4]
        else:
            cur = nxt.pop(i) if i in nxt else net.prepare(locs)
            with torch.cuda.stream(prep_stream):
                nxt[i + 1] = net.prepare(res[(i + 1) % 4][0])
        out = net([cur, x])
        out.backward(g)
        for p in net.parameters():
            p.grad = None

    for reuse in (True, False, True, False):
        nxt = {}
        for i in range(5):
            step(i, reuse, nxt)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 40
        for i in range(5, 5 + n):
            step(i, reuse, nxt)
            if i % 4 == 0:
                torch.cuda.current_stream().synchronize()  # keep the host at most a few steps ahead
        e1.record()
        torch.cuda.synchronize()
        print(f"{'structure reused (no build)' if reuse else 'one build per step (pipelined)'}: {e0.elapsed_time(e1) / n:.3f} ms per step")


if __name__ == "__main__":
    main()
