#!/usr/bin/env python
"""Development aid (library built with MM3D_EXTRA_NVCC_FLAGS=-DMM3D_TRACE): in-step timeline of the main stream -- one
CUDA event after every operation of a forward + backward (warm caches, weight gradients running beside on the side
stream), printed as the time between consecutive events."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mm2d3d_b200 import _lib, synth  # noqa: E402
from mm2d3d_b200 import scn as scn_mod  # noqa: E402
from mm2d3d_b200.unet import UNetSCN  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    scn_mod.set_conv_mode(sys.argv[1] if len(sys.argv) > 1 else "tf32")
    torch.manual_seed(0)
    net = UNetSCN(in_channels=3).to(dev)
    batches = []
    for r in range(3):
        locs, feats = synth.make_batch("nuscenes", batch=8, seed0=8 * r)
        batches.append((torch.from_numpy(locs).to(dev), torch.from_numpy(feats).to(dev)))
    g = [torch.randn(b[0].shape[0], 16, device=dev) for b in batches]
    lib = _lib.lib

    def step(i):
        locs, feats = batches[i % 3]
        prep = net.prepare(locs, wait=True)
        torch.cuda.synchronize()
        for p in net.parameters():
            p.grad = None
        x = feats.clone().requires_grad_(True)
        out = net([prep, x])
        out.backward(g[i % 3])

    for i in range(4):
        step(i)
    torch.cuda.synchronize()
    lib.mm3d_debug_marks(1)
    step(4)
    lib.mm3d_debug_dump_marks()
    lib.mm3d_debug_marks(0)


if __name__ == "__main__":
    main()
