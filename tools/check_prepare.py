"""Diagnostic: inline structure build vs UNetSCN.prepare() on a side stream, step by step."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mm2d3d_b200 import synth, scn as scn_mod
from mm2d3d_b200.unet import UNetSCN

DEV = "cuda:0"
torch.manual_seed(3)
net = UNetSCN(in_channels=3, m=16, num_planes=4, full_scale=256).to(DEV)
batches = []
for r in range(3):
    locs, feats = synth.make_batch("nuscenes", batch=2, seed0=40 + 2 * r)
    locs[:, :3] //= 16
    if os.environ.get("UNIQUE", "1") == "1":
        import numpy as np
        locs, first = np.unique(locs, axis=0, return_index=True)
        feats = feats[first]
    batches.append((torch.from_numpy(locs).to(DEV), torch.from_numpy(feats).to(DEV)))


def run(prep_of):
    outs = []
    for i in range(7):
        locs, feats = batches[i % 3]
        x = feats.clone().requires_grad_(True)
        net.zero_grad(set_to_none=True)
        out = net([prep_of(i, locs), x])
        out.square().sum().backward()
        outs.append((out.detach().clone(), x.grad.clone(), net.layer2.weight.grad.clone()))
    torch.cuda.synchronize()
    return outs


for mode in ("fp32", "tf32"):
    scn_mod.set_conv_mode(mode)
    a = run(lambda i, l: l)
    b = run(lambda i, l: l)
    same = run(lambda i, l: net.prepare(l))
    side = torch.cuda.Stream(device=DEV, priority=-1)
    ahead = {}

    def prep_of(i, locs):
        if i not in ahead:
            with torch.cuda.stream(side):
                ahead[i] = net.prepare(locs)
        cur = ahead.pop(i)
        with torch.cuda.stream(side):
            ahead[i + 1] = net.prepare(batches[(i + 1) % 3][0])
        return cur

    c = run(prep_of)
    for name, other in (("inline2", b), ("prepare-same-stream", same), ("prepare-side", c)):
        print(mode, name, [tuple(float((x - y).abs().max()) for x, y in zip(p, q)) for p, q in zip(a, other)])
