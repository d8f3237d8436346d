"""Find reads of uninitialised memory without a sanitizer: fill the caching allocator's free blocks with a byte
pattern before the forward and before the backward, and compare the results across patterns (0xFF = NaN floats /
-1 indices, 0x00 = zeros / row 0: both are safe as indices)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mm2d3d_b200 import synth, scn as scn_mod
from mm2d3d_b200.unet import UNetSCN

DEV = "cuda:0"
torch.manual_seed(3)
net = UNetSCN(in_channels=3, m=16, num_planes=4, full_scale=256).to(DEV)
batches = []
for r in range(2):
    locs, feats = synth.make_batch("nuscenes", batch=2, seed0=40 + 2 * r)
    locs[:, :3] //= 16
    batches.append((torch.from_numpy(locs).to(DEV), torch.from_numpy(feats).to(DEV)))


def poison(byte):
    if byte is None:
        return
    torch.cuda.synchronize()
    ts = [torch.full((sz,), byte, dtype=torch.uint8, device=DEV) for sz in (1 << 30, 256 << 20, 64 << 20, 16 << 20, 4 << 20, 1 << 20)]
    ts += [torch.full((sz,), byte, dtype=torch.uint8, device=DEV) for sz in (512 << 10, 64 << 10, 4 << 10, 512) for _ in range(64)]
    torch.cuda.synchronize()
    del ts


def step(b, byte_f, byte_b, fused=True):
    locs, feats = batches[b]
    net.fused = fused
    x = feats.clone().requires_grad_(True)
    net.zero_grad(set_to_none=True)
    poison(byte_f)
    out = net([locs, x])
    loss = out.square().sum()
    torch.cuda.synchronize()
    poison(byte_b)
    loss.backward()
    torch.cuda.synchronize()
    g = {n: p.grad.clone() for n, p in net.named_parameters()}
    return out.detach().clone(), x.grad.clone(), g


def cmp(tag, r0, r1):
    o = float((r0[0] - r1[0]).abs().max())
    d = float((r0[1] - r1[1]).abs().max())
    worst = max(((float((r0[2][n] - r1[2][n]).abs().max() / (r0[2][n].abs().max() + 1e-30)), n) for n in r0[2]),
                key=lambda t: (t[0] != t[0], t[0]))
    print(f"{tag}: out {o:.3g} (scale {float(r0[0].abs().max()):.3g})  d_feats {d:.3g} (scale {float(r0[1].abs().max()):.3g})  "
          f"worst param grad rel {worst[0]:.3g} @ {worst[1]}", flush=True)


for mode in sys.argv[1:] or ["fp32", "tf32"]:
    scn_mod.set_conv_mode(mode)
    for fused in (True, False):
        for b in (0, 1):
            base = step(b, 0x00, 0x00, fused)
            cmp(f"{mode} fused={fused} b{b} repeat      ", base, step(b, 0x00, 0x00, fused))
            cmp(f"{mode} fused={fused} b{b} fwd poison  ", base, step(b, 0xFF, 0x00, fused))
            cmp(f"{mode} fused={fused} b{b} bwd poison  ", base, step(b, 0x00, 0xFF, fused))
