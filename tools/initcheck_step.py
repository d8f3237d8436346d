"""One small fwd+bwd per convolution mode, meant to run under
   PYTORCH_NO_CUDA_MEMORY_CACHING=1 compute-sanitizer --tool initcheck python tools/initcheck_step.py [mode]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mm2d3d_b200 import synth, scn as scn_mod
from mm2d3d_b200.unet import UNetSCN

DEV = "cuda:0"
torch.manual_seed(3)
net = UNetSCN(in_channels=3, m=16, num_planes=4, full_scale=256).to(DEV)
locs, feats = synth.make_batch("nuscenes", batch=2, seed0=40)
locs[:, :3] //= 16
locs, feats = torch.from_numpy(locs).to(DEV), torch.from_numpy(feats).to(DEV)
for mode in sys.argv[1:] or ["fp32", "tf32"]:
    scn_mod.set_conv_mode(mode)
    for it in range(2):
        x = feats.clone().requires_grad_(True)
        net.zero_grad(set_to_none=True)
        out = net([locs, x])
        out.square().sum().backward()
        torch.cuda.synchronize()
        print(mode, it, float(out.abs().max()), float(x.grad.abs().max()), flush=True)
