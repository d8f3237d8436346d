"""cProfile of the Python/launch side of one training step (module path)."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mm2d3d_b200 import synth, scn
from mm2d3d_b200.unet import UNetSCN
mode = sys.argv[1] if len(sys.argv) > 1 else "tf32"
scn.set_conv_mode(mode)
locs, feats = synth.make_batch("nuscenes", batch=8)
net = UNetSCN(3).cuda()
locs_d = torch.from_numpy(locs).cuda(); feats_d = torch.from_numpy(feats).cuda()
g = torch.randn(locs.shape[0], 16, device="cuda")
def step():
    x = feats_d.detach().requires_grad_(True)
    out = net([locs_d, x]); out.backward(g)
for _ in range(3): step()
torch.cuda.synchronize()
# CPU-side time with the GPU kept out of the way: tiny input => kernels are negligible
small = torch.from_numpy(locs[:2000]).cuda(); sf = torch.from_numpy(feats[:2000]).cuda(); sg = g[:2000].clone()
def small_step():
    x = sf.detach().requires_grad_(True)
    out = net([small, x]); out.backward(sg)
for _ in range(3): small_step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): small_step()
torch.cuda.synchronize()
print("tiny-input step (CPU/launch bound) ms:", (time.perf_counter() - t0) * 50)
pr = cProfile.Profile(); pr.enable()
for _ in range(10): small_step()
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
