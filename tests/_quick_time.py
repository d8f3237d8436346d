"""Ad-hoc timing of the module path (not the bench): prints ms for fwd and fwd+bwd."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mm2d3d_b200 import synth
from mm2d3d_b200.unet import UNetSCN

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
locs, feats = synth.make_batch("nuscenes", batch=batch)
net = UNetSCN(3).cuda()
locs_d = torch.from_numpy(locs).cuda(); feats_d = torch.from_numpy(feats).cuda()
def step(bwd=True):
    x = feats_d.clone().requires_grad_(True)
    out = net([locs_d, x])
    if bwd: out.sum().backward()
for _ in range(3): step()
torch.cuda.synchronize()
for bwd in (False, True):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time(); e0.record()
    for _ in range(5): step(bwd)
    e1.record(); torch.cuda.synchronize()
    print("batch", batch, "points", locs.shape[0], "bwd" if bwd else "fwd", "gpu ms/step", e0.elapsed_time(e1)/5, "wall ms/step", (time.time()-t0)*200)
