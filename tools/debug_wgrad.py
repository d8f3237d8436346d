import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mm2d3d_b200 import _lib
from mm2d3d_b200 import functional as F
from mm2d3d_b200.metadata import Metadata
lib = _lib.lib
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
n = 300
c = rng.integers(0, 12, (n, 3)) + 50
coords = np.concatenate([c, np.zeros((n, 1), np.int64)], 1).astype(np.int64)
meta = Metadata(torch.from_numpy(coords).to(dev), 4096, 1)
t, _, _ = F.conv_tables(meta, "smc", 4096)
N = t.n_out
cin, cout, K = 16, 16, 27
torch.manual_seed(0)
x = torch.randn(N, cin, device=dev); g = torch.randn(N, cout, device=dev)
ws = F.scratch(1 << 22, dev)
def run(mode):
    dw = torch.full((K, cin, cout), 7.0, device=dev)
    _lib.check(lib.mm3d_conv_wgrad(x.data_ptr(), N, cin, g.data_ptr(), N, cout, dw.data_ptr(), K, t.tbl, t.stride, None, 0, mode, ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
    torch.cuda.synchronize()
    return dw
ref = run(0)
got = run(1)
print("variant", os.environ.get("MM3D_WG_VARIANT"), "N", N, "ref absmax", ref.abs().max().item(), "got absmax", got.abs().max().item(), "nonzero", int((got != 0).sum()), "nan", int(torch.isnan(got).sum()),
      "relerr", ((got - ref).abs().max() / ref.abs().max()).item(), "dev_err", lib.mm3d_take_device_error())
k = 13
print("ref[13][:2,:4]", ref[k][:2, :4].tolist()); print("got[13][:2,:4]", got[k][:2, :4].tolist())
# where are the nonzeros
nz = (got != 0).nonzero()
print("first nonzeros", nz[:5].tolist())
