import os, sys
sys.path.insert(0, "/root/repo")
import torch
from mm2d3d_b200 import _lib, synth
from mm2d3d_b200 import functional as F
from mm2d3d_b200.metadata import Metadata
dev = torch.device("cuda", 0)
locs, _ = synth.make_batch("nuscenes", batch=2)
meta = Metadata(torch.from_numpy(locs).to(dev), 4096, 2)
for ci, co in ((16, 16), (32, 16), (64, 32), (96, 48), (128, 64), (112, 112)):
    torch.manual_seed(0)
    t, _, _ = F.conv_tables(meta, "smc", 4096, plans=True)
    x = torch.randn(t.n_in, ci, device=dev)
    dout = torch.randn(t.n_out, co, device=dev)
    lib, m = _lib.lib, _lib.MODES["bf16"]
    xp, dp = F._planes(x, True), F._planes(dout, True)
    res = {}
    for tag, env in (("bf16", None), ("tf32-fallback", "1")):
        if env: os.environ["MM3D_WGRAD_NO_BF16"] = env
        else: os.environ.pop("MM3D_WGRAD_NO_BF16", None)
        dw = torch.empty(27, 1, ci, co, device=dev)
        _lib.check(lib.mm3d_conv_wgrad(xp.data_ptr(), t.n_in, ci, dp.data_ptr(), t.n_out, co, dw.data_ptr(), 27, t.tbl, t.stride, None,
                                       t.plan, t.plan_cap, 0, m, None, 0, _lib.stream_ptr()))
        torch.cuda.synchronize()
        res[tag] = dw
    # fp32 SIMT reference
    dwr = torch.empty(27, 1, ci, co, device=dev)
    _lib.check(lib.mm3d_conv_wgrad(x.data_ptr(), t.n_in, ci, dout.data_ptr(), t.n_out, co, dwr.data_ptr(), 27, t.tbl, t.stride, None,
                                   None, 0, 0, 0, None, 0, _lib.stream_ptr()))
    torch.cuda.synchronize()
    sc = float(dwr.abs().max())
    print(ci, co, {k: f"{float((v - dwr).abs().max()) / sc:.2e}" for k, v in res.items()}, "device_error", lib.mm3d_take_device_error())
