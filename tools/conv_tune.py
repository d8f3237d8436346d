#!/usr/bin/env python
"""Development aid: time the tcgen05 forward kernel of several network layers under different ring geometries
(MM3D_TC_J / MM3D_TC_S / MM3D_TC_PER_SM are read per launch by conv_tc.cu in -DMM3D_TUNING builds).

    python tools/conv_tune.py [--settings "J=2;J=4,PER_SM=1"]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from mm2d3d_b200 import _lib, synth  # noqa: E402
from mm2d3d_b200 import functional as F  # noqa: E402
from mm2d3d_b200.metadata import Metadata  # noqa: E402

LAYERS = [("smc", 0, 16, 16), ("smc", 0, 32, 16), ("down", 0, 16, 32), ("up", 0, 32, 16), ("smc", 1, 32, 32), ("smc", 1, 64, 32),
          ("smc", 2, 48, 48), ("smc", 2, 96, 48), ("smc", 3, 64, 64), ("smc", 3, 128, 64), ("smc", 4, 80, 80),
          ("smc", 4, 160, 80), ("smc", 5, 96, 96), ("smc", 5, 192, 96), ("smc", 6, 112, 112), ("down", 3, 64, 80), ("up", 3, 80, 64)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--settings", default="")
    ap.add_argument("--dir", default="fwd")
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    settings = [dict(kv.split("=") for kv in s.split(",") if kv) for s in a.settings.split(";")] if a.settings else [{}]
    dev = torch.device("cuda", 0)
    locs, _ = synth.make_batch("nuscenes", batch=8)
    meta = Metadata(torch.from_numpy(locs).to(dev), 4096, 7)
    lib, m = _lib.lib, _lib.MODES["tf32"]
    flush = torch.empty(384 << 20, dtype=torch.uint8, device=dev)
    torch.manual_seed(0)
    print("layer".ljust(22) + "".join(str(s or "default").ljust(26) for s in settings))
    for kind, level, cin, cout in LAYERS:
        spatial = 4096 >> level
        if kind == "up":
            spatial //= 2
        fwd_t, bwd_t, bwd_flags = F.conv_tables(meta, kind, spatial, plans=True)
        K = fwd_t.K
        x = torch.randn(fwd_t.n_in, cin, device=dev)
        w = torch.randn(K, 1, cin, cout, device=dev) / (cin * K) ** 0.5
        out = torch.empty(fwd_t.n_out, cout, device=dev)
        dout = torch.randn(fwd_t.n_out, cout, device=dev)
        dx = torch.empty(bwd_t.n_out, cin, device=dev)
        dw = torch.empty_like(w)
        ws = F.scratch(max(lib.mm3d_conv_workspace_bytes(fwd_t.n_in, fwd_t.n_out, max(cin, cout), max(cin, cout), K, m), 1 << 22), dev)
        sp = _lib.stream_ptr()

        def run():
            if a.dir == "fwd":
                _lib.check(lib.mm3d_conv_fwd(x.data_ptr(), fwd_t.n_in, cin, out.data_ptr(), fwd_t.n_out, cout, w.data_ptr(), K, fwd_t.tbl,
                                             fwd_t.stride, fwd_t.onehot, fwd_t.plan, fwd_t.plan_cap, 0, m, ws.data_ptr(), ws.numel(), sp))
            elif a.dir == "dgrad":
                _lib.check(lib.mm3d_conv_fwd(dout.data_ptr(), bwd_t.n_in, cout, dx.data_ptr(), bwd_t.n_out, cin, w.data_ptr(), K, bwd_t.tbl,
                                             bwd_t.stride, bwd_t.onehot, bwd_t.plan, bwd_t.plan_cap, bwd_flags, m, ws.data_ptr(), ws.numel(), sp))
            else:
                _lib.check(lib.mm3d_conv_wgrad(x.data_ptr(), fwd_t.n_in, cin, dout.data_ptr(), fwd_t.n_out, cout, dw.data_ptr(), K, fwd_t.tbl,
                                               fwd_t.stride, fwd_t.onehot, fwd_t.plan, fwd_t.plan_cap, 0, m, ws.data_ptr(), ws.numel(), sp))

        cells = []
        for st in settings:
            for k in ("J", "S", "PER_SM", "NO_TMA", "SPLIT"):
                os.environ.pop("MM3D_TC_" + k, None)
            for k, v in st.items():
                os.environ["MM3D_TC_" + k] = v
            try:
                run()
                torch.cuda.synchronize()
                ts = []
                for _ in range(a.reps):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    run()
                    e1.record()
                    e1.synchronize()
                    ts.append(e0.elapsed_time(e1))
                cells.append(f"{1e3 * min(ts):7.1f} us")
            except Exception as ex:  # noqa: BLE001
                cells.append("ERR " + str(ex)[:18])
        err = lib.mm3d_take_device_error()
        print(f"{kind} L{level} {cin}->{cout} ({fwd_t.n_out})".ljust(22) + "".join(c.ljust(26) for c in cells) + (" DEVICE_ERROR" if err else ""))


if __name__ == "__main__":
    main()
