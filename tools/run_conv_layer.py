#!/usr/bin/env python
"""Run one sparse-convolution launch configuration repeatedly (for ncu captures / quick timing).

    python tools/run_conv_layer.py --kind smc --level 0 --cin 16 --cout 16 --mode tf32 --dir fwd --reps 5
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from mm2d3d_b200 import _lib, synth  # noqa: E402
from mm2d3d_b200 import functional as F  # noqa: E402
from mm2d3d_b200.metadata import Metadata  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", default="smc", choices=["smc", "down", "up"])
    ap.add_argument("--level", type=int, default=0)
    ap.add_argument("--cin", type=int, default=16)
    ap.add_argument("--cout", type=int, default=16)
    ap.add_argument("--mode", default="tf32")
    ap.add_argument("--dir", default="fwd", choices=["fwd", "dgrad", "wgrad"])
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--shape", default="nuscenes")
    ap.add_argument("--noflush", action="store_true")
    a = ap.parse_args()

    dev = torch.device("cuda", 0)
    locs, _ = synth.make_batch(a.shape, batch=a.batch)
    meta = Metadata(torch.from_numpy(locs).to(dev), 4096, 7)
    spatial = 4096 >> a.level
    if a.kind == "up":
        spatial //= 2  # input lives one level deeper
    fwd_t, bwd_t, bwd_flags = F.conv_tables(meta, a.kind, spatial, plans=a.mode != "fp32")
    K = fwd_t.K
    lib = _lib.lib
    m = _lib.MODES[a.mode]
    torch.manual_seed(0)
    x = torch.randn(fwd_t.n_in, a.cin, device=dev)
    w = torch.randn(K, 1, a.cin, a.cout, device=dev) / (a.cin * K) ** 0.5
    out = torch.empty(fwd_t.n_out, a.cout, device=dev)
    dout = torch.randn(fwd_t.n_out, a.cout, device=dev)
    dx = torch.empty(bwd_t.n_out, a.cin, device=dev)
    dw = torch.empty_like(w)
    ws = F.scratch(max(lib.mm3d_conv_workspace_bytes(fwd_t.n_in, fwd_t.n_out, a.cin, a.cout, K, m), 1 << 22), dev)
    sp = _lib.stream_ptr()

    def run():
        if a.dir == "fwd":
            _lib.check(lib.mm3d_conv_fwd(x.data_ptr(), fwd_t.n_in, a.cin, out.data_ptr(), fwd_t.n_out, a.cout,
                                         w.data_ptr(), K, fwd_t.tbl, fwd_t.stride, fwd_t.onehot, fwd_t.plan, fwd_t.plan_cap, 0, m,
                                         ws.data_ptr(), ws.numel(), sp))
        elif a.dir == "dgrad":
            _lib.check(lib.mm3d_conv_fwd(dout.data_ptr(), bwd_t.n_in, a.cout, dx.data_ptr(), bwd_t.n_out, a.cin,
                                         w.data_ptr(), K, bwd_t.tbl, bwd_t.stride, bwd_t.onehot, bwd_t.plan, bwd_t.plan_cap, bwd_flags, m,
                                         ws.data_ptr(), ws.numel(), sp))
        else:
            _lib.check(lib.mm3d_conv_wgrad(x.data_ptr(), fwd_t.n_in, a.cin, dout.data_ptr(), fwd_t.n_out, a.cout,
                                           dw.data_ptr(), K, fwd_t.tbl, fwd_t.stride, fwd_t.onehot, fwd_t.plan, fwd_t.plan_cap, 0, m,
                                           ws.data_ptr(), ws.numel(), sp))

    flush = torch.empty(384 << 20, dtype=torch.uint8, device=dev)
    run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(a.reps):
        if not a.noflush:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    err = lib.mm3d_take_device_error()
    print(f"{a.kind} {a.dir} L{a.level} {a.cin}->{a.cout} mode {a.mode}: rows in/out {fwd_t.n_in}/{fwd_t.n_out} "
          f"ms {sorted(ts)[len(ts) // 2]:.4f} (min {min(ts):.4f}) device_error {err}")


if __name__ == "__main__":
    main()
