#!/usr/bin/env python
"""Development aid (library built with MM3D_EXTRA_NVCC_FLAGS=-DMM3D_TRACE): per-role clock64 stamps of CTA 0 of one
tcgen05 forward launch -- when each producer warp waited for / got / filled its stage, when the MMA warp waited for /
got / issued each item, when the epilogue got and finished each tile."""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mm2d3d_b200 import _lib, synth  # noqa: E402
from mm2d3d_b200 import functional as F  # noqa: E402
from mm2d3d_b200.metadata import Metadata  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", default="smc")
    ap.add_argument("--level", type=int, default=0)
    ap.add_argument("--cin", type=int, default=16)
    ap.add_argument("--cout", type=int, default=16)
    ap.add_argument("--warm", action="store_true")
    ap.add_argument("--dir", default="fwd", choices=["fwd", "wgrad"])
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    locs, _ = synth.make_batch("nuscenes", batch=8)
    meta = Metadata(torch.from_numpy(locs).to(dev), 4096, 7)
    spatial = 4096 >> a.level
    t, _, _ = F.conv_tables(meta, a.kind, spatial, plans=True)
    lib, m = _lib.lib, _lib.MODES["tf32"]
    x = torch.randn(t.n_in, a.cin, device=dev)
    w = torch.randn(t.K, 1, a.cin, a.cout, device=dev) / (a.cin * t.K) ** 0.5
    out = torch.empty(t.n_out, a.cout, device=dev)
    ws = F.scratch(max(lib.mm3d_conv_workspace_bytes(t.n_in, t.n_out, a.cin, a.cout, t.K, m), 1 << 22), dev)
    trace = torch.zeros(16 * 64 * 4 + 1024, dtype=torch.int64, device=dev)
    flush = torch.empty(384 << 20, dtype=torch.uint8, device=dev)

    dout = torch.randn(t.n_out, a.cout, device=dev)
    dw = torch.empty_like(w)

    def run():
        if a.dir == "wgrad":
            _lib.check(lib.mm3d_conv_wgrad(x.data_ptr(), t.n_in, a.cin, dout.data_ptr(), t.n_out, a.cout, dw.data_ptr(), t.K, t.tbl,
                                           t.stride, t.onehot, t.plan, t.plan_cap, 0, m, ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
            return
        _lib.check(lib.mm3d_conv_fwd(x.data_ptr(), t.n_in, a.cin, out.data_ptr(), t.n_out, a.cout, w.data_ptr(), t.K, t.tbl, t.stride,
                                     t.onehot, t.plan, t.plan_cap, 0, m, ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
    run()
    torch.cuda.synchronize()
    if not a.warm:
        flush.zero_()
    lib.mm3d_debug_set_trace.argtypes = [C.c_void_p]
    lib.mm3d_debug_set_trace(trace.data_ptr())
    run()
    torch.cuda.synchronize()
    lib.mm3d_debug_set_trace(None)
    ctas = trace[4096:].view(-1, 2).cpu().numpy()
    tr = trace[:4096].view(16, 64, 4).cpu().numpy()
    if (ctas > 0).any():  # forward kernel: globaltimer at each CTA's start (after the dependency wait) and end, ns
        live = ctas[ctas[:, 0] > 0]
        g0 = live[:, 0].min()
        print("CTA start/end (ns after the first start):")
        for c, (b, e) in enumerate(live):
            print(f"   cta {c:3d} {b - g0:7d} {e - g0:7d}")
    t0 = tr[tr > 0].min()
    rel = lambda v: "      -" if v == 0 else f"{(v - t0):7d}"
    print("columns: producers (roles 0..7): wait_start got_stage filled -- epilogue (role 9): wait_start got_tile done -- "
          "MMA issuers (roles 10, 11): wait_start got_item issued acc_wait_start\n"
          "wgrad: producers (0..7): item_start a_wait_start got_stage filled -- MMA issuer (role 9): wait_start got_item issued -- "
          "epilogue (role 10): flush_start flush_end")
    for role in range(16):
        if not (tr[role] > 0).any():
            continue
        print(f"role {role}")
        for r in range(64):
            if (tr[role, r] > 0).any():
                print(f"   {r:3d} " + " ".join(rel(v) for v in tr[role, r]))


if __name__ == "__main__":
    main()
