#!/usr/bin/env python
"""Submanifold 3^3 sparse-convolution microbench sweep (BASELINE.json configs[4]): active voxels 10 k .. 2 M,
C_in/C_out 16 .. 192, fwd / dgrad / wgrad through the C ABI in the TF32 tensor-core mode.

Occupancy is surface-like (stacked synthetic LiDAR scans, 3-8 pairs per voxel), not uniform random.  Every launch is
timed on its own with CUDA events after an L2 flush; algorithmic bytes and flops follow SURVEY.md 8(d).  One channel
pair per size is also checked against the FP32 SIMT kernels (relative error of forward, dgrad and wgrad).

    python tools/conv_sweep.py [--out gpurun_out/conv_sweep.json]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from mm2d3d_b200 import _lib, synth  # noqa: E402
from mm2d3d_b200 import functional as F  # noqa: E402
from mm2d3d_b200.metadata import Metadata  # noqa: E402

SIZES = [10_000, 30_000, 100_000, 300_000, 1_000_000, 2_000_000]
PAIRS = [(16, 16), (32, 32), (64, 64), (112, 112), (32, 16), (96, 48), (192, 96)]


_SCANS = {}


def coords_with(n_voxels):
    """Unique voxel coordinates (x, y, z, batch) of stacked nuScenes-shaped scans, cut to n_voxels rows."""
    rows, b, total = [], 0, 0
    while total < n_voxels:
        if b not in _SCANS:
            c = synth.scan_coords("nuscenes", seed=b)
            _, first = np.unique(c, axis=0, return_index=True)
            _SCANS[b] = c[np.sort(first)]  # one row per voxel, in scan (first-occurrence) order
        c = _SCANS[b]
        rows.append(np.concatenate([c, np.full((len(c), 1), b, np.int64)], 1))
        total += len(c)
        b += 1
    return np.concatenate(rows, 0)[:n_voxels]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/conv_sweep.json")
    ap.add_argument("--sizes", default=",".join(map(str, SIZES)))
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    lib = _lib.lib
    peak = 6534.5
    try:
        pk = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
        peak = float(pk.get("hbm_gbs_burst", pk.get("hbm_gbs", peak)))
    except Exception:
        pass
    flush = torch.empty(384 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()
    rows = []
    for nv in [int(s) for s in a.sizes.split(",")]:
        locs = coords_with(nv)
        meta = Metadata(torch.from_numpy(locs).to(dev), 4096, 1, plans=True)
        t_tc, bt_tc, bflags = F.conv_tables(meta, "smc", 4096, plans=True)
        n = t_tc.n_out
        lv = meta.level(4096)
        tbl = meta._lv_view(lv, lv.o_nbr, torch.int32, 27 * t_tc.stride).view(27, t_tc.stride)[:, :n]
        pairs = int((tbl >= 0).sum())
        for ci, co in PAIRS:
            torch.manual_seed(ci * 1000 + co)
            x = torch.randn(n, ci, device=dev)
            w = torch.randn(27, 1, ci, co, device=dev) / (ci * 27) ** 0.5
            dout = torch.randn(n, co, device=dev)
            res = {}
            for mode in ("tf32", "fp32") if (ci, co) in ((32, 32), (192, 96)) else ("tf32",):
                m = _lib.MODES[mode]
                out, dx, dw = torch.empty(n, co, device=dev), torch.empty(n, ci, device=dev), torch.empty_like(w)
                ws = F.scratch(max(lib.mm3d_conv_workspace_bytes(n, n, ci, co, 27, m), lib.mm3d_conv_workspace_bytes(n, n, co, ci, 27, m), 1 << 22), dev)
                sp = _lib.stream_ptr()
                plan, pcap = (t_tc.plan, t_tc.plan_cap) if mode == "tf32" else (None, 0)
                bplan, bpcap = (bt_tc.plan, bt_tc.plan_cap) if mode == "tf32" else (None, 0)
                fns = {
                    "fwd": lambda: _lib.check(lib.mm3d_conv_fwd(x.data_ptr(), n, ci, out.data_ptr(), n, co, w.data_ptr(), 27, t_tc.tbl, t_tc.stride, None, plan, pcap, 0, m, ws.data_ptr(), ws.numel(), sp)),
                    "dgrad": lambda: _lib.check(lib.mm3d_conv_fwd(dout.data_ptr(), n, co, dx.data_ptr(), n, ci, w.data_ptr(), 27, bt_tc.tbl, bt_tc.stride, None, bplan, bpcap, bflags, m, ws.data_ptr(), ws.numel(), sp)),
                    "wgrad": lambda: _lib.check(lib.mm3d_conv_wgrad(x.data_ptr(), n, ci, dout.data_ptr(), n, co, dw.data_ptr(), 27, t_tc.tbl, t_tc.stride, None, plan, pcap, 0, m, ws.data_ptr(), ws.numel(), sp)),
                }
                for d, fn in fns.items():
                    fn()
                    torch.cuda.synchronize()
                    if mode == "fp32":
                        continue  # only its results are needed
                    ts = []
                    for _ in range(a.reps):
                        flush.zero_()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record(stream)
                        fn()
                        e1.record(stream)
                        e1.synchronize()
                        ts.append(e0.elapsed_time(e1))
                    ms = min(ts)
                    alg = 4 * (n * ci + n * co) + 4 * 27 * ci * co + 4 * 27 * n
                    flops = 2 * pairs * ci * co
                    rows.append({"voxels": n, "pairs": pairs, "c_in": ci, "c_out": co, "dir": d, "us": ms * 1e3,
                                 "alg_GBps": alg / ms / 1e6, "frac_hbm": alg / ms / 1e6 / peak, "TFLOPs": flops / ms / 1e9})
                res[mode] = (out.clone(), dx.clone(), dw.clone())
            if "fp32" in res:
                errs = [float((p - q).norm() / q.norm()) for p, q in zip(res["tf32"], res["fp32"])]
                rows.append({"voxels": n, "c_in": ci, "c_out": co, "dir": "parity_vs_fp32_kernels",
                             "rel_err_fwd_dgrad_wgrad": errs, "ok": all(e < 1e-2 for e in errs)})
        if lib.mm3d_take_device_error():
            raise SystemExit("device error flag set")
        del meta
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump({"hbm_peak_GBps": peak, "rows": rows}, open(a.out, "w"), indent=0)
    print(f"# submanifold 3^3 sweep, TF32 tensor-core mode, min of {a.reps} L2-flushed launches; HBM peak {peak:.0f} GB/s")
    print(f"# {'voxels':>8} {'pairs/vox':>9} {'cin->cout':>9} | " + " | ".join(f"{d:>5} us   GB/s  %HBM  TF/s" for d in ("fwd", "dgrad", "wgrad")))
    by = {}
    for r in rows:
        if r["dir"].startswith("parity"):
            print(f"# parity {r['voxels']} voxels {r['c_in']}->{r['c_out']}: rel err fwd/dgrad/wgrad vs FP32 kernels = "
                  + "/".join(f"{e:.1e}" for e in r["rel_err_fwd_dgrad_wgrad"]) + (" ok" if r["ok"] else " FAIL"))
            continue
        by.setdefault((r["voxels"], r["pairs"], r["c_in"], r["c_out"]), {})[r["dir"]] = r
    for (nv, pr, ci, co), d in by.items():
        print(f"  {nv:>8} {pr / nv:>9.2f} {ci:>4}->{co:<4} | " + " | ".join(
            f"{d[k]['us']:>8.1f} {d[k]['alg_GBps']:>6.0f} {100 * d[k]['frac_hbm']:>5.1f} {d[k]['TFLOPs']:>5.1f}" for k in ("fwd", "dgrad", "wgrad")))


if __name__ == "__main__":
    main()
