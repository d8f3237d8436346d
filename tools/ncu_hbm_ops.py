#!/usr/bin/env python
"""The HBM-bound ops of the path, one launch each inside a cudaProfiler range, for an ncu pass:

    ncu --profile-from-start off --clock-control none --metrics <tools/ncu_hbm_ops.py --metrics> --csv \
        --log-file gpurun_out/ncu_hbm_ops.csv python tools/ncu_hbm_ops.py
    python tools/ncu_hbm_ops.py --summarise gpurun_out/ncu_hbm_ops.csv gpurun_out/ncu_hbm_ops_expected.json > profiles/...txt

The run writes gpurun_out/ncu_hbm_ops_expected.json: the ops in launch order with their kernel names and algorithmic
bytes; --summarise joins it with ncu's per-launch rows (DRAM bytes, duration) into one table."""
import csv
import json
import re
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

METRICS = ("gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,"
           "dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,"
           "smsp__inst_executed.sum")


def run():
    import numpy as np
    import torch

    from mm2d3d_b200 import _lib, synth
    from mm2d3d_b200.lift import LiftIndices
    from mm2d3d_b200.metadata import Metadata

    lib, ptr = _lib.lib, _lib.ptr
    dev = torch.device("cuda", 0)
    locs, feats = synth.make_batch("nuscenes", batch=8)
    locs_d, feats_d = torch.from_numpy(locs).to(dev), torch.from_numpy(feats).to(dev)
    n = locs.shape[0]
    meta = Metadata(locs_d, 4096, 7, plans=True)  # warm-up (allocations, function attributes)
    counts = [lv.n for lv in meta._order]
    n0 = counts[0]
    sp = _lib.stream_ptr()
    per = [int((locs[:, 3] == b).sum()) for b in range(8)]
    li = LiftIndices(synth.make_img_indices(per, 225, 400, seed=0), dev)
    ops = []

    def op(label, kernels, alg_bytes, fn, warm=True):
        if warm:
            fn()
            torch.cuda.synchronize()
        torch.cuda.profiler.start()
        fn()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        ops.append({"label": label, "kernels": kernels, "alg_bytes": int(alg_bytes)})

    struct_bytes = 32 * n + sum(8 * c + 27 * 4 * c for c in counts) + sum(8 * c + 8 * 4 * c + 5 * c for c in counts[:-1])
    op("structure build, 7 levels (voxel hash, coarsening chain, 3^3 tables)",
       ["k_clear", "k_insert", "k_count", "k_scan_blocks", "k_assign", "k_ids_level0", "k_ids_coarsen", "k_setup", "k_nbr27"], struct_bytes,
       lambda: Metadata(locs_d, 4096, 7, plans=False))
    plan_bytes = sum(2 * 27 * 4 * c for c in counts) + sum(2 * 2 * 8 * 4 * c for c in counts[:-1])
    m2 = Metadata(locs_d, 4096, 7, plans=False)
    specs = [("smc", 4096 >> l) for l in range(7)] + [("down", 4096 >> l) for l in range(6)] + [("up", 4096 >> l) for l in range(6)]
    op("19 row plans (one launch)", ["k_build_plans"], plan_bytes, lambda: m2.build_plans(specs), warm=False)

    V = torch.empty(n0, 3, device=dev)
    op("InputLayer fwd, C=3", ["k_input_fwd"], n * (4 + 12) + 4 * n0 + 12 * n0,
       lambda: _lib.check(lib.mm3d_input_fwd(ptr(feats_d), meta.p2v_ptr, meta.npts_ptr, n, n0, 3, 4, ptr(V), sp)))
    dV, dF = torch.randn(n0, 3, device=dev), torch.empty(n, 3, device=dev)
    op("InputLayer bwd, C=3", ["k_input_bwd"], n * (4 + 12) + 4 * n0 + 12 * n0,
       lambda: _lib.check(lib.mm3d_input_bwd(ptr(dV), meta.p2v_ptr, meta.npts_ptr, n, 3, 4, ptr(dF), sp)))
    Z, O = torch.randn(n0, 16, device=dev), torch.empty(n, 16, device=dev)
    op("OutputLayer fwd, C=16", ["k_output_fwd"], 64 * n0 + 64 * n + 4 * n,
       lambda: _lib.check(lib.mm3d_output_fwd(ptr(Z), meta.p2v_ptr, n, 16, ptr(O), sp)))
    dO, dZ = torch.randn(n, 16, device=dev), torch.empty(n0, 16, device=dev)
    op("OutputLayer bwd, C=16", ["k_output_bwd"], 64 * n0 + 64 * n + 4 * n,
       lambda: _lib.check(lib.mm3d_output_bwd(ptr(dO), meta.p2v_ptr, n, n0, 16, ptr(dZ), sp)))
    for lvl, c in ((0, 32), (2, 96)):
        rows_l = counts[lvl]
        x = torch.randn(rows_l, c, device=dev)
        y, dy, dx = torch.empty_like(x), torch.randn_like(x), torch.empty_like(x)
        g, b = torch.ones(c, device=dev), torch.zeros(c, device=dev)
        rm, rv = torch.zeros(c, device=dev), torch.ones(c, device=dev)
        save, dgb = torch.empty(2, c, device=dev), torch.empty(2, c, device=dev)
        wsb = lib.mm3d_bnrelu_workspace_bytes(c)
        ws = torch.zeros(wsb, dtype=torch.uint8, device=dev)
        op(f"BatchNormReLU fwd, level {lvl} ({rows_l} rows), C={c}", ["k_bn_fwd_fused"], 3 * 4 * rows_l * c,
           lambda: _lib.check(lib.mm3d_bnrelu_fwd(ptr(x), ptr(y), rows_l, c, ptr(g), ptr(b), ptr(rm), ptr(rv), ptr(save[0]), ptr(save[1]),
                                                  1e-4, 0.9, 0.0, 1, ptr(ws), wsb, sp)))
        op(f"BatchNormReLU bwd, level {lvl} ({rows_l} rows), C={c}", ["k_bn_bwd_fused"], 5 * 4 * rows_l * c,
           lambda: _lib.check(lib.mm3d_bnrelu_bwd(ptr(x), ptr(dy), ptr(dx), rows_l, c, ptr(g), ptr(b), ptr(save[0]), ptr(save[1]),
                                                  ptr(dgb[0]), ptr(dgb[1]), 0.0, 1, ptr(ws), wsb, sp)))
    for C_, fmt, fname in ((6, torch.channels_last, "channels-last"), (6, torch.contiguous_format, "NCHW"), (64, torch.channels_last, "channels-last")):
        fmap = torch.randn(8, C_, 225, 400, device=dev).contiguous(memory_format=fmt)
        out = torch.empty(li.n, C_, device=dev)
        op(f"lift2d fwd [8,{C_},225,400] f32 {fname}", ["k_lift_fwd"], li.n * (16 + 2 * 4 * C_),
           lambda: _lib.check(lib.mm3d_lift2d_fwd(ptr(fmap), 0, 8, C_, 225, 400, *fmap.stride(), ptr(li.idx), ptr(li.offsets), li.n, ptr(out), sp)))
        dmap = torch.zeros_like(fmap)
        op(f"lift2d bwd [8,{C_},225,400] f32 {fname}", ["k_lift_bwd"], li.n * (16 + 3 * 4 * C_),
           lambda: _lib.check(lib.mm3d_lift2d_bwd(ptr(out), 0, 8, C_, 225, 400, *dmap.stride(), ptr(li.idx), ptr(li.offsets), li.n, ptr(dmap), sp)))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump({"n_points": n, "level_rows": counts, "ops": ops}, open(os.path.join(ROOT, "gpurun_out", "ncu_hbm_ops_expected.json"), "w"), indent=1)


def summarise(csv_path, exp_path):
    exp = json.load(open(exp_path))
    rows = [r for r in csv.reader(open(csv_path)) if r and r[0].isdigit() or (r and r[0] == "ID")]
    hdr = rows[0]
    ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    iu = hdr.index("Metric Unit")
    launches = {}
    order = []
    for r in rows[1:]:
        lid = int(r[0])
        if lid not in launches:
            launches[lid] = {"name": r[ik]}
            order.append(lid)
        v = float(r[iv].replace(",", "")) if r[iv] not in ("", "n/a") else 0.0
        u = r[iu]
        scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3,
                 "ns": 1e-9, "nsecond": 1e-9, "s": 1.0, "second": 1.0}.get(u, 1.0)
        launches[lid][r[im]] = v * scale
    peak = 6534.5
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    print(f"# ncu --clock-control none (cold caches per launch), batch of 8 nuScenes-shaped scans: {exp['n_points']} points, level rows {exp['level_rows']}")
    print(f"# algorithmic bytes per SURVEY.md 8(d); HBM peak = measured copy bandwidth {peak:.0f} GB/s; DRAM = dram__bytes_read + dram__bytes_write")
    print(f"# {'op':<68} {'launches':>8} {'us':>8} {'alg MB':>8} {'DRAM MB':>8} {'DRAM/alg':>8} {'alg GB/s':>9} {'%HBM':>6} {'DRAM GB/s':>9}")
    pos = 0
    for o in exp["ops"]:
        t = d = 0.0
        cnt = 0
        per_kernel = {}
        while pos < len(order):
            L = launches[order[pos]]
            mk = re.search(r"\b(k_\w+)", L["name"])
            short = mk.group(1) if mk else L["name"]
            if not any(short == k or short.startswith(k) for k in o["kernels"]):
                if cnt:
                    break
                pos += 1  # (a kernel of no listed op, e.g. a torch fill)
                continue
            dur = L.get("gpu__time_duration.sum", 0.0)
            db = L.get("dram__bytes_read.sum", 0.0) + L.get("dram__bytes_write.sum", 0.0)
            t += dur
            d += db
            cnt += 1
            pk = per_kernel.setdefault(short, [0, 0.0, 0.0])
            pk[0] += 1; pk[1] += dur; pk[2] += db
            pos += 1
        if not cnt:
            print(f"  {o['label']:<68} (no launch found)")
            continue
        a = o["alg_bytes"]
        print(f"  {o['label']:<68} {cnt:>8d} {t * 1e6:>8.1f} {a / 1e6:>8.2f} {d / 1e6:>8.2f} {d / a:>8.2f} {a / t / 1e9:>9.0f} {100 * a / t / 1e9 / peak:>6.1f} {d / t / 1e9:>9.0f}")
        if len(per_kernel) > 1:
            for k, (c_, tt, dd) in sorted(per_kernel.items(), key=lambda kv: -kv[1][1]):
                print(f"      {k:<64} {c_:>8d} {tt * 1e6:>8.1f} {'':>8} {dd / 1e6:>8.2f}")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--metrics":
        print(METRICS)
    elif len(sys.argv) > 1 and sys.argv[1] == "--summarise":
        summarise(sys.argv[2], sys.argv[3])
    else:
        run()
