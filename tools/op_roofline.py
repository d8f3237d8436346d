#!/usr/bin/env python
"""Achieved HBM bandwidth of the memory-bound ops of the path (structure build, Input/OutputLayer, BatchNorm+ReLU,
2D->3D lift, point rasteriser), each timed alone through the C ABI after an L2 flush, on the bench's batch
(8 nuScenes-shaped scans).  Algorithmic bytes follow SURVEY.md 8(d).

    python tools/op_roofline.py > profiles/<name>.txt
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from mm2d3d_b200 import _lib, synth  # noqa: E402
from mm2d3d_b200.metadata import Metadata  # noqa: E402

lib, ptr = _lib.lib, _lib.ptr
DEV = torch.device("cuda", 0)


def main():
    peak = 6534.5
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    flush = torch.empty(384 << 20, dtype=torch.uint8, device=DEV)
    stream = torch.cuda.current_stream()
    sp = _lib.stream_ptr()
    rows = []

    def timed(name, alg_bytes, fn, reps=5, note=""):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = min(ts)
        rows.append((name, alg_bytes, ms * 1e3, alg_bytes / ms / 1e6, note))

    locs, feats = synth.make_batch("nuscenes", batch=8)
    locs_d, feats_d = torch.from_numpy(locs).to(DEV), torch.from_numpy(feats).to(DEV)
    n = locs.shape[0]
    meta = Metadata(locs_d, 4096, 7, plans=True)
    counts = [lv.n for lv in meta._order]
    n0 = counts[0]

    # ---- structure: the whole 7-level build (voxel hash, coarsening chain, 3^3 tables) and the 19 row plans
    struct_bytes = 32 * n + sum(8 * c + 27 * 4 * c for c in counts) + sum(8 * c + 8 * 4 * c + 5 * c for c in counts[:-1])
    timed("structure build, 7 levels (57 launches, 1 host sync)", struct_bytes, lambda: Metadata(locs_d, 4096, 7, plans=False),
          note="latency-bound chain of small kernels")
    plan_bytes = sum(2 * 27 * 4 * c for c in counts) + sum(2 * 2 * 8 * 4 * c for c in counts[:-1])
    timed("  + 19 row plans (1 more launch)", struct_bytes + plan_bytes, lambda: Metadata(locs_d, 4096, 7, plans=True))

    # ---- Input / Output layers (C = 3 in, 16 out)
    V = torch.empty(n0, 3, device=DEV)
    timed("InputLayer fwd (mode 4 mean), C=3", n * (4 + 12) + 4 * n0 + 12 * n0,
          lambda: _lib.check(lib.mm3d_input_fwd(ptr(feats_d), meta.p2v_ptr, meta.npts_ptr, n, n0, 3, 4, ptr(V), sp)))
    dV, dF = torch.randn(n0, 3, device=DEV), torch.empty(n, 3, device=DEV)
    timed("InputLayer bwd, C=3", n * (4 + 12) + 4 * n0 + 12 * n0,
          lambda: _lib.check(lib.mm3d_input_bwd(ptr(dV), meta.p2v_ptr, meta.npts_ptr, n, 3, 4, ptr(dF), sp)))
    Z, O = torch.randn(n0, 16, device=DEV), torch.empty(n, 16, device=DEV)
    timed("OutputLayer fwd, C=16", 64 * n0 + 64 * n + 4 * n,
          lambda: _lib.check(lib.mm3d_output_fwd(ptr(Z), meta.p2v_ptr, n, 16, ptr(O), sp)))
    dO, dZ = torch.randn(n, 16, device=DEV), torch.empty(n0, 16, device=DEV)
    timed("OutputLayer bwd, C=16", 64 * n0 + 64 * n + 4 * n,
          lambda: _lib.check(lib.mm3d_output_bwd(ptr(dO), meta.p2v_ptr, n, n0, 16, ptr(dZ), sp)))

    # the same kernels on 8x the points (2.2 M points): at the bench's size (8 - 34 MB) these launches are bounded by
    # launch latency (~4 us at 6.5 TB/s would already be 100 %), the large size shows what the kernels themselves reach
    rep = 8
    p2v_big = torch.cat([meta.p2v().int() + r * n0 for r in range(rep)]).contiguous()
    npts_big = meta.npts().int().repeat(rep).contiguous()
    nb_, n0b = n * rep, n0 * rep
    feats_b, Vb = feats_d.repeat(rep, 1).contiguous(), torch.empty(n0 * rep, 3, device=DEV)
    timed(f"InputLayer fwd, C=3, {nb_} points", nb_ * (4 + 12) + 4 * n0b + 12 * n0b,
          lambda: _lib.check(lib.mm3d_input_fwd(ptr(feats_b), ptr(p2v_big), ptr(npts_big), nb_, n0b, 3, 4, ptr(Vb), sp)))
    Zb, Ob = torch.randn(n0b, 16, device=DEV), torch.empty(nb_, 16, device=DEV)
    timed(f"OutputLayer fwd, C=16, {nb_} points", 64 * n0b + 64 * nb_ + 4 * nb_,
          lambda: _lib.check(lib.mm3d_output_fwd(ptr(Zb), ptr(p2v_big), nb_, 16, ptr(Ob), sp)))
    dOb, dZb = torch.randn(nb_, 16, device=DEV), torch.empty(n0b, 16, device=DEV)
    timed(f"OutputLayer bwd, C=16, {nb_} points", 64 * n0b + 64 * nb_ + 4 * nb_,
          lambda: _lib.check(lib.mm3d_output_bwd(ptr(dOb), ptr(p2v_big), nb_, n0b, 16, ptr(dZb), sp)))
    del feats_b, Vb, Zb, Ob, dOb, dZb

    # ---- BatchNorm + ReLU, training, at the sizes of levels 0, 2 and 5
    for lvl, c in ((0, 16), (0, 32), (2, 48), (2, 96), (5, 96)):
        rows_l = counts[lvl]
        x = torch.randn(rows_l, c, device=DEV)
        y, dy, dx = torch.empty_like(x), torch.randn_like(x), torch.empty_like(x)
        g, b = torch.ones(c, device=DEV), torch.zeros(c, device=DEV)
        rm, rv = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
        save, dgb = torch.empty(2, c, device=DEV), torch.empty(2, c, device=DEV)
        wsb = lib.mm3d_bnrelu_workspace_bytes(c)
        ws = torch.zeros(wsb, dtype=torch.uint8, device=DEV)
        timed(f"BatchNormReLU fwd, level {lvl} ({rows_l} rows), C={c}", 3 * 4 * rows_l * c,
              lambda: _lib.check(lib.mm3d_bnrelu_fwd(ptr(x), ptr(y), rows_l, c, ptr(g), ptr(b), ptr(rm), ptr(rv), ptr(save[0]),
                                                     ptr(save[1]), 1e-4, 0.9, 0.0, 1, ptr(ws), wsb, sp)),
              note="one launch: stats, grid barrier, apply (2nd read from L2)")
        timed(f"BatchNormReLU bwd, level {lvl} ({rows_l} rows), C={c}", 5 * 4 * rows_l * c,
              lambda: _lib.check(lib.mm3d_bnrelu_bwd(ptr(x), ptr(dy), ptr(dx), rows_l, c, ptr(g), ptr(b), ptr(save[0]), ptr(save[1]),
                                                     ptr(dgb[0]), ptr(dgb[1]), 0.0, 1, ptr(ws), wsb, sp)))

    # ---- 2D -> 3D lift and the point rasteriser
    from mm2d3d_b200.lift import LiftIndices, rasterize_points
    per = [int((locs[:, 3] == b).sum()) for b in range(8)]
    li = LiftIndices(synth.make_img_indices(per, 225, 400, seed=0), DEV)
    for C_, dt, name in ((6, torch.float32, "f32"), (6, torch.float16, "f16"), (64, torch.float32, "f32")):
        for fmt, fname in ((torch.contiguous_format, "NCHW"), (torch.channels_last, "channels-last")):
            fmap = torch.randn(8, C_, 225, 400, device=DEV).to(dt).contiguous(memory_format=fmt)
            out = torch.empty(li.n, C_, dtype=dt, device=DEV)
            es = fmap.element_size()
            code = {torch.float32: 0, torch.float16: 1}[dt]
            note = ("single elements from a channel-major map: sector-bound (32 B moved per 4 B used)" if fname == "NCHW"
                    else "a pixel's channels are one contiguous piece")
            timed(f"lift2d fwd [8,{C_},225,400] {name} {fname}", li.n * (16 + 2 * es * C_),
                  lambda: _lib.check(lib.mm3d_lift2d_fwd(ptr(fmap), code, 8, C_, 225, 400, *fmap.stride(), ptr(li.idx), ptr(li.offsets),
                                                         li.n, ptr(out), sp)), note=note)
            dmap = torch.zeros_like(fmap)
            timed(f"lift2d bwd [8,{C_},225,400] {name} {fname}", li.n * (16 + 3 * es * C_),
                  lambda: _lib.check(lib.mm3d_lift2d_bwd(ptr(out), code, 8, C_, 225, 400, *dmap.stride(), ptr(li.idx), ptr(li.offsets),
                                                         li.n, ptr(dmap), sp)), note="atomics")
    vals = torch.rand(li.n, device=DEV)
    timed("rasterize_points [8,225,400]", li.n * (16 + 4) + 2 * 4 * 8 * 225 * 400, lambda: rasterize_points(li, vals, 225, 400, 0.0),
          note="3 launches + allocation")

    print(f"# memory-bound ops alone (L2 flushed, min of 5), batch of 8 nuScenes-shaped scans: {n} points, level rows {counts}")
    print(f"# HBM peak (measured copy bandwidth) {peak:.0f} GB/s")
    print(f"# {'op':<62} {'alg MB':>8} {'us':>8} {'GB/s':>7} {'%HBM':>6}  note")
    for name, b, us, gbs, note in rows:
        print(f"  {name:<62} {b / 1e6:>8.2f} {us:>8.1f} {gbs:>7.0f} {100 * gbs / peak:>6.1f}  {note}")


if __name__ == "__main__":
    main()
