"""Fused execution of ``UNetSCN``: the whole forward (and the whole backward) is ONE call into
``libmm3d`` (``csrc/unet_exec.cu``) instead of ~110 Python autograd nodes.

``UNetSCN.forward`` uses this path when the network has the reference's shape (VGG blocks,
``block_reps == 1``, InputLayer mode 4 -- ``config/config.yaml:22-28``); any other configuration
runs module by module.  Results are the same kernels in the same order, so parity is unchanged.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from . import functional as F
from ._lib import check, lib
from .metadata import Metadata


def collect_slots(net):
    """Parameters and BN buffers of ``net`` in the executor's slot order (module-tree order):
    stem.w | per level: pre_bn{w,b,rm,rv} pre.w [dn_bn dn.w <deeper> up_bn up.w post_bn post.w] | head_bn."""

    def bn(mod):
        return [mod.weight, mod.bias, mod.running_mean, mod.running_var]

    def level(seq):
        mods = list(seq._modules.values())
        out = bn(mods[0][0]) + [mods[0][1].weight]
        if len(mods) > 1:
            branch = list(mods[1][1]._modules.values())  # ConcatTable[Identity, Sequential[BN, Conv, U, BN, Deconv]]
            out += bn(branch[0]) + [branch[1].weight]
            out += level(branch[2])
            out += bn(branch[3]) + [branch[4].weight]
            out += bn(mods[3][0]) + [mods[3][1].weight]
        return out

    cached = getattr(net, "_exec_slots", None)
    if cached is not None and cached[0] is net.layer2.weight and cached[-1] is net.layer4.running_var:
        return cached  # (the module walk costs ~0.1 ms per call; parameters are updated in place by optimisers / .to())
    slots = [net.layer2.weight] + level(net.layer3) + bn(net.layer4)
    net._exec_slots = slots
    return slots


def fusable(net) -> bool:
    """Whether ``net`` has the shape the executor implements.  The module walk costs ~0.35 ms, so the answer is
    cached on the network together with what it depends on (per-layer conv modes, leakiness, InputLayer mode)."""
    from . import scn
    mods = getattr(net, "_exec_mods", None)
    if mods is None:
        mods = net._exec_mods = [m for m in net.modules() if isinstance(m, (scn._ConvBase, scn.BatchNormLeakyReLU))]
    key = (net.layer1.mode if isinstance(net.layer1, scn.InputLayer) else None,
           tuple((m.mode, hasattr(m, "bias")) if isinstance(m, scn._ConvBase) else m.leakiness for m in mods))
    cached = getattr(net, "_exec_fusable", None)
    if cached is not None and cached[0] == key:
        return cached[1]
    ok = _fusable(net)
    net._exec_fusable = (key, ok)
    return ok


def _fusable(net) -> bool:
    from . import scn
    if not isinstance(net.layer1, scn.InputLayer) or net.layer1.mode != 4:
        return False
    if getattr(net, "_block_reps", 1) != 1 or getattr(net, "_residual", False):
        return False
    for mod in net.modules():
        if isinstance(mod, scn._ConvBase) and (mod.mode is not None or hasattr(mod, "bias")):
            return False
        if isinstance(mod, scn.BatchNormLeakyReLU) and mod.leakiness != 0:
            return False
    return True


def _level_desc(meta: Metadata, num_planes: int, spatial0: int, plans: bool):
    W = _lib.LEVEL_DESC_WORDS
    desc = (C.c_int64 * (W * num_planes))()
    if plans:  # every table of the forward in one launch
        specs, s = [], spatial0
        for l in range(num_planes):
            specs.append(("smc", s))
            if l + 1 < num_planes:
                specs += [("down", s), ("up", s)]
            s //= 2
        meta.build_plans(specs)
    s = spatial0
    for l in range(num_planes):
        lv = meta.nbr(s)
        desc[W * l + 0] = lv.n
        desc[W * l + 1] = lv.ptr(lv.o_nbr)
        desc[W * l + 2] = lv.tstride
        if plans:
            desc[W * l + 6], desc[W * l + 9] = meta.plan("smc", s)
        if l + 1 < num_planes:
            fine, _ = meta.down(s)
            desc[W * l + 3] = fine.ptr(fine.o_parent)
            desc[W * l + 4] = fine.ptr(fine.o_off)
            desc[W * l + 5] = fine.ptr(fine.o_child)
            if plans:
                desc[W * l + 7] = meta.plan("down", s)[0]
                desc[W * l + 8] = meta.plan("up", s)[0]
        s //= 2
    return desc


class UNetSCNFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, meta, cfg, mask_w, mask_b, *slots):
        in_ch, m, L, mode, training, eps, momentum, spatial0 = cfg
        feats = feats.float().contiguous()
        dev = feats.device
        # optional RGB mask of Net3DSeg.forward (3d_net/model.py:46-48), folded into the InputLayer scatter
        masked = mask_w is not None
        if masked and (mask_w.numel() != in_ch or mask_b is None or mask_b.numel() != 1):
            raise ValueError("UNetSCN: rgb_mask must be the weight [1, C] and bias [1] of an nn.Linear(C, 1)")
        n_points = meta.n_points
        desc = _level_desc(meta, L, spatial0, plans=mode != _lib.MODE_FP32)
        with torch.cuda.device(dev):
            # sticky device-error words (mapped host memory): whatever a kernel of an earlier call raised surfaces
            # here -- no synchronisation, a few nanoseconds
            _lib.raise_device_errors("UNetSCN.forward")
            act_bytes = lib.mm3d_unet_act_bytes(in_ch, m, L, mode, desc, n_points)
            act = torch.empty(act_bytes, dtype=torch.uint8, device=dev)
            scr = F.scratch(lib.mm3d_unet_scratch_bytes(in_ch, m, L, mode), dev)
            out = torch.empty(n_points, m, dtype=torch.float32, device=dev)
            params = (C.c_void_p * len(slots))(*[t.data_ptr() for t in slots])
            mask, wb, gate = None, None, None
            if masked:
                wb = torch.cat([mask_w.detach().reshape(-1), mask_b.detach().reshape(-1)]).float().contiguous()
                gate = torch.empty(max(n_points, 1), dtype=torch.float32, device=dev)
                mask = C.byref(_lib.UnetMask(wb.data_ptr(), gate.data_ptr(), None, None, None, 0))
            check(lib.mm3d_unet_forward(in_ch, m, L, mode, int(training), eps, momentum, desc, n_points, meta.p2v_ptr,
                                        meta.npts_ptr, feats.data_ptr(), out.data_ptr(), params, act.data_ptr(), act_bytes,
                                        scr.data_ptr(), scr.numel(), mask, _lib.stream_ptr()), "mm3d_unet_forward")
        ctx.meta, ctx.cfg, ctx.desc, ctx.act, ctx.slots = meta, cfg, desc, act, slots
        ctx.feats_shape = feats.shape
        ctx.mask = (feats, wb, gate, mask_w.shape, mask_b.shape) if masked else None
        return out

    @staticmethod
    def backward(ctx, d_out):
        in_ch, m, L, mode, training, eps, momentum, spatial0 = ctx.cfg
        meta, slots, desc, act = ctx.meta, ctx.slots, ctx.desc, ctx.act
        d_out = d_out.float().contiguous()
        dev = d_out.device
        n_points = meta.n_points
        need_feats = ctx.needs_input_grad[0]
        with torch.cuda.device(dev):
            _lib.raise_device_errors("UNetSCN.backward")
            # one flat buffer for every parameter gradient; the returned grads are views into it
            live = [t for t in slots if t.requires_grad]
            flat = torch.empty(sum(t.numel() for t in live), dtype=torch.float32, device=dev)
            # all views in one native call (what DDP's buckets use) instead of ~130 slice + view calls
            views = iter(torch._utils._unflatten_dense_tensors(flat, live)) if live else iter(())
            gptrs, grads = [], []
            for t in slots:
                if t.requires_grad:
                    g = next(views)
                    gptrs.append(g.data_ptr())
                    grads.append(g)
                else:
                    gptrs.append(None)
                    grads.append(None)
            d_feats = torch.empty(ctx.feats_shape, dtype=torch.float32, device=dev) if need_feats else None
            if need_feats and ctx.feats_shape[0] > n_points:
                d_feats.zero_()
            tmp_bytes = lib.mm3d_unet_bwd_bytes(in_ch, m, L, mode, desc, n_points)
            tmp = torch.empty(tmp_bytes, dtype=torch.uint8, device=dev)
            scr = F.scratch(lib.mm3d_unet_scratch_bytes(in_ch, m, L, mode), dev)
            params = (C.c_void_p * len(slots))(*[t.data_ptr() for t in slots])
            gp = (C.c_void_p * len(slots))(*gptrs)
            mask, d_mw, d_mb = None, None, None
            if ctx.mask is not None:
                feats, wb, gate, w_shape, b_shape = ctx.mask
                d_wb = torch.empty(in_ch + 1, dtype=torch.float32, device=dev)
                mws = torch.empty(in_ch + 1, dtype=torch.float64, device=dev)
                mask = C.byref(_lib.UnetMask(wb.data_ptr(), gate.data_ptr(), feats.data_ptr(), d_wb.data_ptr(), mws.data_ptr(),
                                             mws.numel() * 8))
                d_mw, d_mb = d_wb[:in_ch].reshape(w_shape), d_wb[in_ch:].reshape(b_shape)
            check(lib.mm3d_unet_backward(in_ch, m, L, mode, int(training), desc, n_points, meta.p2v_ptr, meta.npts_ptr,
                                         d_out.data_ptr(), d_feats.data_ptr() if need_feats else None, params, gp,
                                         act.data_ptr(), act.numel(), tmp.data_ptr(), tmp_bytes, scr.data_ptr(),
                                         scr.numel(), mask, _lib.stream_ptr()), "mm3d_unet_backward")
        return (d_feats, None, None, d_mw, d_mb, *grads)


class PreparedScans:
    """Sparse structure of one batch (voxels, rule tables, row plans), built ahead of the forward that uses it --
    typically on a side stream while the previous step still computes, the way a prefetching data loader prepares
    the next batch.  Pass it to ``UNetSCN.forward`` in place of the coordinate tensor."""

    __slots__ = ("meta", "event", "stream", "mode", "spatial0", "levels", "points")

    def __init__(self, meta, mode, spatial0, levels, points=None):
        self.meta, self.mode, self.spatial0, self.levels = meta, mode, spatial0, levels
        self.points = points  # augment.PointStructure when the batch was prepared from raw float points
        self.stream = torch.cuda.current_stream(meta.device)
        self.event = torch.cuda.Event()
        self.event.record(self.stream)

    def kept(self):
        """Prepared from points (``UNetSCN.prepare_points``): ``None`` when every point lies inside the receptive
        field, else the bool ``[N]`` mask of the points the reference's loader keeps -- features and labels passed
        to the forward must be filtered with it.  Waits for the build's row counts (the build's one host sync)."""
        if self.points is None:
            return None
        meta, keep = self.points.resolve()
        if meta is not self.meta:  # rebuilt from the kept points, on the stream current at the time of this call
            self.meta = meta
            self.stream = torch.cuda.current_stream(meta.device)
            self.event = torch.cuda.Event()
            self.event.record(self.stream)
        return keep

    @property
    def min_value(self):
        return self.points.min_value

    @property
    def offset(self):
        return self.points.offset

    @property
    def n_points(self):
        return self.meta.n_points


def prepare(net, coords, wait=False):
    """Enqueue the structure build for ``coords`` ([N, 4] int64: x, y, z, batch) on the CURRENT stream and return a
    :class:`PreparedScans`.  Nothing blocks here unless ``wait``: the one host synchronisation of a build (reading
    the row counts) is deferred to the forward that uses the structure, by which time a build enqueued one step
    ahead has long finished."""
    if not isinstance(coords, torch.Tensor) or not coords.is_cuda:
        raise RuntimeError("UNetSCN.prepare: coordinates must be a CUDA tensor -- mm2d3d_b200 has no CPU path")
    spatial0 = int(net.layer1.spatial_size[0])
    L = net._num_planes
    with torch.cuda.device(coords.device):
        meta = Metadata(coords, spatial0, L, plans=F.DEFAULT_MODE != "fp32", defer_sync=not wait)
        return PreparedScans(meta, F.DEFAULT_MODE, spatial0, L)


def prepare_points(net, points, sample_offsets, rot, transl_u, scale, wait=False):
    """:func:`prepare` from raw float points: the reference's ``augment_and_scale_3d`` + cast + filter fused into the
    voxel-hash insert (``augment.voxelize_points``).  ``rot`` / ``transl_u`` are the host-side random draws
    (``augment.draw_augmentation``).  The returned handle's :meth:`PreparedScans.kept` tells which points survive
    (``None`` = all), ``.min_value`` / ``.offset`` are what ``augment_and_scale_3d`` returns beside the coordinates."""
    from .augment import voxelize_points
    spatial0 = int(net.layer1.spatial_size[0])
    L = net._num_planes
    ps = voxelize_points(points, sample_offsets, rot, transl_u, scale, spatial0, L, plans=F.DEFAULT_MODE != "fp32",
                         defer_sync=not wait)
    with torch.cuda.device(ps.meta.device):
        prep = PreparedScans(ps.meta, F.DEFAULT_MODE, spatial0, L, points=ps)
    if wait:
        prep.kept()
    return prep


def run(net, coords, feats, rgb_mask=None):
    if not feats.is_cuda:
        raise RuntimeError("UNetSCN: features must be a CUDA tensor -- mm2d3d_b200 has no CPU path")
    spatial0 = int(net.layer1.spatial_size[0])
    L = net._num_planes
    if isinstance(coords, PreparedScans):
        prep = coords
        if (prep.mode, prep.spatial0, prep.levels) != (F.DEFAULT_MODE, spatial0, L) or prep.meta.device != feats.device:
            raise ValueError("UNetSCN: the prepared structure was built for another network, device or convolution mode")
        prep.kept()  # (prepared from points: settles which structure is used)
        meta = prep.meta.finish()
        cur = torch.cuda.current_stream(feats.device)
        if cur != prep.stream:
            cur.wait_event(prep.event)
            meta.record_stream(cur)
    else:
        if coords.device != feats.device:
            coords = coords.to(feats.device, non_blocking=True)
        meta = Metadata(coords, spatial0, L, plans=F.DEFAULT_MODE != "fp32")
    if feats.shape[0] < meta.n_points:
        raise ValueError(f"UNetSCN: {feats.shape[0]} feature rows for {meta.n_points} points")
    bn0 = net.layer4
    cfg = (net.in_channels, net.out_channels, L, _lib.MODES[F.DEFAULT_MODE], bool(net.training), float(bn0.eps),
           float(bn0.momentum), spatial0)
    mask_w, mask_b = rgb_mask if rgb_mask is not None else (None, None)
    with torch.autocast("cuda", enabled=False):
        return UNetSCNFn.apply(feats, meta, cfg, mask_w, mask_b, *collect_slots(net))
