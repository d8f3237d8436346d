"""Points (metres) -> SparseConvNet voxel coordinates on the GPU: the reference's ``augment_and_scale_3d``
(``lib/utils/augmentation_3d.py:83-158``) with the integer cast and receptive-field filter that follow it in the
data loaders (``lib/dataset/nuscenes_dataloader.py:312-327``), for a whole collated batch in one call.

The random draws stay on the host and follow the reference's order on ``numpy.random`` exactly
(:func:`draw_augmentation`); the data-dependent work -- rotation, scaling, per-sample min / max, translation,
``astype(int64)``, the ``[0, full_scale)`` test -- is ``mm3d_scale_points`` (``csrc/augment.cu``), or, fused with the
voxel-hash insert of ``scn.InputLayer`` so that the int64 ``[N, 4]`` tensor never exists, ``mm3d_voxelize_points``
(:func:`voxelize_points`, ``UNetSCN.prepare_points``).  SURVEY.md section 8(f), row 1.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import check, lib


def draw_augmentation(noisy_rot=0.0, flip_x=0.0, flip_y=0.0, rot_z=0.0, rot_y=0.0, transl=False, rng=np.random):
    """The random part of ``augment_and_scale_3d``: returns ``(rot_matrix float32 [3,3], u float64 [3] or None)``.
    Same calls, in the same order, on ``rng`` (default: the global ``numpy.random`` the reference uses), so a seeded
    stream gives the reference's matrix and translation draws."""
    rot = np.eye(3, dtype=np.float32)
    if noisy_rot > 0 or flip_x > 0 or flip_y > 0 or rot_z > 0 or rot_y > 0:
        if noisy_rot > 0:
            rot += rng.randn(3, 3) * noisy_rot
        if flip_x > 0:
            rot[0][0] *= rng.randint(0, 2) * 2 - 1
        if flip_y > 0:
            rot[1][1] *= rng.randint(0, 2) * 2 - 1
        if rot_z > 0:
            theta = rng.rand() * rot_z
            rot = rot.dot(np.array([[np.cos(theta), -np.sin(theta), 0], [np.sin(theta), np.cos(theta), 0], [0, 0, 1]],
                                   dtype=np.float32))
        if rot_y > 0:
            theta = rng.rand() * rot_y
            rot = rot.dot(np.array([[np.cos(theta), 0, np.sin(theta)], [0, 1, 0], [-np.sin(theta), 0, np.cos(theta)]],
                                   dtype=np.float32))
    u = rng.rand(3) if transl else None
    return rot, u


def _device_args(points: torch.Tensor, sample_offsets, rot, transl_u, what):
    if not isinstance(points, torch.Tensor) or not points.is_cuda:
        raise RuntimeError(f"{what}: expected a CUDA tensor -- mm2d3d_b200 has no CPU path")
    dev = points.device
    points = points.float().contiguous()
    offs = torch.as_tensor(np.asarray(sample_offsets, dtype=np.int64)).to(dev)
    B = offs.numel() - 1
    if B < 1 or int(np.asarray(sample_offsets)[-1]) != points.shape[0]:
        raise ValueError(f"{what}: sample_offsets must be [B+1] row ranges covering all {points.shape[0]} points")
    rot_d = torch.as_tensor(np.ascontiguousarray(np.asarray(rot, dtype=np.float32).reshape(B, 9))).to(dev)
    u_d = None if transl_u is None else torch.as_tensor(np.ascontiguousarray(np.asarray(transl_u, dtype=np.float64).reshape(B, 3))).to(dev)
    return dict(points=points, offsets=offs, rot=rot_d, u=u_d, B=B,
                keep=torch.empty(points.shape[0], dtype=torch.uint8, device=dev),
                min_value=torch.zeros(B, 3, dtype=torch.float32, device=dev),
                offset=torch.zeros(B, 3, dtype=torch.float64, device=dev))


def scale_points(points: torch.Tensor, sample_offsets, rot, transl_u, scale: float, full_scale: int):
    """``points`` float32 CUDA ``[N, 3]`` (samples concatenated), ``sample_offsets`` ``[B+1]`` row ranges, ``rot``
    ``[B, 3, 3]`` float32 (identity = no augmentation), ``transl_u`` ``[B, 3]`` float64 uniform draws or ``None``.

    Returns ``(coords int64 [N, 4] = (x, y, z, sample), keep bool [N], min_value float32 [B, 3], offset float64
    [B, 3])``; ``coords[keep]`` is what the reference's collate hands to ``scn.InputLayer``."""
    q = _device_args(points, sample_offsets, rot, transl_u, "scale_points")
    dev, n, B = q["points"].device, q["points"].shape[0], q["B"]
    coords = torch.empty(n, 4, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        ws = torch.empty(lib.mm3d_scale_points_workspace_bytes(B), dtype=torch.uint8, device=dev)
        check(lib.mm3d_scale_points(q["points"].data_ptr(), q["offsets"].data_ptr(), B, n, q["rot"].data_ptr(), float(scale),
                                    int(full_scale), None if q["u"] is None else q["u"].data_ptr(), coords.data_ptr(),
                                    q["keep"].data_ptr(), q["min_value"].data_ptr(), q["offset"].data_ptr(), ws.data_ptr(),
                                    ws.numel(), _lib.stream_ptr()),
              "mm3d_scale_points")
    return coords, q["keep"].bool(), q["min_value"], q["offset"]


def voxelize_points(points: torch.Tensor, sample_offsets, rot, transl_u, scale: float, full_scale: int,
                    prebuild_levels: int = 1, plans: bool = False, defer_sync: bool = False):
    """:func:`scale_points` and the structure half of ``scn.InputLayer(3, full_scale, mode)`` in one pass
    (``mm3d_voxelize_points``): the float points go straight into the voxel hash.  Arguments as
    :func:`scale_points`; returns a :class:`PointStructure`.  Row numbering, point -> voxel map and every table are
    bit-identical to ``Metadata(scale_points(...)[0][keep], full_scale, ...)``."""
    from .metadata import Metadata
    q = _device_args(points, sample_offsets, rot, transl_u, "voxelize_points")
    q["scale"] = float(scale)
    with torch.cuda.device(q["points"].device):
        meta = Metadata(None, int(full_scale), prebuild_levels, plans=plans, defer_sync=defer_sync, points=q)
    return PointStructure(meta, q, int(full_scale), prebuild_levels, plans)


class PointStructure:
    """Result of :func:`voxelize_points`.  :meth:`resolve` (one host synchronisation, the same one every structure
    build has) returns ``(metadata, keep)``: ``keep`` is ``None`` when every point lies inside the receptive field --
    nuScenes / SemanticKITTI scans at scale 20 in a 4096^3 grid always do -- and otherwise the bool ``[N]`` mask of
    the survivors, in which case the structure has been rebuilt from them (``scale_points`` + ``InputLayer`` build:
    the reference drops such points before the collate, ``nuscenes_dataloader.py:326-327``) and the caller filters
    features and labels with it, as the reference does."""

    def __init__(self, meta, args, full_scale, prebuild_levels, plans):
        self.meta, self._args, self._cfg = meta, args, (full_scale, prebuild_levels, plans)
        self.keep = None
        self.min_value, self.offset = args["min_value"], args["offset"]
        self._resolved = False

    def resolve(self):
        from .metadata import Metadata
        if not self._resolved:
            self.meta.finish()
            if self.meta.dropped:
                q, (full_scale, levels, plans) = self._args, self._cfg
                with torch.cuda.device(q["points"].device):
                    coords, keep, _, _ = scale_points(q["points"], q["offsets"].cpu().numpy(), q["rot"].cpu().numpy(),
                                                      None if q["u"] is None else q["u"].cpu().numpy(), q["scale"],
                                                      full_scale)
                    self.meta = Metadata(coords[keep], full_scale, levels, plans=plans)
                self.keep = keep
            self._resolved = True
        return self.meta, self.keep
