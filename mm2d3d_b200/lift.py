"""2D->3D feature lift: sample the 2D network's ``[B, C, H, W]`` map at the pixel every LiDAR
point projects to.

Replaces the per-sample advanced-indexing loop of the reference
(``2d_net/model.py:131-137`` and ``:163-173``)::

    for i in range(B): out.append(fmap.permute(0, 2, 3, 1)[i][img_indices[i][:, 0], img_indices[i][:, 1]])
    out = torch.cat(out, 0)

with one gather kernel over all samples (forward) and one scatter-add kernel (backward,
duplicates accumulate like ``index_put_(accumulate=True)``).  Indices are integer pixels
(row, col) -- the reference never interpolates.  ``LiftIndices`` uploads the per-sample numpy
index arrays ONCE per batch; the reference converts and uploads them at every use.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib  # noqa: F401  (raises when the library is missing)
from .functional import Lift2DFn


class LiftIndices:
    """Device-resident concatenation of the reference's ``img_indices`` list."""

    def __init__(self, img_indices, device):
        counts = [int(len(ix)) for ix in img_indices]
        offs = np.zeros(len(counts) + 1, dtype=np.int64)
        np.cumsum(counts, out=offs[1:])
        if counts and offs[-1] > 0:
            cat = np.concatenate([np.asarray(ix, dtype=np.int64).reshape(-1, 2) for ix in img_indices], 0)
        else:
            cat = np.zeros((0, 2), dtype=np.int64)
        self.counts = counts
        self.idx = torch.from_numpy(np.ascontiguousarray(cat)).to(device, non_blocking=True)
        self.offsets = torch.from_numpy(offs).to(device, non_blocking=True)

    @property
    def n(self):
        return int(self.idx.shape[0])


def lift2d(fmap: torch.Tensor, img_indices) -> torch.Tensor:
    """``[B, C, H, W]`` map + per-sample (row, col) indices -> ``[sum N_i, C]`` (same dtype)."""
    if not isinstance(img_indices, LiftIndices):
        img_indices = LiftIndices(img_indices, fmap.device)
    if len(img_indices.counts) != fmap.shape[0]:
        raise ValueError("lift2d: one index array per sample expected")
    return Lift2DFn.apply(fmap, img_indices.idx, img_indices.offsets)
