"""CPU restatement of the reference's 2D->3D lift.  TEST INFRASTRUCTURE ONLY.

Follows ``2d_net/model.py:131-137`` (and the second copy at ``:163-173``): for every
sample ``i`` take the ``[C, H, W]`` map, view it channels-last and pick the C-vector at
integer pixel ``(img_indices[i][:, 0], img_indices[i][:, 1])`` = (row, col); concatenate
the samples.  Backward is a scatter-add (duplicates accumulate) -- torch autograd of
advanced indexing, as in the reference.

Pinned: ``tests/golden/lift_*.npz`` were produced by importing the reference's own
``L2G_classifier_2D`` (see ``tests/golden/make_golden.py``).
"""
import torch


def lift2d(fmap: torch.Tensor, img_indices) -> torch.Tensor:
    out = []
    for i in range(fmap.shape[0]):
        idx = torch.as_tensor(img_indices[i], dtype=torch.long)
        out.append(fmap.permute(0, 2, 3, 1)[i][idx[:, 0], idx[:, 1]])
    return torch.cat(out, 0)
