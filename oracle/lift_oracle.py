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


def lift2d_bilinear(fmap: torch.Tensor, pixel_coords) -> torch.Tensor:
    """Bilinear variant -- NOT in the reference (which only gathers at integer pixels); the definition is torch's
    ``F.grid_sample(mode="bilinear", padding_mode="zeros", align_corners=True)`` and ``tests/test_oracle.py`` pins this
    explicit four-tap restatement to it.  ``pixel_coords[i]``: float ``[N_i, 2]`` (row, col), pixel centres at integers."""
    B, C, H, W = fmap.shape
    out = []
    for i in range(B):
        rc = torch.as_tensor(pixel_coords[i], dtype=torch.float32).reshape(-1, 2)
        r0, c0 = torch.floor(rc[:, 0]), torch.floor(rc[:, 1])
        fr, fc = rc[:, 0] - r0, rc[:, 1] - c0
        acc = torch.zeros(rc.shape[0], C, dtype=torch.float32)
        for dr in (0, 1):
            for dc in (0, 1):
                rr, cc = (r0 + dr).long(), (c0 + dc).long()
                inside = (rr >= 0) & (rr < H) & (cc >= 0) & (cc < W)
                w = (fr if dr else 1 - fr) * (fc if dc else 1 - fc) * inside
                v = fmap[i].float()[:, rr.clamp(0, H - 1), cc.clamp(0, W - 1)].t()  # [N, C]
                acc = acc + w[:, None] * v
        out.append(acc.to(fmap.dtype))
    return torch.cat(out, 0)
