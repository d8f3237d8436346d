#!/usr/bin/env python
"""Raw float points -> level-0 structure: the fused path (mm3d_voxelize_points) against the two steps it
replaces (mm3d_scale_points writing int64 [N, 4], then mm3d_voxelize reading it back).  Batch of 8 nuScenes-shaped scans,
CUDA events around 50 calls of each, L2 left warm for both (the same small working set).  A development measurement."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from mm2d3d_b200 import synth  # noqa: E402
from mm2d3d_b200.augment import scale_points, voxelize_points  # noqa: E402
from mm2d3d_b200.metadata import Metadata  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    scans = [synth.raycast_points("nuscenes", s) for s in range(8)]
    offs = np.concatenate([[0], np.cumsum([len(p) for p in scans])]).astype(np.int64)
    pts = torch.from_numpy(np.concatenate(scans, 0)).to(dev)
    rot = np.stack([np.eye(3, dtype=np.float32)] * 8)
    n = pts.shape[0]

    def two_step(levels):
        coords, keep, _, _ = scale_points(pts, offs, rot, None, 20.0, 4096)
        return Metadata(coords, 4096, levels, defer_sync=True)

    def fused(levels):
        return voxelize_points(pts, offs, rot, None, 20.0, 4096, prebuild_levels=levels, defer_sync=True).meta

    for levels in (1, 7):
        for name, fn in (("scale_points + voxelize", two_step), ("voxelize_points (fused)", fused)) * 2:
            for _ in range(5):
                fn(levels)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(50):
                fn(levels)
            e1.record()
            torch.cuda.synchronize()
            print(f"{levels} level(s), {n} points: {name:26s} {e0.elapsed_time(e1) / 50 * 1e3:8.1f} us per build")


if __name__ == "__main__":
    main()
