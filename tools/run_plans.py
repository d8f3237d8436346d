#!/usr/bin/env python
"""Build the structure and the 19 row plans of one batch-8 nuScenes-shaped input (for ncu captures)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mm2d3d_b200 import synth  # noqa: E402
from mm2d3d_b200.metadata import Metadata  # noqa: E402

locs, _ = synth.make_batch(sys.argv[1] if len(sys.argv) > 1 else "nuscenes", batch=8)
locs = torch.from_numpy(locs).cuda()
for _ in range(2):
    m = Metadata(locs, 4096, 7)
    s = 4096
    for l in range(7):
        m.plan("smc", s)
        if l < 6:
            m.plan("down", s)
            m.plan("up", s)
        s //= 2
    torch.cuda.synchronize()
print("ok")
