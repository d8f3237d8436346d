"""CPU restatement (test infrastructure only) of the data-dependent part of the reference's
``augment_and_scale_3d`` (``lib/utils/augmentation_3d.py:143-158``) + the loaders' integer cast and range filter
(``lib/dataset/nuscenes_dataloader.py:323-327``), given the augmentation matrix and the translation draws.
Pinned against ``tests/golden/augment_ref.npz``, which the reference function itself produced."""
import numpy as np


def scale_points(points, rot, transl_u, scale, full_scale):
    """One sample.  points float32 [N,3]; rot float32 [3,3]; transl_u float64 [3] or None.
    Returns (coords int64 [N,3], keep bool [N], min_value float32 [3], offset float64 [3])."""
    points = np.asarray(points, dtype=np.float32)
    if not np.array_equal(rot, np.eye(3, dtype=np.float32)):
        points = points.dot(np.asarray(rot, dtype=np.float32))      # augmentation_3d.py:141
    coords = points * scale                                          # :144
    min_value = coords.min(0)                                        # :146
    coords -= min_value
    offset = np.zeros(3)
    if transl_u is not None:                                         # :150-155
        offset = np.clip(full_scale - coords.max(0) - 0.001, a_min=0, a_max=None) * np.asarray(transl_u)
        coords += offset
    ci = coords.astype(np.int64)                                     # nuscenes_dataloader.py:324
    keep = (ci.min(1) >= 0) * (ci.max(1) < full_scale)               # :327
    return ci, keep, min_value, offset
