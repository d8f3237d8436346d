"""Synthetic LiDAR scans with the shape statistics of the reference's datasets.

There are no datasets on the build or GPU boxes; the benchmark and the parity tests use
ray-cast scans (ground plane, two street walls, 12 car-sized boxes) pushed through exactly
the reference's coordinate transform: ``coords = floor((p - min p) * scale)`` as int64, rows
outside ``[0, full_scale)`` dropped (``lib/utils/augmentation_3d.py:144-148``,
``lib/dataset/nuscenes_dataloader.py:324-327``), batch index appended as the LAST column
(``lib/dataset/__init__.py:62-67``).  Generator spec: SURVEY.md Appendix C.
"""
from __future__ import annotations

import numpy as np

SHAPES = {
    # name: (beams, azimuth steps, elevation lo/hi [deg], sensor height [m])
    "nuscenes": (32, 1085, -30.67, 10.67, 1.84),
    "semantickitti": (64, 1900, -24.8, 2.0, 1.73),
}


def raycast_points(shape: str = "nuscenes", seed: int = 0) -> np.ndarray:
    """float32 ``[N, 3]`` points of one synthetic sweep (sensor frame, metres)."""
    nbeams, naz, elo, ehi, h = SHAPES[shape]
    rng = np.random.default_rng(seed)
    el = np.deg2rad(np.linspace(elo, ehi, nbeams))[:, None]
    az = np.linspace(-np.pi, np.pi, naz, endpoint=False)[None, :]
    d = np.stack(
        [np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el) * np.ones_like(az)], -1
    ).reshape(-1, 3)
    r = np.full(d.shape[0], np.inf)
    with np.errstate(divide="ignore", invalid="ignore"):
        rg = np.where(d[:, 2] < 0, -h / d[:, 2], np.inf)
        r = np.minimum(r, rg)
        for w in (9.0, -12.0):
            rw = np.where(d[:, 1] * w > 0, w / d[:, 1], np.inf)
            r = np.minimum(r, rw)
        for _ in range(12):
            xc = rng.uniform(-40, 40)
            yc = rng.uniform(-7, 7)
            rc = np.where(d[:, 0] * xc > 0, xc / d[:, 0], np.inf)
            z = h + rc * d[:, 2]
            ok = (np.abs(rc * d[:, 1] - yc) < 1.0) & (z > 0) & (z < 1.6)
            r = np.minimum(r, np.where(ok, rc, np.inf))
    r = r + rng.normal(0.0, 0.02, r.shape[0])
    keep = (r > 1.0) & (r < 80.0)
    p = r[keep, None] * d[keep] + np.array([0.0, 0.0, h])
    return p.astype(np.float32)


def scan_coords(shape="nuscenes", seed=0, scale=20, full_scale=4096) -> np.ndarray:
    """int64 ``[N, 3]`` voxel coordinates of one scan (reference transform)."""
    p = raycast_points(shape, seed).astype(np.float64)
    c = np.floor((p - p.min(0)) * scale).astype(np.int64)
    ok = np.all((c >= 0) & (c < full_scale), axis=1)
    return c[ok]


def make_batch(shape="nuscenes", batch=8, seed0=0, in_channels=3, scale=20, full_scale=4096):
    """Collated batch in the reference's input contract:
    ``locs`` int64 ``[N, 4]`` (x, y, z, batch) and ``feats`` float32 ``[N, C]`` ~ U[0,1)."""
    locs, feats = [], []
    for b in range(batch):
        c = scan_coords(shape, seed0 + b, scale, full_scale)
        locs.append(np.concatenate([c, np.full((c.shape[0], 1), b, np.int64)], 1))
        rng = np.random.default_rng(10_000 + seed0 + b)
        feats.append(rng.random((c.shape[0], in_channels), dtype=np.float32))
    return np.concatenate(locs, 0), np.concatenate(feats, 0)


def make_img_indices(n_points_per_sample, height=225, width=400, seed=0, window=None):
    """Per-sample int64 ``[N_i, 2]`` (row, col) pixel indices like the reference's
    ``img_indices`` (``lib/dataset/nuscenes_dataloader.py:274``); ``window`` = side of a
    duplicate-heavy square all points fall into."""
    rng = np.random.default_rng(seed)
    out = []
    for n in n_points_per_sample:
        if window:
            r = rng.integers(0, min(window, height), n)
            c = rng.integers(0, min(window, width), n)
        else:
            r = rng.integers(0, height, n)
            c = rng.integers(0, width, n)
        out.append(np.stack([r, c], 1).astype(np.int64))
    return out
