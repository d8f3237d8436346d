"""CPU restatement of the SparseConvNet semantics the reference's 3D branch relies on.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): the product never imports this.
PARITY UNPINNED by reference-owned tests -- SparseConvNet (pinned at dcf6a7ff in
``/root/reference/environment.yml:37``) is not vendored; what is restated here is its
published behaviour as used by ``3d_net/scn_unet.py``, and it is
pinned by the dense-conv3d equivalence / known-answer / gradcheck tests in ``tests/``.

Everything is numpy (integer structure) + torch CPU ops (floating point, any dtype, with
autograd).  The floating-point ops follow SparseConvNet's CPU algorithm (Appendix A.8):
per kernel offset ``index_select -> matmul -> index_add_`` over that offset's rule list.

Conventions
-----------
* ``coords``: int64 ``[N, 4]`` = (x, y, z, batch), batch LAST
  (``lib/dataset/__init__.py:62-67``, ``3d_net/scn_unet.py:131-133``).
* voxel rows are numbered by FIRST OCCURRENCE in the concatenated point list (A.2);
  coarse rows by first occurrence when scanning the fine rows in row order (A.4,
  canonical choice -- SparseConvNet's own order there is hash-iteration order).
* 3^3 offset index ``k = ((dx+1)*3 + (dy+1))*3 + (dz+1)`` (A.3); 2^3 offset index
  ``k = ((x&1)*2 + (y&1))*2 + (z&1)`` (A.4).
"""
from __future__ import annotations

import numpy as np
import torch

# --------------------------------------------------------------------------------------
# integer structure: voxel ids, level pyramid, rule tables
# --------------------------------------------------------------------------------------


def pack_keys(coords: np.ndarray) -> np.ndarray:
    """(x,y,z,b) int64 rows -> one int64 key ``b<<48 | x<<32 | y<<16 | z``."""
    c = np.asarray(coords, dtype=np.int64)
    return (c[:, 3] << 48) | (c[:, 0] << 32) | (c[:, 1] << 16) | c[:, 2]


def unpack_keys(keys: np.ndarray) -> np.ndarray:
    k = np.asarray(keys, dtype=np.int64)
    out = np.empty((k.shape[0], 4), dtype=np.int64)
    out[:, 3] = k >> 48
    out[:, 0] = (k >> 32) & 0xFFFF
    out[:, 1] = (k >> 16) & 0xFFFF
    out[:, 2] = k & 0xFFFF
    return out


def first_occurrence_ids_loop(keys):
    """Literal restatement of SparseConvNet's InputLayer / SparseGrid numbering (A.2):
    walk the items in order, the first time a key is seen it gets ``nActive++``.
    Pure Python -- for small cases and for validating the vectorised version."""
    table = {}
    ids = np.empty(len(keys), dtype=np.int64)
    uniq = []
    for i, k in enumerate(keys.tolist()):
        j = table.get(k)
        if j is None:
            j = len(uniq)
            table[k] = j
            uniq.append(k)
        ids[i] = j
    return ids, np.asarray(uniq, dtype=np.int64)


def first_occurrence_ids(keys: np.ndarray):
    """Vectorised equivalent of :func:`first_occurrence_ids_loop`."""
    keys = np.asarray(keys, dtype=np.int64)
    if keys.shape[0] == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    uniq, first, inv = np.unique(keys, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(order.shape[0])
    return rank[inv.reshape(-1)], uniq[order]


def voxelize(coords: np.ndarray):
    """InputLayer structure (A.2): returns (p2v [N], voxel coords [N0,4], points-per-voxel [N0])."""
    ids, uniq = first_occurrence_ids(pack_keys(coords))
    npts = np.bincount(ids, minlength=uniq.shape[0]).astype(np.int64)
    return ids, unpack_keys(uniq), npts


def coarsen(vcoords: np.ndarray):
    """Stride-2/size-2 structure (A.4): parent row, 2^3 offset index and coarse coords."""
    c = np.asarray(vcoords, dtype=np.int64)
    par = c.copy()
    par[:, :3] >>= 1
    off = ((c[:, 0] & 1) * 2 + (c[:, 1] & 1)) * 2 + (c[:, 2] & 1)
    parent, uniq = first_occurrence_ids(pack_keys(par))
    return parent, off, unpack_keys(uniq)


OFFSETS27 = np.array(
    [(dx, dy, dz) for dx in (-1, 0, 1) for dy in (-1, 0, 1) for dz in (-1, 0, 1)], dtype=np.int64
)


def nbr_table(vcoords: np.ndarray, spatial_size: int) -> np.ndarray:
    """3^3 submanifold rule table (A.3): ``tbl[j, k]`` = row of the active voxel at
    ``pos(j) + d_k`` in the same sample, or -1.  Equivalent to SparseConvNet's rulebook:
    ``rules[k] = {(tbl[j,k], j) : tbl[j,k] >= 0}``."""
    c = np.asarray(vcoords, dtype=np.int64)
    n = c.shape[0]
    tbl = np.full((n, 27), -1, dtype=np.int32)
    if n == 0:
        return tbl
    keys = pack_keys(c)
    order = np.argsort(keys, kind="stable")
    skeys = keys[order]
    for k, d in enumerate(OFFSETS27):
        q = c.copy()
        q[:, :3] += d
        ok = np.all((q[:, :3] >= 0) & (q[:, :3] < spatial_size), axis=1)
        qk = pack_keys(q)
        pos = np.searchsorted(skeys, qk)
        pos[pos >= n] = n - 1
        hit = ok & (skeys[pos] == qk)
        tbl[hit, k] = order[pos[hit]].astype(np.int32)
    return tbl


def nbr_table_loop(vcoords, spatial_size):
    """Literal dict-probe version of :func:`nbr_table` (27 probes per active voxel)."""
    grid = {tuple(int(v) for v in row): i for i, row in enumerate(vcoords)}
    tbl = np.full((len(vcoords), 27), -1, dtype=np.int32)
    for j, (x, y, z, b) in enumerate(np.asarray(vcoords).tolist()):
        for k, (dx, dy, dz) in enumerate(OFFSETS27.tolist()):
            q = (x + dx, y + dy, z + dz, b)
            if min(q[:3]) < 0 or max(q[:3]) >= spatial_size:
                continue
            i = grid.get(q)
            if i is not None:
                tbl[j, k] = i
    return tbl


def child_table(parent: np.ndarray, off: np.ndarray, n_coarse: int) -> np.ndarray:
    """``tbl[q, k]`` = fine row whose parent is q and whose 2^3 offset is k, or -1."""
    tbl = np.full((n_coarse, 8), -1, dtype=np.int32)
    tbl[parent, off] = np.arange(parent.shape[0], dtype=np.int32)
    return tbl


def up_table(parent: np.ndarray, off: np.ndarray) -> np.ndarray:
    """One-hot table of the deconvolution: ``tbl[i, k] = parent[i] if k == off[i] else -1``."""
    tbl = np.full((parent.shape[0], 8), -1, dtype=np.int32)
    tbl[np.arange(parent.shape[0]), off] = parent.astype(np.int32)
    return tbl


def rules_from_table(tbl: np.ndarray):
    """SparseConvNet rulebook form: per offset k the (in rows, out rows) of its pairs,
    canonically sorted by out row."""
    rules = []
    for k in range(tbl.shape[1]):
        out_rows = np.nonzero(tbl[:, k] >= 0)[0]
        rules.append((tbl[out_rows, k].astype(np.int64), out_rows.astype(np.int64)))
    return rules


class Metadata:
    """Per-forward structure shared by every layer (A.1, a12): level-0 voxelisation and,
    lazily, per spatial size the active set, its 3^3 rules and its stride-2 rules."""

    def __init__(self, coords, spatial_size: int):
        coords = np.asarray(coords, dtype=np.int64)
        assert coords.ndim == 2 and coords.shape[1] == 4
        self.spatial_size0 = int(spatial_size)
        self.n_points = coords.shape[0]
        self.p2v, v0, self.npts = voxelize(coords)
        self.batch_size = int(coords[:, 3].max()) + 1 if coords.shape[0] else 0
        self.level_coords = {self.spatial_size0: v0}
        self._nbr = {}
        self._down = {}
        self._rules_t = {}

    def coords_at(self, spatial_size):
        return self.level_coords[int(spatial_size)]

    def nbr(self, spatial_size):
        s = int(spatial_size)
        if s not in self._nbr:
            self._nbr[s] = nbr_table(self.level_coords[s], s)
        return self._nbr[s]

    def down(self, spatial_size):
        """(parent, off, n_coarse) for the 2/2 convolution from ``spatial_size``."""
        s = int(spatial_size)
        if s not in self._down:
            parent, off, cc = coarsen(self.level_coords[s])
            self.level_coords.setdefault(s // 2, cc)
            self._down[s] = (parent, off, cc.shape[0])
        return self._down[s]

    # torch index tensors of the rule lists, cached like SparseConvNet's rulebooks
    def rules_t(self, kind, spatial_size):
        key = (kind, int(spatial_size))
        if key not in self._rules_t:
            if kind == "smc":
                rules = rules_from_table(self.nbr(spatial_size))
            else:  # "down": in = fine row, out = coarse row, per 2^3 offset
                parent, off, _ = self.down(spatial_size)
                rules = []
                for k in range(8):
                    fine = np.nonzero(off == k)[0].astype(np.int64)
                    rules.append((fine, parent[fine].astype(np.int64)))
            self._rules_t[key] = [
                (torch.from_numpy(a), torch.from_numpy(b)) for a, b in rules
            ]
        return self._rules_t[key]


# --------------------------------------------------------------------------------------
# floating-point ops (torch CPU, differentiable, dtype follows the inputs)
# --------------------------------------------------------------------------------------


def input_layer(meta: Metadata, feats: torch.Tensor, mode: int = 4) -> torch.Tensor:
    """A.2: mode 4 = mean of the points of a voxel, mode 3 = sum."""
    n0 = meta.npts.shape[0]
    p2v = torch.from_numpy(meta.p2v)
    feats = feats[: meta.n_points]
    if mode == 4:
        w = (1.0 / torch.from_numpy(meta.npts).to(feats.dtype))[p2v]
        feats = feats * w[:, None]
    elif mode != 3:
        raise NotImplementedError("oracle restates InputLayer modes 3 and 4 only")
    out = torch.zeros(n0, feats.shape[1], dtype=feats.dtype)
    return out.index_add(0, p2v, feats)


def output_layer(meta: Metadata, vfeats: torch.Tensor) -> torch.Tensor:
    """A.7: every input point gets a copy of its voxel's row."""
    return vfeats.index_select(0, torch.from_numpy(meta.p2v))


def _rule_conv(x, rules, weight, n_out, reverse=False):
    """A.8: per offset, gather rows, dense matmul, scatter-add."""
    out = torch.zeros(n_out, weight.shape[-1], dtype=x.dtype)
    for k, (r_in, r_out) in enumerate(rules):
        if reverse:
            r_in, r_out = r_out, r_in
        if r_in.numel() == 0:
            continue
        out = out.index_add(0, r_out, x.index_select(0, r_in) @ weight[k])
    return out


def submanifold_conv(meta, spatial_size, x, weight):
    """A.3.  ``weight``: [27, 1, C_in, C_out] (SparseConvNet parameter shape) or [27, C_in, C_out]."""
    w = weight.reshape(weight.shape[0], weight.shape[-2], weight.shape[-1])
    return _rule_conv(x, meta.rules_t("smc", spatial_size), w, x.shape[0])


def conv_down(meta, spatial_size, x, weight):
    """A.4 Convolution(size 2, stride 2) from ``spatial_size`` to ``spatial_size // 2``."""
    w = weight.reshape(weight.shape[0], weight.shape[-2], weight.shape[-1])
    _, _, n_coarse = meta.down(spatial_size)
    return _rule_conv(x, meta.rules_t("down", spatial_size), w, n_coarse)


def deconv_up(meta, spatial_size_out, x, weight):
    """A.4 Deconvolution(size 2, stride 2): same rulebook as the matching Convolution,
    run backwards; the output lives on the existing grid of ``spatial_size_out``."""
    w = weight.reshape(weight.shape[0], weight.shape[-2], weight.shape[-1])
    parent, _, _ = meta.down(spatial_size_out)
    return _rule_conv(x, meta.rules_t("down", spatial_size_out), w, parent.shape[0], reverse=True)


def batchnorm_relu(x, gamma, beta, running_mean, running_var, eps=1e-4, momentum=0.9,
                   training=True, leakiness=0.0):
    """A.5.  ``momentum`` is SparseConvNet's (weight of the OLD running value)."""
    if training:
        n = x.shape[0]
        mean = x.mean(0)
        var = ((x - mean) ** 2).mean(0)
        if running_mean is not None:
            with torch.no_grad():
                running_mean.mul_(momentum).add_((1 - momentum) * mean.detach().to(running_mean.dtype))
                unbiased = var.detach() * (n / max(n - 1, 1))
                running_var.mul_(momentum).add_((1 - momentum) * unbiased.to(running_var.dtype))
    else:
        mean, var = running_mean.to(x.dtype), running_var.to(x.dtype)
    y = (x - mean) * torch.rsqrt(var + eps) * gamma + beta
    return torch.where(y > 0, y, y * leakiness)
