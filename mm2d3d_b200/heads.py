"""Point-wise prologue and loss around the 3D network, each one kernel per direction.

* :func:`rgb_mask` -- ``Net3DSeg.forward``'s ``feats *= sigmoid(linear_rgb_mask(feats))``
  (``3d_net/model.py:46-48``); takes the ``nn.Linear(C, 1)`` parameters as they are.
* :func:`cross_modal_kl` -- one term of ``TrainModel.cross_modal_loss`` (``train.py:157-184``):
  ``F.kl_div(F.log_softmax(pred, 1), F.softmax(target.detach(), 1), reduction="none").sum(1).mean()``.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, lib, ptr


def _f32c(t):
    return t.detach().to(torch.float32).contiguous()


class _RgbMaskFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, weight, bias):
        if not feats.is_cuda:
            raise RuntimeError("rgb_mask: CUDA tensors expected (there is no CPU path)")
        x, w, b = _f32c(feats), _f32c(weight).reshape(-1), _f32c(bias).reshape(-1)
        n, c = x.shape
        if w.numel() != c or b.numel() != 1:
            raise ValueError("rgb_mask: weight must be [1, C] and bias [1] (nn.Linear(C, 1))")
        y = torch.empty_like(x)
        s = torch.empty(n, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            check(lib.mm3d_rgb_mask_fwd(ptr(x), n, c, ptr(w), ptr(b), ptr(y), ptr(s), _lib.stream_ptr()), "mm3d_rgb_mask_fwd")
        ctx.save_for_backward(x, s, w)
        ctx.shapes = (weight.shape, bias.shape)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, s, w = ctx.saved_tensors
        dy = _f32c(dy)
        n, c = x.shape
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dwb = torch.empty(c + 1, dtype=torch.float32, device=x.device)
        ws = torch.empty(c + 1, dtype=torch.float64, device=x.device)
        with torch.cuda.device(x.device):
            check(lib.mm3d_rgb_mask_bwd(ptr(x), ptr(s), ptr(dy), n, c, ptr(w), ptr(dx), ptr(dwb), dwb.data_ptr() + 4 * c,
                                        ptr(ws), ws.numel() * 8, _lib.stream_ptr()), "mm3d_rgb_mask_bwd")
        return dx, dwb[:c].reshape(ctx.shapes[0]), dwb[c:].reshape(ctx.shapes[1])


def rgb_mask(feats: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """``feats * sigmoid(feats @ weight.T + bias)``; feats ``[N, C]`` (C <= 8), ``weight [1, C]``, ``bias [1]``."""
    return _RgbMaskFn.apply(feats, weight, bias)


class _KlFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target):
        if not pred.is_cuda:
            raise RuntimeError("cross_modal_kl: CUDA tensors expected (there is no CPU path)")
        p, t = _f32c(pred), _f32c(target)
        if p.shape != t.shape or p.dim() != 2 or p.shape[0] == 0:
            raise ValueError("cross_modal_kl: two non-empty [N, C] logit tensors expected")
        n, C = p.shape
        loss = torch.empty((), dtype=torch.float32, device=p.device)
        ws = torch.empty(1, dtype=torch.float64, device=p.device)
        with torch.cuda.device(p.device):
            check(lib.mm3d_kl_logits_fwd(ptr(p), ptr(t), n, C, ptr(loss), ptr(ws), 8, _lib.stream_ptr()), "mm3d_kl_logits_fwd")
        ctx.save_for_backward(p, t)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        p, t = ctx.saved_tensors
        n, C = p.shape
        d = _f32c(dloss).reshape(1)
        dpred = torch.empty_like(p)
        with torch.cuda.device(p.device):
            check(lib.mm3d_kl_logits_bwd(ptr(p), ptr(t), n, C, ptr(d), ptr(dpred), _lib.stream_ptr()), "mm3d_kl_logits_bwd")
        return dpred, None


def cross_modal_kl(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """Mean over rows of ``KL(softmax(target) || softmax(pred))``; ``target`` is treated as detached."""
    return _KlFn.apply(pred, target)
