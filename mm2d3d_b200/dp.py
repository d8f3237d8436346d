"""Scan-sharded data parallelism: one process per GPU, each rank runs the whole hot path on its
own scans (samples never interact in hashing, rulebooks, convolutions or the I/O layers; BN
statistics stay per rank exactly like the reference's DDP without SyncBN, ``run.py:262-268``).
The only exchange step is the gradient mean.

Instead of DDP's 25 MB buckets and per-forward buffer broadcasts, all parameter gradients live in
ONE flat float32 buffer (10.76 MB for UNetSCN) and a single ``all_reduce`` over NCCL (NVLink 5 /
NVSwitch) averages everything.  At this size the collective is latency-bound, so it is not split.

Two ways to get there.  The whole-network executor (``executor.py``) already writes every parameter
gradient of a backward into one flat tensor and hands autograd views of it: with ``p.grad = None``
before backward (``zero_grad(set_to_none=True)``, PyTorch's default) those views BECOME ``p.grad``
without any copy or add kernel, and :meth:`FlatGradAllReduce.all_reduce_mean` reduces their common
base tensor in place.  Otherwise (module-by-module path, other networks) ``p.grad`` is pointed at
views of a buffer owned here and backward accumulates into it.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class FlatGradAllReduce:
    def __init__(self, module: torch.nn.Module, process_group=None):
        self.params = [p for p in module.parameters() if p.requires_grad]
        self.group = process_group
        if not self.params:
            raise ValueError("module has no trainable parameters")
        dev, dtype = self.params[0].device, self.params[0].dtype
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=dtype, device=dev)
        off = 0
        for p in self.params:
            if p.device != dev or p.dtype != dtype:
                raise ValueError("all parameters must share device and dtype")
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * self.flat.element_size()

    def zero_(self):
        """Start a step: drop the gradients (no kernel) so that backward's flat gradient tensor is adopted
        as ``p.grad`` as is; parameters that never receive a gradient keep ``None``."""
        for p in self.params:
            p.grad = None

    def _common_base(self):
        """One flat tensor over the storage all ``p.grad`` tile back to back, if they do (executor path:
        autograd adopts the views of the executor's flat gradient tensor as ``p.grad``)."""
        grads = [p.grad for p in self.params]
        if any(g is None or not g.is_contiguous() for g in grads):
            return None
        st = grads[0].untyped_storage()
        pos = sorted(grads, key=lambda g: g.storage_offset())
        start = expect = pos[0].storage_offset()
        for g in pos:
            if g.untyped_storage().data_ptr() != st.data_ptr() or g.storage_offset() != expect or g.dtype != grads[0].dtype:
                return None
            expect += g.numel()
        return torch.empty(0, dtype=grads[0].dtype, device=grads[0].device).set_(st, start, (expect - start,))

    def broadcast_parameters(self, src: int = 0):
        """One-time parameter sync at start-up (what DDP does when it wraps a module)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            for p in self.params:
                dist.broadcast(p.data, src, group=self.group)

    def all_reduce_mean(self, async_op: bool = False):
        """Average the flat gradient over the ranks (no-op for a single process)."""
        if not (dist.is_available() and dist.is_initialized()):
            return None
        world = dist.get_world_size(self.group)
        if world == 1:
            return None
        flat = self._common_base()
        if flat is None:  # gather whatever gradients exist into the owned buffer and re-point p.grad at it
            off = 0
            for p in self.params:
                view = self.flat[off:off + p.numel()].view_as(p)
                if p.grad is None:
                    view.zero_()
                elif p.grad.data_ptr() != view.data_ptr():
                    view.copy_(p.grad)
                p.grad = view
                off += p.numel()
            flat = self.flat
        if flat.is_cuda and dist.get_backend(self.group) == "nccl":
            # NCCL averages inside the collective: no separate scaling kernel
            return dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group, async_op=async_op)
        flat.div_(world)
        return dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)


def shard_scans(n_scans: int, rank: int, world: int):
    """Indices of the scans rank ``rank`` owns when a global batch is split evenly
    (``batch_size // gpus`` per process, ``run.py:52-54``)."""
    if n_scans % world:
        raise ValueError(f"global batch {n_scans} is not divisible by world size {world}")
    per = n_scans // world
    return list(range(rank * per, (rank + 1) * per))
