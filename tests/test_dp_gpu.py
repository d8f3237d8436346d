"""Scan-sharded data parallelism on the GPU: two processes run the CUDA executor on their shards, exchange the flat
gradient (``FlatGradAllReduce``: autograd adopts views of the executor's flat gradient tensor, ``_common_base`` rebuilds
one tensor over them, one in-place all-reduce) and take an SGD step; a single process that evaluates both shards one
after the other and averages is the reference.  NCCL (``ReduceOp.AVG``) when the box has two GPUs; on a one-GPU box
the two ranks share the device and exchange through gloo (same CUDA tensors, same host logic)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

NET = dict(in_channels=3, m=16, num_planes=4, full_scale=256)
LR = 0.05


def _shard(scans):
    """One point per voxel (no float atomics in the I/O layers): forward and d_feats are bit-reproducible."""
    from mm2d3d_b200 import synth
    locs, feats = [], []
    for i, s in enumerate(scans):
        l, f = synth.make_batch("nuscenes", batch=1, seed0=60 + s)
        l[:, :3] //= 16
        l[:, 3] = i
        l, first = np.unique(l, axis=0, return_index=True)
        locs.append(l)
        feats.append(f[first])
    return torch.from_numpy(np.concatenate(locs)), torch.from_numpy(np.concatenate(feats))


def _step(net, flat, scans, dev, mode):
    import mm2d3d_b200.scn as scn
    coords, feats = _shard(scans)
    scn.set_conv_mode(mode)
    try:
        if flat is not None:
            flat.zero_()
        else:
            for p in net.parameters():
                p.grad = None
        out = net([coords.to(dev), feats.to(dev)])
        out.square().mean().backward()
    finally:
        scn.set_conv_mode("fp32")


def _worker(rank, world, port, backend, mode, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = torch.device("cuda", rank if backend == "nccl" else 0)
    torch.cuda.set_device(dev)
    if backend == "nccl":
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    from mm2d3d_b200.dp import FlatGradAllReduce, shard_scans
    from mm2d3d_b200.unet import UNetSCN
    torch.manual_seed(100 + rank)  # different initial weights per rank: the broadcast must fix them
    net = UNetSCN(**NET).to(dev)
    flat = FlatGradAllReduce(net)
    flat.broadcast_parameters(0)
    _step(net, flat, shard_scans(4, rank, world), dev, mode)
    base = flat._common_base()
    adopted = base is not None and base.data_ptr() != flat.flat.data_ptr()  # the executor's tensor, not a copy
    flat.all_reduce_mean()
    with torch.no_grad():
        for p in net.parameters():
            p.add_(p.grad, alpha=-LR)
    torch.cuda.synchronize()
    g = torch.cat([p.grad.flatten() for p in net.parameters()]).cpu().numpy()
    w = torch.cat([p.detach().flatten() for p in net.parameters()]).cpu().numpy()
    q.put((rank, g, w, adopted))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["fp32", "tf32"])
def test_flat_gradient_allreduce_on_gpu(mode):
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, backend, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(2):
        rank, g, w, adopted = q.get(timeout=600)
        got[rank] = (g, w, adopted)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    # every rank ends with the same averaged gradient and the same parameters
    assert np.array_equal(got[0][0], got[1][0]) and np.array_equal(got[0][1], got[1][1])
    assert got[0][2] and got[1][2], "p.grad were not views of the executor's flat gradient tensor"

    from mm2d3d_b200.dp import shard_scans
    from mm2d3d_b200.unet import UNetSCN
    dev = torch.device("cuda", 0)
    torch.manual_seed(100)  # rank 0's initial weights
    net = UNetSCN(**NET).to(dev)
    w0 = torch.cat([p.detach().flatten() for p in net.parameters()]).cpu().numpy()
    want = None
    for r in range(2):
        _step(net, None, shard_scans(4, r, 2), dev, mode)
        g = torch.cat([p.grad.flatten() for p in net.parameters()]).double().cpu().numpy()
        want = g if want is None else want + g
    want = want / 2
    g_dp, w_dp = got[0][0].astype(np.float64), got[0][1].astype(np.float64)
    rel = np.linalg.norm(g_dp - want) / np.linalg.norm(want)
    print(f"\n[{backend}, {mode}] averaged gradient vs single-process evaluation of both shards: rel-L2 {rel:.2e}")
    # same kernels on the same inputs: only the order of the weight-gradient / BatchNorm atomics differs
    assert rel < 1e-4, rel
    assert np.allclose(w_dp, w0 - LR * want, rtol=0, atol=1e-6 + 1e-4 * np.abs(LR * want).max())
