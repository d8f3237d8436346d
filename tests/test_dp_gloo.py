"""N>1 host logic on CPU: world_size-2 gloo run of the scan-sharded data-parallel path
(flat gradient buffer + one all-reduce), checked against a single process that evaluates the same
two shards one after the other and averages -- BN statistics are per rank in both, like the
reference's DDP without SyncBN (run.py:262-268)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import scn_cpu
from mm2d3d_b200.dp import FlatGradAllReduce, shard_scans
from mm2d3d_b200.unet import UNetSCN

NET = dict(in_channels=3, m=4, num_planes=3, full_scale=64, backend=scn_cpu)


def _scan(seed, b):
    rng = np.random.default_rng(seed)
    u = rng.integers(0, 20, (150, 2))
    z = (u[:, 0] // 3 + rng.integers(0, 2, 150)) % 20
    c = np.concatenate([np.stack([u[:, 0], u[:, 1], z], 1) + 8, np.full((150, 1), b)], 1).astype(np.int64)
    return torch.from_numpy(c), torch.from_numpy(rng.random((150, 3), dtype=np.float32))


def _shard_grads(net, scans):
    coords = torch.cat([_scan(s, i)[0] for i, s in enumerate(scans)])
    feats = torch.cat([_scan(s, i)[1] for i, s in enumerate(scans)])
    out = net([coords, feats])
    out.square().sum().backward()


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)  # deliberately different initial weights: broadcast must fix them
    net = UNetSCN(**NET)
    flat = FlatGradAllReduce(net)
    flat.broadcast_parameters(0)
    flat.zero_()
    _shard_grads(net, shard_scans(4, rank, world))
    flat.all_reduce_mean()
    if rank == 0:
        q.put((flat.flat.clone().numpy(), torch.cat([p.detach().flatten() for p in net.parameters()]).numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_gradient_allreduce_two_ranks():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    flat, params = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    torch.manual_seed(100)  # rank 0's initial weights
    net = UNetSCN(**NET)
    assert np.allclose(torch.cat([p.detach().flatten() for p in net.parameters()]).numpy(), params)
    want = None
    for r in range(2):
        for p in net.parameters():
            p.grad = None
        _shard_grads(net, shard_scans(4, r, 2))
        g = torch.cat([p.grad.flatten() for p in net.parameters()])
        want = g if want is None else want + g
    want = (want / 2).numpy()
    assert np.allclose(flat, want, rtol=1e-5, atol=1e-6)


def test_shard_scans():
    assert shard_scans(8, 1, 4) == [2, 3]
    assert sorted(sum((shard_scans(16, r, 8) for r in range(8)), [])) == list(range(16))
    try:
        shard_scans(6, 0, 4)
        raise AssertionError("uneven split must raise")
    except ValueError:
        pass
