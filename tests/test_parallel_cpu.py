"""CPU tests (gloo, world size 2) of the data-parallel plumbing: cloud sharding, flat gradient buffer and its
all-reduce, parameter broadcast (3d_recognizer_b200/parallel.py)."""
import importlib
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_range_partitions_everything():
    par = importlib.import_module("3d_recognizer_b200.parallel")
    for n in (0, 1, 7, 32, 64):
        for world in (1, 2, 3, 8):
            spans = [par.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    par = importlib.import_module("3d_recognizer_b200.parallel")
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(rank)                       # ranks start from different weights
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.BatchNorm1d(7), torch.nn.Linear(7, 3))
    par.broadcast_parameters(net, src=0)
    flat = par.FlatGradients(net)
    x = torch.full((4, 5), float(rank + 1))
    flat.zero()
    net(x).sum().backward()
    flat.rebind()
    local = flat.flat.clone()
    flat.allreduce_mean()
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    ok_mean = torch.allclose(flat.flat, sum(gathered) / world, atol=1e-6)
    ok_views = all(p.grad.data_ptr() >= flat.flat.data_ptr() and
                   p.grad.data_ptr() < flat.flat.data_ptr() + flat.flat.numel() * 4 for p in flat.params)
    w0 = [torch.zeros_like(net[0].weight) for _ in range(world)]
    dist.all_gather(w0, net[0].weight.data)
    ok_bcast = all(torch.equal(w0[0], w) for w in w0)
    out[rank] = bool(ok_mean and ok_views and ok_bcast)
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_flat_gradient_allreduce_gloo_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}
