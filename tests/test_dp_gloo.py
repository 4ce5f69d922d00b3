"""N>1 host logic on CPU: world_size-2 gloo run of the flat-bucket gradient all-reduce."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out, attached=False):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "deep-multiview-depth-estimation_b200"))
    import torch.distributed as dist
    from mvs_b200.harness import FlatGradAllReduce
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)                       # different init per rank on purpose
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 1))
    params = list(net.parameters())                     # a plain list, like MVSNet.parameters (model.py:164)
    red = FlatGradAllReduce(params)
    red.broadcast_parameters()
    torch.manual_seed(7)
    x_all = torch.randn(8, 6); y_all = torch.randn(8, 1)
    x, y = x_all[rank::world], y_all[rank::world]       # shard the batch by rank
    if attached:                                        # .grad of every parameter is a view into the flat bucket
        red.attach_grads()
        assert all(p.grad.untyped_storage().data_ptr() == red.bucket.untyped_storage().data_ptr() for p in params)
        ((net(x) - y) ** 2).mean().backward()           # a first pass that the next step's bucket.zero_() must wipe
        red.bucket.zero_()
    ((net(x) - y) ** 2).mean().backward()
    red.reduce()
    if attached:
        assert all(p.grad is v for p, v in zip(params, red.views))       # still the views: nothing was re-allocated
    out[rank] = [p.grad.clone() for p in params] + [p.detach().clone() for p in params]
    dist.destroy_process_group()


import pytest


@pytest.mark.parametrize("attached", [False, True])
def test_flat_bucket_allreduce_equals_full_batch_gradient(attached):
    world, port = 2, _free_port()
    mgr = mp.Manager(); out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out, attached), nprocs=world, join=True)
    g0, g1 = out[0], out[1]
    for a, b in zip(g0, g1):
        assert torch.equal(a, b)                         # identical grads AND identical (broadcast) params
    n = len(g0) // 2
    torch.manual_seed(100)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 1))
    torch.manual_seed(7)
    x_all = torch.randn(8, 6); y_all = torch.randn(8, 1)
    ((net(x_all) - y_all) ** 2).mean().backward()        # mean of per-rank means == full-batch mean (equal shards)
    for p, g in zip(net.parameters(), g0[:n]):
        assert torch.allclose(p.grad, g, atol=1e-6)
