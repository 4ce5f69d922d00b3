"""Depth-slab sharding on CPU: the plane arithmetic by brute force, and world_size 2/3/4 gloo runs of the
sharded regulariser against (a) the reference's golden logits and (b) the single-process module on larger volumes.
The conv backend here is torch's CPU conv and the per-slab arithmetic is the torch branch of depth_slab.py; the product
entry point (DepthSlabCostVolumeReg.forward) refuses CPU tensors."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


# ---------------------------------------------------------------------------------------------------------------------
# plan arithmetic (no communication)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("D,R", [(8, 2), (7, 2), (16, 4), (24, 3), (30, 4), (192, 8), (256, 8), (256, 4), (250, 8), (64, 5)])
def test_slab_plan_covers_every_dependency(D, R):
    from mvs_b200.depth_slab import SlabPlan
    pl = SlabPlan(D, R)
    p = D // 2 + 1
    assert pl.L == p - 2 * pl.lo
    # partitions
    for part, lo, hi in ((pl.canvas, 0, D), (pl.box, 0, pl.nC), (pl.tbox, pl.e0, pl.e1)):
        assert part[0][0] == lo and part[-1][1] == hi
        assert all(a[1] == b[0] for a, b in zip(part, part[1:])) and all(b > a for a, b in part)
    pcs = [pl.box_piece(r) for r in range(R)]
    assert pcs[0][0] == 0 and pcs[-1][1] == pl.nC and all(a[1] == b[0] for a, b in zip(pcs, pcs[1:]))
    for r in range(R):
        a, b = pl.canvas[r]
        ja, jb = pl.box[r]
        k0, k1 = pl.cost_planes(r)
        assert 0 <= k0 <= max(0, a - 1) and min(D, b + 1) <= k1 <= D
        # stride-2 branch: box plane j (canvas o = lo + j) reads canvas planes 2o - p + k (model.py:104-110, config.py:20)
        c0, c1, g, n_out = pl.s2_input(r)
        assert k0 <= c0 and c1 <= k1 and n_out == jb - ja + g
        for jl in range(g, n_out):                      # local output plane -> global box plane
            j = ja - g + jl
            for k in range(3):
                src = 2 * (pl.lo + j) - p + k
                loc = 2 * jl - pl.L + k                 # what the local call reads
                assert src == c0 + loc
                if 0 <= src < D:
                    assert c0 <= src < c1              # real data is inside the local input
                else:
                    assert loc < 0 or loc >= c1 - c0   # the canvas border is the local zero padding
        # conv_k_1 halo
        ta, tb = pl.tbox[r]
        sa, sb = pl.s_halo(r)
        xa, xb, zlo, zhi = pl.x_range(r)
        assert (xb - xa) + zlo + zhi == (tb - ta) + 2
        for t in range(ta, tb):
            for k in (-1, 0, 1):
                j = t + k
                if 0 <= j < pl.nC:
                    assert sa <= j < sb
                on_canvas = 0 <= pl.lo + j < D
                assert on_canvas == (xa <= j < xb)
        # transposed conv: canvas plane o gets box plane j through tap k iff o = 2 (lo + j) - p + k  (ConvTranspose3d, pad p)
        ua, ub, L_loc = pl.up_input(r)
        assert L_loc in (1, 2)
        for o in range(a, b):
            for k in range(3):
                if (o + p - k) % 2 == 0:
                    j = (o + p - k) // 2 - pl.lo
                    if 0 <= j < pl.nC:
                        assert ua <= j < ub
                        assert o - a == 2 * (j - ua) - L_loc + k


# ---------------------------------------------------------------------------------------------------------------------
# gloo runs
# ---------------------------------------------------------------------------------------------------------------------
def _init(rank, world, port):
    import sys
    for p in (os.path.join(ROOT, "deep-multiview-depth-estimation_b200"), os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    return dist


def _reslab_worker(rank, world, port, out):
    dist = _init(rank, world, port)
    from mvs_b200.depth_slab import reslab, reshard_rows, gather_rows, row_partition, TorchDistComm
    comm = TorchDistComm()
    torch.manual_seed(3)
    full = torch.randn(2, 4, 11, 3, 5)                              # B = 2: exercises the staged (non-contiguous) path
    have = [(0, 4), (4, 4), (4, 11)][:world] if world == 3 else [(0, 5), (5, 11)]
    want = [(-1, 6), (3, 9), (8, 13)][:world] if world == 3 else [(-2, 7), (4, 12)]
    mine = full[:, :, have[rank][0]:have[rank][1]].contiguous(memory_format=torch.channels_last_3d)
    got = reslab(mine, have, want, comm, fill=7.0)
    w0, w1 = want[rank]
    exp = torch.full((2, 4, w1 - w0, 3, 5), 7.0)
    lo_, hi_ = max(w0, 0), min(w1, 11)
    exp[:, :, lo_ - w0:hi_ - w0] = full[:, :, lo_:hi_]
    ok = torch.equal(got, exp) and got.is_contiguous(memory_format=torch.channels_last_3d)
    # plane slabs -> row slabs -> gathered map
    vol = torch.randn(1, 1, 11, 7, 5)
    planes = have
    rows = row_partition(7, world)
    lr = reshard_rows(vol[:, :, planes[rank][0]:planes[rank][1]].contiguous(), planes, rows, comm)
    ok = ok and torch.equal(lr, vol[:, :, :, rows[rank][0]:rows[rank][1]])
    g = gather_rows(lr.sum(2), rows, comm)
    ok = ok and torch.allclose(g, vol.sum(2))
    out[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [3])
def test_reslab_and_row_reshard(world):
    port = _free_port()
    mgr = mp.Manager(); out = mgr.dict()
    mp.spawn(_reslab_worker, args=(world, port, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))


def _reg(golden_dir, train=True):
    import mvs_b200
    reg = mvs_b200.CostVolumeReg(device="cpu", precision="fp32")
    w0 = np.load(os.path.join(golden_dir, "reg_weights.npz"))
    reg.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in w0.items()})
    return reg.train(train)


def _logits_worker(rank, world, port, cost, golden_dir, train, out):
    dist = _init(rank, world, port)
    from mvs_b200 import conv3d as conv_backends
    from mvs_b200.depth_slab import DepthSlabCostVolumeReg, SlabPlan, reshard_rows, row_partition
    cost = torch.from_numpy(cost)
    B, _, D, h, w = cost.shape
    reg = _reg(golden_dir, train)
    if not train:                                                   # non-trivial running statistics for eval mode
        g = torch.Generator().manual_seed(5)
        for bn in (reg.BN_0, reg.BN_1, reg.BN_2, reg.BN_3):
            bn.running_mean.copy_(0.1 * torch.randn(bn.num_features, generator=g))
            bn.running_var.copy_(0.5 + torch.rand(bn.num_features, generator=g))
    sharded = DepthSlabCostVolumeReg(reg)
    asked = []

    def cost_fn(c0, c1):
        asked.append((c0, c1))
        return cost[:, :, c0:c1].clone()

    lg = sharded.slab_logits(cost_fn, B, D, h, w, conv_backends.get("cudnn"))
    planes, rows = SlabPlan(D, world).canvas, row_partition(h, world)
    lr = reshard_rows(lg, planes, rows, sharded.comm)
    out[rank] = (lg.numpy(), lr.numpy(), asked, {k: v.numpy().copy() for k, v in reg.state_dict().items() if "running" in k})
    dist.destroy_process_group()


def _run_sharded(world, cost, golden_dir, train=True):
    port = _free_port()
    mgr = mp.Manager(); out = mgr.dict()
    mp.spawn(_logits_worker, args=(world, port, cost, golden_dir, train, out), nprocs=world, join=True)
    res = [out[r] for r in range(world)]
    logits = np.concatenate([r[0] for r in res], 2)
    by_rows = np.concatenate([r[1] for r in res], 3)
    return logits, by_rows, res


def _relmax(a, b):
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max() / np.abs(b).max())


@pytest.mark.parametrize("name,world", [("tiny_b1v3", 2), ("b2v3", 2), ("v7_odd", 2)])
def test_sharded_logits_match_reference_golden(golden_dir, name, world):
    g = dict(np.load(os.path.join(golden_dir, name + ".npz")))
    logits, by_rows, res = _run_sharded(world, g["cost"], golden_dir)
    assert logits.shape == g["logits"].shape
    assert _relmax(logits, g["logits"]) < 1e-4                      # north_star tolerance, fp32 path
    assert np.array_equal(by_rows, logits)                          # plane slabs -> row slabs is a pure permutation
    # every rank ends the pass with the same running statistics as the reference run
    for k, v in res[0][3].items():
        assert np.array_equal(v, res[1][3][k])
        assert np.allclose(v, g["bn_after/" + k], rtol=1e-4, atol=1e-6), k


@pytest.mark.parametrize("D,h,w,world,train", [(24, 12, 14, 3, True), (30, 9, 11, 4, True), (24, 12, 14, 2, False)])
def test_sharded_logits_match_single_process(golden_dir, D, h, w, world, train):
    from mvs_b200 import conv3d as conv_backends
    torch.manual_seed(11)
    cost = (torch.rand(1, 32, D, h, w) * torch.rand(1, 32, D, h, w)).numpy()          # variance-like: non-negative
    reg = _reg(golden_dir, train)
    if not train:
        g = torch.Generator().manual_seed(5)
        for bn in (reg.BN_0, reg.BN_1, reg.BN_2, reg.BN_3):
            bn.running_mean.copy_(0.1 * torch.randn(bn.num_features, generator=g))
            bn.running_var.copy_(0.5 + torch.rand(bn.num_features, generator=g))
    with torch.no_grad():
        ref = reg.logits(torch.from_numpy(cost), conv_backends.get("cudnn")).numpy()
    logits, by_rows, res = _run_sharded(world, cost, golden_dir, train)
    assert _relmax(logits, ref) < 2e-5
    assert np.array_equal(by_rows, logits)
    from mvs_b200.depth_slab import SlabPlan
    pl = SlabPlan(D, world)
    for r in range(world):                                           # each rank swept only its own planes + halo
        assert res[r][2] == [pl.cost_planes(r)]
        a, b = pl.canvas[r]
        assert pl.cost_planes(r)[1] - pl.cost_planes(r)[0] <= (b - a) + 3
    if train:
        for k, v in res[0][3].items():
            assert np.allclose(v, reg.state_dict()[k].numpy(), rtol=1e-4, atol=1e-6), k
