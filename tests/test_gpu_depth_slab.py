"""Depth-slab path on the GPU: R virtual ranks driven as R threads on ONE device through an in-process
comm (the driver's GPU box has one GPU; the NCCL run of the same code is tools/run_depth_slab.py under torchrun).
Every slab runs the libmvs_b200.so kernels (K1 on its planes + halo, tcgen05 convolutions, K3d statistics / affine,
K3c, K4); the result is compared with the single-GPU module on the same inputs and with the CPU oracle."""
import threading

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class ThreadComm:
    """depth_slab.TorchDistComm's interface for R threads of one process sharing one CUDA device and one stream."""

    class Shared:
        def __init__(self, world):
            self.world = world
            self.barrier = threading.Barrier(world)
            self.turn = threading.Lock()       # held by the one rank that is computing (ranks share the process-wide
            self.box = {}                      # library workspaces, allocator pools and module caches: they take turns)
            self.slots = [None] * world
            self.lock = threading.Lock()

    def __init__(self, shared, rank):
        self.sh, self.rank, self.world = shared, rank, shared.world

    def _wait(self):
        self.sh.turn.release()
        try:
            self.sh.barrier.wait()
        finally:
            self.sh.turn.acquire()

    def exchange(self, sends, recvs):
        with self.sh.lock:
            for d, t in sends:
                assert (self.rank, d) not in self.sh.box
                self.sh.box[(self.rank, d)] = t
        self._wait()
        for s, buf in recvs:
            with self.sh.lock:
                t = self.sh.box.pop((s, self.rank))
            assert t.shape == buf.shape and t.dtype == buf.dtype
            buf.copy_(t)
        self._wait()

    def all_gather(self, t):
        self.sh.slots[self.rank] = t
        self._wait()
        parts = [p.clone() for p in self.sh.slots]
        self._wait()
        return parts

    def all_reduce_sum(self, t):
        parts = self.all_gather(t.clone())
        t.copy_(torch.stack(parts).sum(0))
        return t

    def broadcast(self, t, src):
        parts = self.all_gather(t)
        if self.rank != src:
            t.copy_(parts[src])
        return t


def _run_ranks(world, fn):
    """R virtual ranks as R threads of this process.  Ranks share what real ranks (separate processes) do not -- the library's
    per-device workspaces, the allocator, module-level caches -- so they take turns: a rank computes while it holds the turn
    and hands it over only while it waits for the others inside an exchange."""
    shared = ThreadComm.Shared(world)
    out, err = [None] * world, []

    def work(r):
        shared.turn.acquire()
        try:
            torch.cuda.set_device(0)
            out[r] = fn(ThreadComm(shared, r))
        except BaseException as e:      # noqa: BLE001 -- re-raised below; a dead rank must not leave the others at a barrier
            err.append(e)
            shared.barrier.abort()
        finally:
            shared.turn.release()

    ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    if err:
        raise err[0]
    return out


def _setup(B, V, D, h, w, precision, train, seed=0):
    import mvs_b200
    import plane_sweep as ps
    dev = "cuda:0"
    torch.manual_seed(seed)
    K, R, T = ps.synthetic_cameras(B, V, h, w)
    d_min, d_int = torch.full((B, 1, 1, 1), 425.0), torch.ones(B, 1, 1, 1)
    feat = torch.randn(B * V, 32, h, w, device=dev)
    reg = mvs_b200.CostVolumeReg(device=dev, precision=precision).train(train)
    if not train:
        g = torch.Generator().manual_seed(5)
        for bn in (reg.BN_0, reg.BN_1, reg.BN_2, reg.BN_3):
            bn.running_mean.copy_(0.05 * torch.randn(bn.num_features, generator=g))
            bn.running_var.copy_(0.5 + torch.rand(bn.num_features, generator=g))
    sweep = mvs_b200.PlaneSweep(K, R, T, d_min, d_int, B, V, D, 480.0 / D, h, w, torch.device(dev))
    return reg, feat, sweep


@pytest.mark.parametrize("B,V,D,h,w,world,precision,train", [
    (1, 3, 32, 24, 40, 2, "bf16", True),
    (1, 5, 64, 40, 56, 4, "bf16", True),
    (1, 3, 48, 37, 50, 3, "bf16", False),
    (2, 3, 32, 24, 40, 2, "bf16", True),
    (1, 3, 32, 24, 40, 4, "fp32", True),
])
def test_depth_slab_equals_single_gpu(B, V, D, h, w, world, precision, train, monkeypatch):
    import copy
    import mvs_b200
    from mvs_b200 import ops
    from mvs_b200.depth_slab import DepthSlabCostVolumeReg, slab_cost_fn, SlabPlan
    reg, feat, sweep = _setup(B, V, D, h, w, precision, train)
    vol_dtype = torch.bfloat16 if precision == "bf16" else torch.float32
    regs = [copy.deepcopy(reg) for _ in range(world)]                 # one replica per rank, like one process per GPU
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)    # fp32 case: the library convs in true fp32
    # the unsharded reference takes its BatchNorm statistics from the stored canvases, as the slabs do (their sums are all-reduced):
    # the single-GPU default -- sums of the fp32 accumulators in the transposed convolution's epilogue -- differs by the rounding of
    # the stores, which is enough to swap near-tied planes of a random-init network in the depth comparison below
    monkeypatch.setenv("MVSB200_DECONV_STATS", "0")
    with torch.no_grad():
        cost = ops.warp_variance(feat, sweep, vol_dtype)
        ref_logits = reg.logits(cost, mvs_b200.conv3d.get(reg.conv_backend))
        ref_prob, ref_depth = ops.softmax_depth(ref_logits, sweep.d_batch_dev, 5)
    n0 = mvs_b200.launch_count()

    def rank_fn(comm):
        sharded = DepthSlabCostVolumeReg(regs[comm.rank], comm)
        depth, prob_rows, rows = sharded.forward(slab_cost_fn(feat, sweep, vol_dtype), sweep.d_batch_dev, B, D, h, w)
        torch.cuda.synchronize()
        return sharded.last_logits, depth, prob_rows, rows

    outs = _run_ranks(world, rank_fn)
    assert mvs_b200.launch_count() > n0                                # the slabs ran libmvs_b200.so kernels
    logits = torch.cat([o[0] for o in outs], 2)
    tol = 2e-2 if precision == "bf16" else 1e-4                        # bf16 convs: statistics summed in another order
    scale = float(ref_logits.abs().max())
    assert float((logits - ref_logits).abs().max()) < tol * scale
    prob = torch.cat([o[2] for o in outs], 3)
    assert float((prob - ref_prob).abs().max()) < tol * float(ref_prob.abs().max())
    step = 480.0 / D
    for o in outs:                                                     # every rank holds the full depth map
        assert torch.equal(o[1], outs[0][1])
    derr = (outs[0][1] - ref_depth).abs()
    if precision == "fp32":
        assert float(derr.max()) < 0.005 * step
    else:                                                              # bf16 logits may swap near-tied planes: bulk agreement
        assert float((derr < 0.05 * step).float().mean()) > 0.96       # (a rank swap moves a kept plane: chaotic in the rounding)
    if train:                                                          # identical running statistics on every replica
        for k, v in regs[0].state_dict().items():
            assert torch.equal(v, regs[-1].state_dict()[k]), k
        for k, v in reg.state_dict().items():
            if "running" in k:
                assert torch.allclose(regs[0].state_dict()[k], v, rtol=2e-2 if precision == "bf16" else 1e-4, atol=1e-5), k


@pytest.mark.parametrize("name", ["tiny_b1v3", "b2v3", "v7_odd"])
def test_depth_slab_fp32_matches_reference_golden(golden_dir, name, monkeypatch):
    """fp32 slab path (2 slabs) on the golden inputs against the outputs of the unmodified reference (tests/golden/*.npz,
    oracle/make_golden.py): probability volume <= 1e-4 relative, depth within 0.5 % of the depth interval off tie pixels."""
    import copy
    import os
    import mvs_b200
    import plane_sweep as ps
    from mvs_b200.depth_slab import DepthSlabCostVolumeReg, slab_cost_fn
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    g = dict(np.load(os.path.join(golden_dir, name + ".npz")))
    t = lambda a: torch.from_numpy(np.asarray(a))
    B, V, D, world = int(g["B"]), int(g["V"]), int(g["D"]), 2
    feat = t(g["feat"]).to("cuda:0")
    h, w = feat.shape[-2:]
    reg = mvs_b200.CostVolumeReg(device="cuda:0", precision="fp32").train()
    w0 = np.load(os.path.join(golden_dir, "reg_weights.npz"))
    reg.load_state_dict({k: t(v).to("cuda:0") for k, v in w0.items()})
    regs = [copy.deepcopy(reg) for _ in range(world)]
    sweep = mvs_b200.PlaneSweep(t(g["K"]), t(g["R"]), t(g["T"]), t(g["d_min"]), t(g["d_int"]), B, V, D, int(g["d_scale"]),
                                h, w, torch.device("cuda:0"))

    def rank_fn(comm):
        sharded = DepthSlabCostVolumeReg(regs[comm.rank], comm)
        depth, prob_rows, rows = sharded.forward(slab_cost_fn(feat, sweep, torch.float32), sweep.d_batch_dev, B, D, h, w)
        torch.cuda.synchronize()
        return depth, prob_rows

    outs = _run_ranks(world, rank_fn)
    prob = torch.cat([o[1] for o in outs], 3).cpu().numpy()
    assert float(np.abs(prob - g["prob"]).max() / np.abs(g["prob"]).max()) < 1e-4
    ok = ~ps.tie_pixels(g["prob"])
    step = float(g["d_scale"]) * float(np.asarray(g["d_int"]).reshape(-1)[0])
    assert np.abs(outs[0][0].cpu().numpy()[:, 0] - g["depth"][:, 0])[ok].max() < 0.005 * step
    for k, v in regs[0].state_dict().items():                          # BatchNorm buffers after the pass == reference's
        if "running" in k:
            assert np.allclose(v.cpu().numpy(), g["bn_after/" + k], rtol=1e-4, atol=1e-6), k


def test_depth_slab_mvsnet_wrapper(monkeypatch):
    """harness.DepthSlabMVSNet: view-sharded encoding + broadcast reproduces the encoder's feature maps; the whole sharded
    forward (encode -> slab sweep -> slab regulariser -> depth -> refinement) yields the same maps on every rank.  (Depth maps
    of a random-init network are not compared with the unsharded run: its probabilities are nearly flat, so the rank-based
    extraction amplifies 1e-6 feature differences; with identical features the comparison is the test above.)"""
    import copy
    import plane_sweep as ps
    from mvs_b200.harness import MVSNet, DepthSlabMVSNet
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    V, D, H, W, world = 3, 32, 96, 160, 2
    torch.manual_seed(0)
    model = MVSNet(D, 480.0 / D, precision="fp32").to("cuda:0").eval()
    model.cost_volume_reg.train()                                      # batch statistics in the sharded part
    K, R, T = ps.synthetic_cameras(1, V, H // 4, W // 4)
    d_min, d_int = torch.full((1, 1, 1, 1), 425.0), torch.ones(1, 1, 1, 1)
    img = torch.randn(V, 3, H, W, device="cuda:0")
    models = [copy.deepcopy(model) for _ in range(world)]
    with torch.no_grad():
        ref_feats = model.feature_encoder(img)

    def rank_fn(comm):
        net = DepthSlabMVSNet(models[comm.rank], comm)
        feats = net.encode(img)                                        # eval-mode encoder: by view, broadcast
        out = net.forward(img, K, R, T, d_min, d_int, V)
        torch.cuda.synchronize()
        return feats, out

    outs = _run_ranks(world, rank_fn)
    for feats, (initial, refined) in outs:
        assert torch.allclose(feats, ref_feats, rtol=1e-4, atol=1e-5)
        assert initial.shape == (1, 1, H // 4, W // 4) and refined.shape == initial.shape
        assert torch.equal(initial, outs[0][1][0]) and torch.isfinite(refined).all()
        assert float(initial.min()) >= 425.0 - 1e-3 and float(initial.max()) <= 425.0 + 480.0


def test_graphed_slab_forward_replays():
    """depth_slab.GraphedSlabForward: the rank's forward captured as one CUDA graph gives the eager result, also after the
    features and the cameras change between replays (one rank here; the NCCL capture is exercised by tools/run_depth_slab.py
    --graph and bench.py --workload cfg4 --gpus N)."""
    import copy
    import mvs_b200
    import plane_sweep as ps
    from mvs_b200.depth_slab import DepthSlabCostVolumeReg, GraphedSlabForward, slab_cost_fn
    B, V, D, h, w = 1, 3, 32, 24, 40
    reg, feat, sweep = _setup(B, V, D, h, w, "bf16", True)
    twin = copy.deepcopy(reg)

    def rank_fn(comm):
        sharded = DepthSlabCostVolumeReg(reg, comm)
        eager = DepthSlabCostVolumeReg(twin, comm)
        graphed = GraphedSlabForward(sharded, sweep, tuple(feat.shape), feat.device, torch.bfloat16, warmup=1)
        outs = []
        for it in range(3):
            torch.manual_seed(10 + it)
            f = torch.randn_like(feat)
            K, R, T = ps.synthetic_cameras(B, V, h, w, seed=it)
            sweep.update(K, R, T, torch.full((B, 1, 1, 1), 425.0 + 10 * it), torch.ones(B, 1, 1, 1))
            n0 = mvs_b200.launch_count()
            depth, prob_rows, rows = graphed(f)
            torch.cuda.synchronize()
            if it > 0:
                assert mvs_b200.launch_count() == n0                 # a replay issues no C-ABI call
            d2, p2, _ = eager.forward(slab_cost_fn(f, sweep, torch.bfloat16), sweep.d_batch_dev, B, D, h, w)
            outs.append((depth.clone(), prob_rows.clone(), d2, p2))
        graphed.release()
        return outs

    (outs,) = _run_ranks(1, rank_fn)
    for depth, prob, d2, p2 in outs:
        assert float((prob - p2).abs().max()) < 2e-2 * float(p2.abs().max())
        assert float(((depth - d2).abs() < 0.05 * 480.0 / D).float().mean()) > 0.97
