"""harness.GraphedTrainStep: the train step replayed as one CUDA graph gives the loss and gradients of the eager step,
also after the inputs (images, ground truth, cameras, depth range) change between replays."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _batch(B, V, H, W, seed, d0):
    import plane_sweep as ps
    g = torch.Generator().manual_seed(seed)
    K, R, T = ps.synthetic_cameras(B, V, H // 4, W // 4, seed=seed)
    d_min, d_int = torch.full((B, 1, 1, 1), d0), torch.ones(B, 1, 1, 1)
    img = torch.randn(B * V, 3, H, W, generator=g).pin_memory()
    gt = (d0 + 480.0 * torch.rand(B, 1, H // 4, W // 4, generator=g))
    gt = (gt * (torch.rand(B, 1, H // 4, W // 4, generator=g) > 0.3)).pin_memory()
    return img, gt, K, R, T, d_min, d_int


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_graphed_step_equals_eager_step(precision):
    import mvs_b200
    from mvs_b200.harness import MVSNet, GraphedTrainStep, loss_fcn
    B, V, H, W, D = 2, 3, 96, 128, 16
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = MVSNet(D, 480.0 / D, precision=precision).to(dev).train()
    twin = copy.deepcopy(model)
    gstep = GraphedTrainStep(model, B, V, H, W, dev)
    for it, (seed, d0) in enumerate([(1, 425.0), (2, 500.0), (3, 425.0)]):
        img, gt, K, R, T, d_min, d_int = _batch(B, V, H, W, seed, d0)
        n0 = mvs_b200.launch_count()
        loss = gstep.run(img, gt, K, R, T, d_min, d_int)
        torch.cuda.synchronize()
        if it > 0:
            assert mvs_b200.launch_count() == n0           # a replay issues no C-ABI call: the launches are in the graph
        assert gstep.launches > 20
        twin.load_state_dict(model.state_dict()) if it == 0 else None
        for p in twin.parameters():
            p.grad = None
        # BatchNorm running statistics advance in both; parameters are never stepped here, so the two nets stay equal
        initial, refined = twin(img.to(dev), K, R, T, d_min, d_int, B, V)
        ref_loss, _, _ = loss_fcn(gt.to(dev), initial, refined)
        ref_loss.backward()
        tol = 3e-2 if precision == "bf16" else 2e-3
        assert abs(float(loss) - float(ref_loss)) <= tol * abs(float(ref_loss))
        worst = 0.0
        for (n, p), q in zip(model.named_parameters(), twin.parameters()):
            assert (p.grad is None) == (q.grad is None), n
            if p.grad is not None:
                worst = max(worst, float((p.grad - q.grad).abs().max() / (q.grad.abs().max() + 1e-12)))
        # the loss goes through the rank-based depth extraction (depthmap.py:11-15): reordered atomic sums can flip a rank, so the
        # two runs agree to a few percent, not to rounding
        assert worst < (0.15 if precision == "bf16" else 5e-2), worst


def test_graphed_inference_equals_eager_forward():
    """harness.GraphedInference: the no-grad forward replayed as one CUDA graph returns the eager forward's depth maps, also
    after images and cameras change between replays."""
    import mvs_b200
    from mvs_b200.harness import MVSNet, GraphedInference
    B, V, H, W, D = 1, 3, 96, 128, 16
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = MVSNet(D, 480.0 / D, precision="bf16").to(dev).train()      # train-mode BatchNorm also at test time (test.py:61)
    twin = copy.deepcopy(model)
    ginf = GraphedInference(model, B, V, H, W, dev, warmup=1)
    for it, (seed, d0) in enumerate([(1, 425.0), (2, 500.0), (3, 425.0)]):
        img, _, K, R, T, d_min, d_int = _batch(B, V, H, W, seed, d0)
        n0 = mvs_b200.launch_count()
        initial, refined = ginf.run(img, K, R, T, d_min, d_int)
        torch.cuda.synchronize()
        if it > 0:
            assert mvs_b200.launch_count() == n0
        with torch.no_grad():
            ref_i, ref_r = twin(img.to(dev), K, R, T, d_min, d_int, B, V)
        step = 480.0 / D
        assert float(((initial - ref_i).abs() < 0.05 * step).float().mean()) > 0.97
        assert torch.isfinite(refined).all() and refined.shape == ref_r.shape
