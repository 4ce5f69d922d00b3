"""K3b parity: fused train-mode BatchNorm3d+ReLU kernels (C ABI: mvsb200_bn_*) against torch's batch_norm on the same
seeded inputs, forward and backward, fp32 and bf16 storage, every channel count the regulariser uses, ragged row
counts; plus the whole regulariser's gradients against the goldens produced by the unmodified reference."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import mvs_b200
from mvs_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ref(x, w, b, eps, relu):
    y = F.batch_norm(x, None, None, w, b, True, 0.1, eps)
    return F.relu(y) if relu else y


@pytest.mark.parametrize("C", [8, 16, 32, 64])
@pytest.mark.parametrize("shape", [(1, 5, 7, 9), (2, 8, 16, 20), (1, 1, 1, 3)])
@pytest.mark.parametrize("relu", [True, False])
def test_bn_relu_fp32_forward_backward(C, shape, relu):
    B, D, h, w = shape
    g = torch.Generator().manual_seed(C * 100 + D)
    x = (torch.randn(B, C, D, h, w, generator=g) * 2 + 0.5).to(DEV).contiguous(memory_format=torch.channels_last_3d)
    wt = (torch.rand(C, generator=g) + 0.5).to(DEV).requires_grad_(True)
    bs = torch.randn(C, generator=g).to(DEV).requires_grad_(True)
    gy = torch.randn(B, C, D, h, w, generator=g).to(DEV)
    x1 = x.clone().requires_grad_(True)
    y, mean, var = ops.batchnorm_relu_train(x1, wt, bs, 1e-5, relu)
    y.backward(gy)
    x2 = x.clone().requires_grad_(True)
    w2, b2 = wt.detach().clone().requires_grad_(True), bs.detach().clone().requires_grad_(True)
    yr = _ref(x2, w2, b2, 1e-5, relu)
    yr.backward(gy)
    assert y.shape == yr.shape and y.is_contiguous(memory_format=torch.channels_last_3d)
    assert torch.allclose(mean, x.mean((0, 2, 3, 4)), rtol=1e-5, atol=1e-6)
    assert torch.allclose(var, x.var((0, 2, 3, 4), unbiased=False), rtol=1e-4, atol=1e-6)
    tol = dict(rtol=1e-4, atol=1e-5)
    assert torch.allclose(y, yr, **tol)
    scale = max(1.0, float(x2.grad.abs().max()))
    assert (x1.grad - x2.grad).abs().max().item() < 2e-4 * scale
    assert torch.allclose(wt.grad, w2.grad, rtol=1e-3, atol=1e-3 * float(w2.grad.abs().max()))
    assert torch.allclose(bs.grad, b2.grad, rtol=1e-3, atol=1e-3 * float(b2.grad.abs().max()))


@pytest.mark.parametrize("C", [8, 32])
def test_bn_relu_bf16_storage(C):
    g = torch.Generator().manual_seed(C)
    x = torch.randn(2, C, 6, 10, 12, generator=g).to(DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
    wt, bs = (torch.rand(C, generator=g) + 0.5).to(DEV), torch.randn(C, generator=g).to(DEV)
    gy = torch.randn(2, C, 6, 10, 12, generator=g).to(DEV).to(torch.bfloat16)
    x1 = x.clone().requires_grad_(True)
    y, _, _ = ops.batchnorm_relu_train(x1, wt, bs)
    y.backward(gy)
    x2 = x.float().requires_grad_(True)
    yr = _ref(x2, wt, bs, 1e-5, True)
    yr.backward(gy.float())
    assert y.dtype == torch.bfloat16 and x1.grad.dtype == torch.bfloat16
    assert (y.float() - yr).abs().max().item() < 1e-2 * float(yr.abs().max())
    assert (x1.grad.float() - x2.grad).abs().max().item() < 1e-2 * float(x2.grad.abs().max())


def test_bn_full_canvas_statistics_are_deterministic():
    x = torch.randn(1, 8, 192, 128, 160, device=DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
    wt, bs = torch.ones(8, device=DEV), torch.zeros(8, device=DEV)
    _, m1, v1 = ops.batchnorm_relu_train(x, wt, bs)
    _, m2, v2 = ops.batchnorm_relu_train(x, wt, bs)
    assert torch.equal(m1, m2) and torch.equal(v1, v2)
    assert torch.allclose(m1, x.float().mean((0, 2, 3, 4)), atol=1e-5)
    assert torch.allclose(v1, x.float().var((0, 2, 3, 4), unbiased=False), rtol=1e-4)


def test_bad_channel_count_raises():
    x = torch.zeros(1, 12, 2, 2, 2, device=DEV)
    with pytest.raises(mvs_b200.MvsB200Error, match="C must be"):
        ops.batchnorm_relu_train(x, torch.ones(12, device=DEV), torch.zeros(12, device=DEV))


def test_regulariser_gradients_match_reference(golden_dir):
    """Whole CostVolumeReg on the GPU (fused BN kernels + convs + K4) against the reference's autograd."""
    g = dict(np.load(os.path.join(golden_dir, "tiny_b1v3.npz")))
    reg = mvs_b200.CostVolumeReg(device=DEV, precision="fp32")       # owns its precision: no TF32 flag flip needed
    w0 = np.load(os.path.join(golden_dir, "reg_weights.npz"))
    reg.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in w0.items()})
    reg.train()
    cv = torch.from_numpy(g["cost"]).to(DEV).requires_grad_(True)
    prob = reg(cv)
    depth = mvs_b200.extract_depth_map(prob, torch.from_numpy(g["d_batch"]).to(DEV))
    names = [n for n, _ in reg.named_parameters()]
    grads = torch.autograd.grad((depth * torch.from_numpy(g["gdepth"]).to(DEV)).sum(), [cv] + list(reg.parameters()))
    ref = g["gcv"]
    assert np.abs(grads[0].cpu().numpy() - ref).max() < 5e-4 * np.abs(ref).max()
    for n, gr in zip(names, grads[1:]):
        ref = g["gparam/" + n]
        assert np.abs(gr.cpu().numpy() - ref).max() <= 5e-4 * max(np.abs(ref).max(), 1e-3), n


@pytest.mark.parametrize("C,dtype", [(16, torch.float32), (32, torch.bfloat16)])
def test_cropped_output_equals_dense_then_slice(C, dtype):
    """Crop-aware BN: full-volume statistics, result and incoming gradient only on a box."""
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, C, 9, 11, 13, generator=g).to(DEV).to(dtype).contiguous(memory_format=torch.channels_last_3d)
    wt, bs = (torch.rand(C, generator=g) + 0.5).to(DEV).requires_grad_(True), torch.randn(C, generator=g).to(DEV).requires_grad_(True)
    crop = ((2, 7), (1, 9), (3, 12))
    sl = (slice(None), slice(None), slice(2, 7), slice(1, 9), slice(3, 12))
    gy = torch.randn(2, C, 5, 8, 9, generator=g).to(DEV).to(dtype)
    x1 = x.clone().requires_grad_(True)
    y1, m1, v1 = ops.batchnorm_relu_train(x1, wt, bs, crop=crop)
    y1.backward(gy)
    gw1, gb1 = wt.grad.clone(), bs.grad.clone()
    wt.grad = bs.grad = None
    x2 = x.clone().requires_grad_(True)
    y2, m2, v2 = ops.batchnorm_relu_train(x2, wt, bs)
    y2[sl].backward(gy)
    assert torch.equal(m1, m2) and torch.equal(v1, v2)
    assert torch.equal(y1, y2[sl])
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert (x1.grad.float() - x2.grad.float()).abs().max().item() <= tol * max(1.0, float(x2.grad.float().abs().max()))
    assert torch.allclose(gw1, wt.grad, rtol=1e-4, atol=1e-4) and torch.allclose(gb1, bs.grad, rtol=1e-4, atol=1e-4)


def test_canvas_inside_larger_allocation():
    """The un-cropped transposed conv returns the canvas plus one slack plane/line/column: statistics and results must be
    those of the canvas alone, and the gradient into the slack must be zero."""
    g = torch.Generator().manual_seed(11)
    C = 16
    xa = torch.randn(2, C, 9, 7, 11, generator=g).to(DEV).contiguous(memory_format=torch.channels_last_3d)
    canvas = (8, 6, 10)
    wt, bs = (torch.rand(C, generator=g) + 0.5).to(DEV), torch.randn(C, generator=g).to(DEV)
    for crop in (None, ((1, 7), (2, 5), (0, 9))):
        sl = (slice(None), slice(None)) + tuple(slice(a, b) for a, b in (crop or tuple((0, n) for n in canvas)))
        x1 = xa.clone().requires_grad_(True)
        y1, m1, v1 = ops.batchnorm_relu_train(x1, wt, bs, crop=crop, canvas=canvas)
        gy = torch.randn(y1.shape, generator=torch.Generator().manual_seed(3)).to(DEV)
        y1.backward(gy)
        x2 = xa.clone().requires_grad_(True)
        xc = x2[..., :8, :6, :10].contiguous(memory_format=torch.channels_last_3d)
        y2, m2, v2 = ops.batchnorm_relu_train(xc, wt, bs)
        y2[sl].backward(gy)
        assert torch.allclose(m1, m2, atol=1e-6) and torch.allclose(v1, v2, rtol=1e-5)
        assert torch.allclose(y1, y2[sl], atol=1e-5)
        assert torch.allclose(x1.grad, x2.grad, atol=1e-5)
        assert x1.grad[..., 8:, :, :].abs().max() == 0 and x1.grad[..., :, 6:, :].abs().max() == 0


@pytest.mark.parametrize("C", [8, 32])
@pytest.mark.parametrize("geometry", ["plain", "canvas+crop"])
def test_running_statistics_updated_by_the_finalize_launch(C, geometry):
    """mvsb200_bn_stats_affine: the statistics' finalize launch also performs the running-statistics update of torch.nn.BatchNorm3d
    in train mode (momentum, unbiased variance, num_batches_tracked) -- against nn.BatchNorm3d on the same canvas, two passes."""
    g = torch.Generator().manual_seed(7 * C)
    canvas = (6, 9, 10)
    alloc = canvas if geometry == "plain" else (7, 10, 11)
    ref = torch.nn.BatchNorm3d(C, momentum=0.1).to(DEV).train()
    with torch.no_grad():
        ref.weight.copy_(torch.rand(C, generator=g) + 0.5)
        ref.bias.copy_(torch.randn(C, generator=g))
    rm, rv = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    nbt = torch.zeros((), dtype=torch.int64, device=DEV)
    for _ in range(2):
        x = (torch.randn(2, C, *alloc, generator=g) * 1.5 + 0.3).to(DEV).contiguous(memory_format=torch.channels_last_3d)
        xc = x[..., :canvas[0], :canvas[1], :canvas[2]]
        yr = F.relu(ref(xc.contiguous()))
        if geometry == "plain":
            y, mean, var = ops.batchnorm_relu_train(x, ref.weight, ref.bias, ref.eps, running=(rm, rv, nbt), momentum=0.1)
        else:
            crop = ((1, 5), (2, 8), (0, 10))
            y, mean, var = ops.batchnorm_relu_train(x, ref.weight, ref.bias, ref.eps, crop=crop, canvas=canvas, running=(rm, rv, nbt),
                                                    momentum=0.1)
            yr = yr[..., 1:5, 2:8, 0:10]
        assert torch.allclose(y, yr, rtol=1e-4, atol=1e-5)
        assert torch.allclose(mean, xc.mean((0, 2, 3, 4)), rtol=1e-5, atol=1e-6)
    assert int(nbt) == 2 == int(ref.num_batches_tracked)
    assert torch.allclose(rm, ref.running_mean, rtol=1e-5, atol=1e-6)
    assert torch.allclose(rv, ref.running_var, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("crop", [None, ((1, 6), (2, 9), (0, 11))])
def test_skip_addition_rides_on_the_apply_pass(dtype, crop):
    """SURVEY §8b bn_relu_add_apply (model.py:117-123): y = ReLU(BatchNorm(x)) + add in one pass -- bit-identical to adding the
    stored tensor afterwards; the addend's gradient is the output's."""
    C, B, D, h, w = 16, 2, 7, 10, 12
    g = torch.Generator().manual_seed(31)
    x = torch.randn(B, C, D, h, w, generator=g).to(DEV).to(dtype).contiguous(memory_format=torch.channels_last_3d)
    wt = (torch.rand(C, generator=g) + 0.5).to(DEV)
    bs = torch.randn(C, generator=g).to(DEV)
    out_dims = (D, h, w) if crop is None else tuple(b - a for a, b in crop)
    add = torch.randn(B, C, *out_dims, generator=g).to(DEV).to(dtype).contiguous(memory_format=torch.channels_last_3d)
    x1, a1 = x.clone().requires_grad_(True), add.clone().requires_grad_(True)
    y1, _, _ = ops.batchnorm_relu_train(x1, wt, bs, crop=crop, add=a1)
    x2, a2 = x.clone().requires_grad_(True), add.clone().requires_grad_(True)
    y2 = ops.batchnorm_relu_train(x2, wt, bs, crop=crop)[0] + a2
    assert y1.shape == y2.shape and torch.equal(y1, y2)
    gy = torch.randn(y1.shape, generator=g).to(DEV).to(dtype)
    y1.backward(gy)
    y2.backward(gy)
    assert torch.equal(x1.grad, x2.grad) and torch.equal(a1.grad, a2.grad) and torch.equal(a1.grad, gy)
    with pytest.raises(mvs_b200.MvsB200Error):
        ops.batchnorm_relu_train(x, wt, bs, crop=crop, add=add[:, :8])


@pytest.mark.parametrize("C", [16, 32, 64])
def test_outside_sums_match_the_torch_expressions(C):
    """mvsb200_outside_sums_{fwd,bwd} (one launch forward, two backward) against the einsum / fp64 expressions they replace
    (regulariser._outside_classes), values and gradients."""
    from mvs_b200.regulariser import CostVolumeReg
    g = torch.Generator().manual_seed(C)
    dims, E_lo, E_hi, B = (24, 16, 20), [5, 0, 4], [17, 11, 19], 2          # one axis touches both canvas borders
    W1 = (torch.randn(C, C, 3, 3, 3, generator=g) / 20).to(DEV).requires_grad_(True)
    bg1 = torch.rand(C, generator=g).to(DEV).requires_grad_(True)
    cnt = CostVolumeReg._outside_geometry(dims, E_lo, E_hi, B, torch.device(DEV))[1]
    A1, A2 = ops.outside_sums(W1, bg1, cnt)
    W2, bg2 = W1.detach().clone().requires_grad_(True), bg1.detach().clone().requires_grad_(True)
    val, cnt2 = CostVolumeReg._outside_classes(W2.float(), bg2, dims, E_lo, E_hi, B)
    vc = val.double() * cnt2.double()
    R1, R2 = vc.sum((1, 2, 3)), (vc * val).sum((1, 2, 3))
    assert A1.dtype == torch.float64 and torch.allclose(A1, R1, rtol=1e-5, atol=1e-6 * float(R1.abs().max()))
    assert torch.allclose(A2, R2, rtol=1e-5, atol=1e-6 * float(R2.abs().max()))
    g1 = torch.randn(C, generator=g, dtype=torch.float64).to(DEV)
    g2 = torch.randn(C, generator=g, dtype=torch.float64).to(DEV) * 1e-3
    (A1 * g1).sum().backward(retain_graph=True); (A2 * g2).sum().backward()
    ((R1 * g1).sum() + (R2 * g2).sum()).backward()
    assert torch.allclose(W1.grad, W2.grad, rtol=1e-4, atol=1e-5 * float(W2.grad.abs().max()))
    assert torch.allclose(bg1.grad, bg2.grad, rtol=1e-4, atol=1e-5 * float(bg2.grad.abs().max()))
