"""K3d parity: per-channel sums and affine+ReLU over boxes (mvsb200_channel_sums*, mvsb200_affine_relu_geo_*) against
plain torch expressions, forward and gradients, on strided channel-slice views like the ones the regulariser passes."""
import pytest
import torch
import torch.nn.functional as F

from mvs_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _slice_view(dtype, seed=0):
    """A channel slice (16 of 48 channels) and spatial crop of a bigger channels_last_3d tensor."""
    g = torch.Generator().manual_seed(seed)
    big = torch.randn(2, 48, 9, 8, 11, generator=g).to(DEV).to(dtype).contiguous(memory_format=torch.channels_last_3d)
    return big, (slice(None), slice(16, 32), slice(1, 8), slice(0, 7), slice(2, 11))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_channel_sums_and_gradient(dtype):
    big, sl = _slice_view(dtype)
    b1 = big.clone().requires_grad_(True)
    s1, s2 = ops.channel_sums(b1[sl])
    w1, w2 = torch.randn(16, device=DEV), torch.randn(16, device=DEV)
    ((s1 * w1).sum() + (s2 * w2).sum()).backward()
    b2 = big.float().requires_grad_(True)
    x = b2[sl]
    r1, r2 = x.sum((0, 2, 3, 4)), (x * x).sum((0, 2, 3, 4))
    ((r1 * w1).sum() + (r2 * w2).sum()).backward()
    assert torch.allclose(s1, r1, rtol=1e-5, atol=1e-3) and torch.allclose(s2, r2, rtol=1e-5, atol=1e-3)
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert (b1.grad.float() - b2.grad).abs().max().item() <= tol * float(b2.grad.abs().max())
    s1b, s2b = ops.channel_sums(big[sl])
    assert torch.equal(s1, s1b) and torch.equal(s2, s2b)            # deterministic


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("mode", ["grow", "crop", "shifted"])
def test_affine_relu_geo_matches_torch(dtype, mode):
    big, sl = _slice_view(dtype, seed=3)
    xin = big[sl]                                                     # input box 7 x 7 x 9 at frame origin (5, 6, 7)
    io = (5, 6, 7)
    if mode == "grow":      # output box = input box dilated by 2 (data on the box, BatchNorm'd zero around it)
        oo, od = (3, 4, 5), (11, 11, 13)
    elif mode == "crop":    # output box inside the input box
        oo, od = (6, 7, 9), (5, 4, 6)
    else:                   # partial overlap
        oo, od = (8, 2, 7), (6, 9, 5)
    g = torch.Generator().manual_seed(5)
    scale = (torch.rand(16, generator=g) + 0.5).to(DEV).requires_grad_(True)
    shift = (torch.randn(16, generator=g) * 0.5).to(DEV).requires_grad_(True)
    gy = torch.randn((2, 16) + od, generator=g).to(DEV).to(dtype)

    b1 = big.clone().requires_grad_(True)
    y = ops.affine_relu_geo(b1[sl], scale, shift, io, oo, od)
    y.backward(gy)
    g_scale, g_shift = scale.grad.clone(), shift.grad.clone()
    scale.grad = shift.grad = None

    b2 = big.float().requires_grad_(True)
    x = b2[sl]
    # embed the input box in a frame, cut the output box out of it
    lo = [min(a, b) for a, b in zip(io, oo)]
    hi = [max(a + n, b + m) for a, n, b, m in zip(io, xin.shape[2:], oo, od)]
    pad = []
    for ax in (2, 1, 0):
        pad += [io[ax] - lo[ax], hi[ax] - (io[ax] + xin.shape[2 + ax])]
    frame = F.pad(x, pad)
    cut = (slice(None), slice(None)) + tuple(slice(oo[ax] - lo[ax], oo[ax] - lo[ax] + od[ax]) for ax in range(3))
    yr = F.relu(frame[cut] * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1))
    yr.backward(gy.float())
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    assert y.shape == yr.shape and y.is_contiguous(memory_format=torch.channels_last_3d)
    assert (y.float() - yr).abs().max().item() <= tol * max(1.0, float(yr.abs().max()))
    assert (b1.grad.float() - b2.grad).abs().max().item() <= tol * max(1.0, float(b2.grad.abs().max()))
    assert torch.allclose(g_scale, scale.grad, rtol=1e-3, atol=1e-3 * float(scale.grad.abs().max()))
    assert torch.allclose(g_shift, shift.grad, rtol=1e-3, atol=1e-3 * float(shift.grad.abs().max()))
