"""SURVEY §8 row f1 -- the 2D feature encoder (model.py:20-65) on the path's tensor-core kernels (mvs_b200/nets2d.py).

Golden: tests/golden/encoder.npz, produced by the UNMODIFIED reference on the CPU in fp32 (oracle/make_golden_encoder.py).

CPU: the harness's stock-torch encoder reproduces the golden; the host-side re-indexing that turns a 5x5 stride-2 convolution
into a 3x3 one on the space-to-depth form is exact.  GPU: layout kernels bit-exact against torch; single layers against torch's
convolution on bf16-rounded operands (forward, data gradient, weight gradient); the whole native encoder against the golden
within the bf16 convolution tolerance (1e-2 of the largest feature), gradients as close to the fp32 golden as the stock bf16
layers are."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import mvs_b200
from mvs_b200 import harness, nets2d

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "encoder.npz")
DEV = "cuda:0"


def _golden_net(g, device):
    net = harness.FeatureEncoder().to(device).train()
    net.load_state_dict({k[3:]: torch.from_numpy(np.asarray(g[k])) for k in g.files if k.startswith("w0.")})
    return net


def _s2d_torch(x):
    """[N, C, 2h, 2w] -> [N, 4C, h, w] with channel (py*2 + px)*C + c."""
    N, C, H, W = x.shape
    return x.view(N, C, H // 2, 2, W // 2, 2).permute(0, 3, 5, 1, 2, 4).reshape(N, 4 * C, H // 2, W // 2)


def _rel_l2(a, b):
    a, b = a.detach().float(), b.detach().float()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def test_stock_encoder_reproduces_the_reference():
    g = np.load(GOLD)
    net = _golden_net(g, "cpu")
    feats = net(torch.from_numpy(g["images"]))
    assert torch.allclose(feats, torch.from_numpy(g["features"]), rtol=1e-4, atol=1e-5)
    feats.backward(torch.from_numpy(g["g_features"]))
    for n, p in net.named_parameters():
        ref = torch.from_numpy(g["g." + n])
        assert (p.grad - ref).abs().max() <= 1e-3 * max(float(ref.abs().max()), 1e-3), n


@pytest.mark.parametrize("shape", [(2, 8, 16, 12, 20), (1, 16, 32, 8, 8)])
def test_k5s2_is_a_3x3_convolution_on_the_space_to_depth_form(shape):
    N, ci, co, H, W = shape
    gen = torch.Generator().manual_seed(ci)
    x = torch.randn(N, ci, H, W, generator=gen, dtype=torch.float64)
    w5 = torch.randn(co, ci, 5, 5, generator=gen, dtype=torch.float64, requires_grad=True)
    want = F.conv2d(x, w5, stride=2, padding=2)
    w3 = nets2d.k5s2_as_3x3(w5.float()).double()
    assert w3.shape == (co, 4 * ci, 3, 3)
    got = F.conv2d(_s2d_torch(x), w3, padding=1)
    assert torch.allclose(got, want.detach(), rtol=1e-5, atol=1e-5)                 # fp32 re-indexing of an fp64 filter
    assert int((w3 != 0).sum()) == co * ci * 25                                    # 25 of the 36 taps are the filter's


def test_c_abi_exports_the_encoder_kernels():
    from mvs_b200 import _lib
    lib = _lib.load()
    for name in ("mvsb200_image_to_rows8", "mvsb200_s2d_rows_bf16", "mvsb200_conv3d_s1_wgrad_ex"):
        assert hasattr(lib, name), name


@pytest.mark.gpu
def test_layout_kernels_bit_exact():
    gen = torch.Generator().manual_seed(3)
    img = torch.rand(3, 3, 12, 20, generator=gen).to(DEV)
    for images in (img, img.contiguous(memory_format=torch.channels_last), img[:, :, ::1, :]):
        rows = nets2d.image_rows(images)
        assert rows.shape == (1, 8, 3, 12, 20) and rows.is_contiguous(memory_format=torch.channels_last_3d)
        assert torch.equal(rows[0, :3].permute(1, 0, 2, 3), images.to(torch.bfloat16))
        assert torch.count_nonzero(rows[0, 3:]).item() == 0
    for C in (8, 16):
        x = torch.randn(1, C, 3, 12, 20, generator=gen).to(DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
        x.requires_grad_(True)
        y = nets2d.space_to_depth(x)
        want = _s2d_torch(x.detach()[0].permute(1, 0, 2, 3))                      # [N, 4C, h, w]
        assert y.shape == (1, 4 * C, 3, 6, 10) and torch.equal(y[0].permute(1, 0, 2, 3), want)
        gy = torch.randn(y.shape, generator=gen).to(DEV).to(torch.bfloat16)
        y.backward(gy)
        back = torch.zeros_like(x.detach().float())                               # adjoint of a permutation = its inverse
        back[0] = _s2d_inverse_torch(gy[0].permute(1, 0, 2, 3).float(), C).permute(1, 0, 2, 3)
        assert torch.equal(x.grad.float(), back)


def _s2d_inverse_torch(y, C):
    N, C4, h, w = y.shape
    return y.view(N, 2, 2, C, h, w).permute(0, 3, 4, 1, 5, 2).reshape(N, C, 2 * h, 2 * w)


@pytest.mark.gpu
@pytest.mark.parametrize("case", [(3, 8, 8, 8), (8, 8, 8, 8), (16, 16, 16, 16), (32, 32, 32, 32), (4, 32, 16, 32), (32, 1, 32, 8),
                                  (32, 16, 32, 16), (64, 32, 64, 32)])
@pytest.mark.parametrize("dims", [(3, 12, 20), (2, 33, 47)])
def test_conv3x3_rows_against_torch(case, dims):
    """One layer: forward, data gradient, weight gradient against torch.conv2d in fp32 on the bf16-rounded operands."""
    ci, co, cx, cy = case
    N, H, W = dims
    gen = torch.Generator().manual_seed(ci * 64 + co + H)
    x = torch.zeros(N, cx, H, W)
    x[:, :ci] = torch.randn(N, ci, H, W, generator=gen)
    x = x.to(torch.bfloat16)
    w = (torch.randn(co, ci, 3, 3, generator=gen) / (3.0 * ci ** 0.5)).to(DEV).requires_grad_(True)
    rows = x.to(DEV).permute(1, 0, 2, 3).unsqueeze(0).contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
    y = nets2d.conv3x3(rows, w, cy)
    assert y.shape == (1, cy, N, H, W) and y.dtype == torch.bfloat16
    gy = torch.zeros(N, cy, H, W)
    gy[:, :co] = torch.randn(N, co, H, W, generator=gen)
    gy = gy.to(torch.bfloat16)
    y.backward(gy.to(DEV).permute(1, 0, 2, 3).unsqueeze(0))
    xr = x.float().to(DEV)[:, :ci].requires_grad_(True)
    wr = w.detach().to(torch.bfloat16).float().requires_grad_(True)
    prev = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        yr = F.conv2d(xr, wr, padding=1)
        yr.backward(gy.float().to(DEV)[:, :co])
    finally:
        torch.backends.cudnn.allow_tf32 = prev
    got = y[0].permute(1, 0, 2, 3).float()
    assert (got[:, :co] - yr).abs().max().item() <= 1e-2 * max(1.0, float(yr.detach().abs().max()))       # bf16 rounding of the output
    assert torch.count_nonzero(got[:, co:]).item() == 0
    gx = rows.grad[0].permute(1, 0, 2, 3).float()
    assert (gx[:, :ci] - xr.grad).abs().max().item() <= 1e-2 * max(1.0, float(xr.grad.abs().max()))
    assert torch.count_nonzero(gx[:, ci:]).item() == 0
    assert w.grad.shape == w.shape
    assert (w.grad - wr.grad).abs().max().item() <= 2e-3 * max(1.0, float(wr.grad.abs().max()))  # fp32 accumulation, other order


@pytest.mark.gpu
def test_native_encoder_against_the_golden(monkeypatch):
    from mvs_b200 import ops
    g = np.load(GOLD)
    images = torch.from_numpy(g["images"]).to(DEV)
    want = torch.from_numpy(g["features"]).to(DEV)
    gf = torch.from_numpy(g["g_features"]).to(DEV)
    runs = {}
    for mode in ("native", "torch"):
        net = _golden_net(g, DEV)
        ops.EVENTS = {}
        try:
            if mode == "native":
                assert nets2d.encoder_ok(net, images)
                feats = nets2d.encode_native(net, images).float()
            else:
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    feats = net(images.contiguous(memory_format=torch.channels_last)).float()
            feats.backward(gf)
            torch.cuda.synchronize()
            launched = {k: len(v) for k, v in ops.EVENTS.items()}
        finally:
            ops.EVENTS = None
        runs[mode] = (feats.detach(), dict(net.named_parameters()), net.state_dict(), launched)
    feats, params, state, launched = runs["native"]
    # 8 convolutions forward, 7 data gradients (the images need none), 8 weight gradients; the space-to-depth permutations ride on
    # the BatchNorm passes
    assert launched.get("conv2d_tc") == 15 and launched.get("conv2d_wgrad_tc") == 8 and "s2d_rows" not in launched, launched
    assert feats.shape == want.shape and feats.is_contiguous(memory_format=torch.channels_last)
    err, err_stock = (feats - want).abs().max().item(), (runs["torch"][0] - want).abs().max().item()
    scale = max(1.0, float(want.abs().max()))
    assert err <= max(1e-2 * scale, 1.5 * err_stock), (err, err_stock, scale)
    for n, p in params.items():
        ref = torch.from_numpy(g["g." + n]).to(DEV)
        assert p.grad is not None and p.grad.shape == ref.shape, n
        e_nat, e_stock = _rel_l2(p.grad, ref), _rel_l2(runs["torch"][1][n].grad, ref)
        assert e_nat <= max(1.5 * e_stock, 2e-2), (n, e_nat, e_stock)
    for k in g.files:
        if k.startswith("w1."):
            got, exp = state[k[3:]].float().cpu(), torch.from_numpy(np.asarray(g[k])).float()
            assert torch.allclose(got, exp, rtol=2e-2, atol=2e-3), k


class _ReferenceLayout(torch.nn.Module):
    """The module layout of model.py:20-65 written with stock torch.nn (Conv2d / BatchNorm2d / ReLU triples)."""

    def __init__(self):
        super().__init__()
        nn = torch.nn
        layers = []
        for ci, co, k, s in nets2d._ENC_LAYERS[:-1]:
            layers += [nn.Conv2d(ci, co, k, stride=s, padding=k // 2, bias=False), nn.BatchNorm2d(co), nn.ReLU()]
        self.model = nn.Sequential(*layers, nn.Conv2d(32, 32, 3, padding=1, bias=False))

    def forward(self, x):
        return self.model(x)


@pytest.mark.gpu
def test_encode_features_binds_to_a_module_of_the_reference_layout():
    g = np.load(GOLD)
    ours = _golden_net(g, DEV)
    theirs = _ReferenceLayout().to(DEV).train()
    theirs.load_state_dict(ours.state_dict())
    images = torch.from_numpy(g["images"]).to(DEV)
    n0 = mvs_b200.launch_count()
    a = mvs_b200.encode_features(theirs, images)
    assert mvs_b200.launch_count() - n0 >= 8 + 8 + 7 * 2             # convolutions, filter packing, BatchNorm: the library ran
    b = mvs_b200.encode_features(ours, images)
    assert a.dtype == torch.bfloat16 and a.shape == (3, 32, 6, 10) and torch.equal(a, b)
    for (ka, va), (kb, vb) in zip(theirs.state_dict().items(), ours.state_dict().items()):
        assert ka == kb and torch.equal(va, vb), ka
    theirs.eval()
    n0 = mvs_b200.launch_count()
    c = mvs_b200.encode_features(theirs, images)                      # eval-mode BatchNorm: the module's own torch layers
    assert mvs_b200.launch_count() == n0 and c.shape == a.shape and torch.isfinite(c.float()).all()
    with pytest.raises(mvs_b200.MvsB200Error):
        nets2d.image_rows(images.cpu())                               # no CPU path


@pytest.mark.gpu
@pytest.mark.parametrize("C", [8, 16, 32])
def test_batchnorm_writes_the_space_to_depth_form(C):
    """BatchNorm + ReLU with s2d=True == space_to_depth(BatchNorm + ReLU), bit for bit, forward and backward."""
    from mvs_b200 import ops
    gen = torch.Generator().manual_seed(C)
    N, H, W = 3, 12, 20
    x = (torch.randn(1, C, N, H, W, generator=gen) * 1.5 + 0.3).to(DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
    wt = (torch.rand(C, generator=gen) + 0.5).to(DEV).requires_grad_(True)
    bs = torch.randn(C, generator=gen).to(DEV).requires_grad_(True)
    gy = torch.randn(1, 4 * C, N, H // 2, W // 2, generator=gen).to(DEV).to(torch.bfloat16)
    outs = []
    for fused in (True, False):
        x1 = x.clone().requires_grad_(True)
        w1, b1 = wt.detach().clone().requires_grad_(True), bs.detach().clone().requires_grad_(True)
        y = ops.batchnorm_relu_train(x1, w1, b1, s2d=fused)[0]
        if not fused:
            y = nets2d.space_to_depth(y)
        y.backward(gy)
        outs.append((y.detach(), x1.grad, w1.grad, b1.grad))
    for a, b_ in zip(*outs):
        assert a.shape == b_.shape and torch.equal(a, b_)
    with pytest.raises(mvs_b200.MvsB200Error):
        ops.batchnorm_relu_train(x[:, :, :, :11], wt, bs, s2d=True)          # odd map height


@pytest.mark.gpu
@pytest.mark.parametrize("case", [(12, 512, 640, True), (5, 1184, 1600, False)])
def test_native_encoder_matches_the_stock_layers_at_full_size(case):
    """BASELINE sizes: cfg2's 12 maps of 640x512 (forward + backward) and cfg4's 5 maps of 1600x1184 (forward), native against the
    stock bf16 layers (cuDNN) on the same weights: two bf16 evaluations of the same network agree to a few bf16 steps per layer."""
    N, H, W, with_grad = case
    torch.manual_seed(11)
    ours = harness.FeatureEncoder().to(DEV).train()
    theirs = harness.FeatureEncoder().to(DEV).train()
    theirs.load_state_dict(ours.state_dict())
    images = torch.rand(N, 3, H, W, device=DEV)
    assert nets2d.encoder_ok(ours, images)
    with torch.set_grad_enabled(with_grad):
        a = nets2d.encode_native(ours, images).float()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            b = theirs(images.contiguous(memory_format=torch.channels_last)).float()
    assert a.shape == b.shape == (N, 32, H // 4, W // 4)
    assert _rel_l2(a, b) <= 2e-2, _rel_l2(a, b)
    assert (a - b).abs().max().item() <= 6e-2 * float(b.abs().max())
    for (ka, va), (kb, vb) in zip(ours.state_dict().items(), theirs.state_dict().items()):
        if "running" in ka:
            assert torch.allclose(va, vb, rtol=2e-2, atol=2e-3), ka
    if with_grad:
        g = torch.randn_like(a)
        a.backward(g)
        b.backward(g)
        for (n, p), (_, q) in zip(ours.named_parameters(), theirs.named_parameters()):
            assert _rel_l2(p.grad, q.grad) <= 1e-1, (n, _rel_l2(p.grad, q.grad))
