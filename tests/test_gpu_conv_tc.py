"""K3 parity: the tcgen05/TMA implicit-GEMM convolution (C ABI: mvsb200_conv3d_s1_fwd, called through the tcgen05 conv
backend) against torch's fp32 convolution of the same bf16-rounded operands -- forward and data gradient, every channel
pairing the regulariser uses, padded and valid, ragged sizes, batch > 1.  Tolerance: BASELINE.json's 1e-2 relative for
the bf16 conv path (observed ~3e-3: one bf16 rounding of the fp32 accumulator)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import mvs_b200
from mvs_b200 import conv3d as conv_backends
from mvs_b200 import conv3d_sm100

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-2


def _case(B, cin, cout, D, h, w, seed=0):
    g = torch.Generator().manual_seed(seed + cin * 131 + cout * 7 + D)
    x = torch.randn(B, cin, D, h, w, generator=g).to(DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
    wt = (torch.randn(cout, cin, 3, 3, 3, generator=g) / (27 * cin) ** 0.5).to(DEV).to(torch.bfloat16)
    return x, wt


def _rel(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max())


@pytest.fixture(autouse=True, params=["kdn", "taps"])
def _no_tf32_both_s1_forms(request, monkeypatch):
    """Every test runs with both forms of the stride-1 kernel: "kdn" (depth tap folded into the MMA N extent, the default)
    and "taps" (one MMA per tap)."""
    monkeypatch.setenv("MVSB200_CONV_S1", request.param)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32 = old


@pytest.mark.parametrize("cin,cout", [(32, 8), (16, 16), (32, 32), (64, 64), (32, 16), (64, 32), (16, 32), (32, 24)])
@pytest.mark.parametrize("pad", [1, 0])
def test_forward_matches_fp32_conv(cin, cout, pad):
    x, wt = _case(1, cin, cout, 7, 21, 38)
    be = conv_backends.get("tcgen05")
    n0 = mvs_b200.launch_count()
    y = be.conv3d(x, wt, 1, (pad,) * 3)
    assert mvs_b200.launch_count() > n0, "the tcgen05 kernel did not run"
    ref = F.conv3d(x.float(), wt.float(), padding=pad)
    assert y.shape == ref.shape and y.dtype == torch.bfloat16
    assert y.is_contiguous(memory_format=torch.channels_last_3d)
    assert _rel(y, ref) < TOL


@pytest.mark.parametrize("shape", [(2, 5, 33, 47), (1, 3, 3, 3), (1, 9, 64, 83), (3, 4, 10, 130), (1, 12, 131, 9)])
def test_forward_ragged_shapes_and_batches(shape):
    B, D, h, w = shape
    x, wt = _case(B, 32, 32, D, h, w, seed=5)
    y = conv_backends.get("tcgen05").conv3d(x, wt, 1, (1, 1, 1))
    assert _rel(y, F.conv3d(x.float(), wt.float(), padding=1)) < TOL
    if min(D, h, w) >= 3:
        y0 = conv_backends.get("tcgen05").conv3d(x, wt, 1, (0, 0, 0))
        assert _rel(y0, F.conv3d(x.float(), wt.float())) < TOL


@pytest.mark.parametrize("cin,cout,pad", [(16, 16, 0), (32, 32, 0), (64, 64, 0), (32, 32, 1), (32, 8, 1), (64, 32, 1)])
def test_data_and_weight_gradients(cin, cout, pad):
    x, wt = _case(1, cin, cout, 6, 14, 23, seed=9)
    gy_shape = F.conv3d(x.float(), wt.float(), padding=pad).shape
    gy = torch.randn(gy_shape, device=DEV).to(torch.bfloat16)
    x1, w1 = x.clone().requires_grad_(True), wt.clone().requires_grad_(True)
    conv_backends.get("tcgen05").conv3d(x1, w1, 1, (pad,) * 3).backward(gy)
    x2, w2 = x.float().requires_grad_(True), wt.float().requires_grad_(True)
    F.conv3d(x2, w2, padding=pad).backward(gy.float())
    assert _rel(x1.grad, x2.grad) < TOL
    assert _rel(w1.grad, w2.grad) < 2 * TOL


def test_large_volume_linearity():
    """Full-size layer (conv_0_0 at cfg1): no golden at this size, so check a size-independent property -- the kernel
    is linear in its input: conv(2x) == 2 conv(x) exactly (powers of two commute with every rounding)."""
    x, wt = _case(1, 32, 8, 192, 128, 160, seed=3)
    be = conv_backends.get("tcgen05")
    y1, y2 = be.conv3d(x, wt, 1, (1, 1, 1)), be.conv3d(x * 2, wt, 1, (1, 1, 1))
    assert torch.equal(y2, y1 * 2)
    sub = F.conv3d(x[:, :, 90:100].float(), wt.float(), padding=1)[:, :, 1:-1]
    assert _rel(y1[:, :, 91:99], sub) < TOL


def test_unsupported_operands_fall_back_to_library_conv():
    x, wt = _case(1, 8, 1, 4, 6, 8)
    n0 = mvs_b200.launch_count()
    y = conv_backends.get("tcgen05").conv3d(x, wt, 1, (1, 1, 1))          # 8 -> 1 channels: not a tensor-core shape
    assert mvs_b200.launch_count() == n0 and y.shape == (1, 1, 4, 6, 8)
    assert conv3d_sm100.available()


@pytest.mark.parametrize("shape", [(1, 6, 9, 14), (2, 5, 33, 47), (1, 1, 1, 1), (1, 3, 8, 160)])
def test_conv_out_forward_and_gradients(shape):
    """K3c: Conv3d(8, 1, 3, padding=1) (model.py:91,123) -- forward, data and weight gradient vs torch fp32."""
    from mvs_b200 import ops
    B, D, h, w = shape
    g = torch.Generator().manual_seed(D * 100 + w)
    z = torch.randn(B, 8, D, h, w, generator=g).to(DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
    wt = (torch.randn(1, 8, 3, 3, 3, generator=g) / 216 ** 0.5).to(DEV)
    gy = torch.randn(B, 1, D, h, w, generator=g).to(DEV)
    z1, w1 = z.clone().requires_grad_(True), wt.clone().requires_grad_(True)
    y = ops.conv_out(z1, w1)
    y.backward(gy)
    z2, w2 = z.float().requires_grad_(True), wt.clone().requires_grad_(True)
    yr = F.conv3d(z2, w2, padding=1)
    yr.backward(gy)
    assert y.dtype == torch.float32 and y.shape == yr.shape
    assert _rel(y, yr) < 1e-5                                   # fp32 accumulation of exact bf16 inputs
    assert _rel(z1.grad, z2.grad) < TOL                         # stored as bf16
    assert _rel(w1.grad, w2.grad) < 1e-4
    y2 = ops.conv_out(z1.detach(), w1.detach())
    assert torch.equal(y, y2)


@pytest.mark.parametrize("form", ["fused", "classes"])
@pytest.mark.parametrize("cin,cout", [(64, 32), (32, 16), (16, 8)])
@pytest.mark.parametrize("dims", [(12, 10, 14), (11, 9, 15), (8, 7, 12), (6, 33, 47), (20, 70, 38)])
def test_transposed_conv_by_parity_classes(cin, cout, dims, form, monkeypatch):
    """Stride-2 ConvTranspose3d (model.py:229-234) from the central box to the canvas on the tcgen05 kernels -- "fused": one
    launch with the 8 output-parity classes side by side in TMEM, "classes": 8 launches of the stride-1 kernel -- vs torch's
    conv_transpose3d (fp32) on the same bf16 operands, forward and both gradients; the two forms agree to one bf16 rounding."""
    from mvs_b200.regulariser import central_region
    monkeypatch.setenv("MVSB200_DECONV", form)
    monkeypatch.setenv("MVSB200_S2_WGRAD", "lines" if form == "fused" else "cudnn")       # own and library weight gradient
    reg = [central_region(n) for n in dims]
    m = [hi - lo + 1 for lo, hi, _ in reg]
    pads = tuple(L for _, _, L in reg)                     # left padding of the equivalent small transposed conv
    g = torch.Generator().manual_seed(cin + cout + dims[0])
    x = torch.randn(2, cin, *m, generator=g).to(DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
    wt = (torch.randn(cin, cout, 3, 3, 3, generator=g) / (8 * cin) ** 0.5).to(DEV).to(torch.bfloat16)
    be = conv_backends.get("tcgen05")
    x1, w1 = x.clone().requires_grad_(True), wt.clone().requires_grad_(True)
    n0 = mvs_b200.launch_count()
    y = be.conv_transpose3d(x1, w1, 2, pads, dims)
    assert mvs_b200.launch_count() >= n0 + (8 if form == "classes" else 1)
    if form == "fused":
        monkeypatch.setenv("MVSB200_DECONV", "classes")
        # same products, another summation order of the taps in the fp32 accumulator: equal up to one bf16 rounding
        assert _rel(y, be.conv_transpose3d(x1.detach(), w1.detach(), 2, pads, dims)) < 4e-3
    x2, w2 = x.float().requires_grad_(True), wt.float().requires_grad_(True)
    ref = conv_backends.TorchConvBackend.conv_transpose3d(x2, w2, 2, pads, dims)
    assert y.shape == ref.shape
    assert _rel(y, ref) < TOL
    gy = torch.randn(ref.shape, generator=g).to(DEV).to(torch.bfloat16)
    y.backward(gy)
    ref.backward(gy.float())
    assert _rel(x1.grad, x2.grad) < TOL
    assert _rel(w1.grad, w2.grad) < 2 * TOL


@pytest.mark.parametrize("wgrad", ["lines", "tcgen05", "cudnn"])
@pytest.mark.parametrize("cin,cout", [(32, 112), (32, 16), (16, 32), (64, 32)])
@pytest.mark.parametrize("dims", [(12, 10, 14), (11, 9, 15), (8, 7, 12), (6, 33, 47), (4, 6, 400)])
def test_stride2_conv_on_the_central_box(cin, cout, dims, wgrad, monkeypatch):
    """The stride-2 branches (model.py:104-110, padding dim/2+1) evaluated on the central box by the tcgen05 stride-2
    kernel (parity sub-lattice slabs) vs the reference formulation conv3d(stride 2, padding dim/2+1) cropped to the box.
    Gradients: "lines" = the shipped own kernels (weight gradient on conv3d_s2_wgrad_lines_kernel, data gradient as a
    transposed convolution over K chunks on deconv3d_s2_kc_kernel; the 112 stacked channels as the three branches 16+32+64),
    "tcgen05" = the parity-class weight gradient, "cudnn" = both gradients through the library."""
    from mvs_b200.regulariser import central_region
    monkeypatch.setenv("MVSB200_S2_WGRAD", wgrad)
    monkeypatch.setenv("MVSB200_S2_DGRAD", "cudnn" if wgrad == "cudnn" else "tcgen05")
    reg = [central_region(n) for n in dims]
    box = tuple(slice(lo, hi + 1) for lo, hi, _ in reg)
    g = torch.Generator().manual_seed(cin + cout + dims[1])
    x = torch.randn(2, cin, *dims, generator=g).to(DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
    wt = (torch.randn(cout, cin, 3, 3, 3, generator=g) / (27 * cin) ** 0.5).to(DEV).to(torch.bfloat16)
    x1, w1 = x.clone().requires_grad_(True), wt.clone().requires_grad_(True)
    splits = [16, 32, 64] if cout == 112 else None
    n0 = mvs_b200.launch_count()
    y = conv_backends.get("tcgen05").conv3d_s2_box(x1, w1, tuple(L for _, _, L in reg), tuple(hi - lo + 1 for lo, hi, _ in reg), splits)
    assert y is not None
    if splits is not None:
        assert [t.shape[1] for t in y] == splits
        y = torch.cat(y, 1)
    x2, w2 = x.float().requires_grad_(True), wt.float().requires_grad_(True)
    full = F.conv3d(x2, w2, None, 2, tuple(n // 2 + 1 for n in dims))        # the reference's layer on the full canvas
    ref = full[(slice(None), slice(None)) + box]
    assert y.shape == ref.shape
    assert _rel(y, ref) < TOL
    outside = full.clone()
    outside[(slice(None), slice(None)) + box] = 0
    assert outside.abs().max() == 0                                            # ... which is zero outside the box
    gy = torch.randn(ref.shape, generator=g).to(DEV).to(torch.bfloat16)
    n1 = mvs_b200.launch_count()
    y.backward(gy)
    if wgrad == "lines" and dims[2] % 2 == 0 and cin <= 32 and not (dims[2] == 400 and cout == 112):   # (201-voxel lines of the
        # 112-channel gradient do not fit two ring stages: that shape runs on the parity-class kernel)
        # backward: one weight-gradient launch + one data-gradient launch (+ one filter-packing launch per K chunk of the
        # transposed convolution) -- no library convolution
        assert mvs_b200.launch_count() - n1 == 2 + (3 if cout == 112 else 1)
    ref.backward(gy.float())
    assert _rel(x1.grad, x2.grad) < TOL
    assert _rel(w1.grad, w2.grad) < 2 * TOL


@pytest.mark.parametrize("co,ci", [(8, 32), (16, 16), (112, 32), (64, 64)])
def test_filter_packing_kernel(co, ci):
    """mvsb200_pack_filter (one launch per filter) against the torch expressions it replaces: natural, flipped + transposed (data
    gradient), depth-innermost (kdn) tap orders, zero rows / columns / slots."""
    from mvs_b200 import conv3d_sm100 as c
    g = torch.Generator().manual_seed(co * 100 + ci)
    w = torch.randn(co, ci, 3, 3, 3, generator=g).to(DEV)
    n_rows = (co + 15) // 16 * 16
    ref = torch.zeros(27, n_rows, ci, dtype=torch.bfloat16, device=DEV)
    ref[:, :co] = w.permute(2, 3, 4, 0, 1).reshape(27, co, ci).to(torch.bfloat16)
    assert torch.equal(c._pack(w, 0, n_rows, c._NAT), ref)
    wd = w.flip(2, 3, 4).transpose(0, 1)                                     # [ci, co, ...]: the data gradient's filter
    n_rows_d = (ci + 15) // 16 * 16
    refd = torch.zeros(27, n_rows_d, max(co, 16), dtype=torch.bfloat16, device=DEV)
    refd[:, :ci, :co] = wd.permute(2, 3, 4, 0, 1).reshape(27, ci, co).to(torch.bfloat16)
    got = c._pack(w, 1, n_rows_d, c._FLIP, n_cols=max(co, 16))
    assert torch.equal(got, refd)
    kdn = refd.view(3, 3, 3, n_rows_d, -1).permute(1, 2, 0, 3, 4).reshape(27, n_rows_d, -1)
    assert torch.equal(c._pack(w, 1, n_rows_d, c._kdn_order(c._FLIP), n_cols=max(co, 16)), kdn)
    # transposed-convolution layout [k][co][ci] from [Cin, Cout, ...] with a trailing zero slot, a column window
    wt = torch.randn(ci, co, 3, 3, 3, generator=g).to(DEV)
    c0, n = (16, 16) if ci >= 32 else (0, ci)
    reft = torch.zeros(28, n_rows, n, dtype=torch.bfloat16, device=DEV)
    reft[:27, :co] = wt.permute(2, 3, 4, 1, 0).reshape(27, co, ci)[:, :, c0:c0 + n].to(torch.bfloat16)
    assert torch.equal(c._pack(wt, 1, n_rows, c._NAT + [-1], c0=c0, cols_real=n), reft)


@pytest.mark.parametrize("cin,cout", [(64, 32), (32, 16), (16, 8)])
def test_transposed_conv_epilogue_statistics_feed_the_batchnorm(cin, cout):
    """mvsb200_deconv3d_s2_fwd_stats + mvsb200_bn_finalize_affine: the per-channel sums the transposed convolution's epilogue takes
    of what it writes give the following train-mode BatchNorm the mean / variance / output (and running statistics) of the
    statistics pass over the canvas, up to the rounding of the stores."""
    from mvs_b200 import ops
    from mvs_b200.regulariser import central_region
    dims = (12, 10, 14)
    reg = [central_region(n) for n in dims]
    m = [hi - lo + 1 for lo, hi, _ in reg]
    pads = tuple(L for _, _, L in reg)
    g = torch.Generator().manual_seed(cin * 7 + cout)
    x = torch.randn(2, cin, *m, generator=g).to(DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
    wt = (torch.randn(cin, cout, 3, 3, 3, generator=g) / (8 * cin) ** 0.5).to(DEV)
    gamma, beta = (torch.rand(cout, generator=g) + 0.5).to(DEV), torch.randn(cout, generator=g).to(DEV)
    U = conv_backends.get("tcgen05").conv_transpose3d_alloc(x, wt, 2, pads, dims)
    part = getattr(U, "_mvs_bn_partials", None)
    assert part is not None and part[2] == dims and part[0].shape[-1] == cout
    rm1, rv1, n1 = torch.zeros(cout, device=DEV), torch.ones(cout, device=DEV), torch.zeros((), dtype=torch.int64, device=DEV)
    rm2, rv2, n2 = rm1.clone(), rv1.clone(), n1.clone()
    n0 = mvs_b200.launch_count()
    y1, mean1, var1 = ops.batchnorm_relu_train(U, gamma, beta, canvas=dims, running=(rm1, rv1, n1), partials=part)
    used = mvs_b200.launch_count() - n0
    y2, mean2, var2 = ops.batchnorm_relu_train(U, gamma, beta, canvas=dims, running=(rm2, rv2, n2))
    assert used == (mvs_b200.launch_count() - n0 - used) - 1                  # one launch fewer: no statistics pass
    # the epilogue sums its fp32 accumulators, the statistics pass the bf16 values that were stored: they differ by the (unbiased,
    # 2^-9 relative) rounding of the stores -- a few 1e-5 of a standard deviation on the mean, ~1e-4 relative on the variance
    Uf = U.float()
    std = float(Uf.std())
    assert float((mean1 - Uf.mean((0, 2, 3, 4))).abs().max()) < 2e-4 * std
    assert torch.allclose(var1, Uf.var((0, 2, 3, 4), unbiased=False), rtol=1e-3, atol=1e-6)
    assert float((mean1 - mean2).abs().max()) < 2e-4 * std and torch.allclose(var1, var2, rtol=1e-3, atol=1e-6)
    assert torch.allclose(y1.float(), y2.float(), rtol=2e-2, atol=2e-2)
    assert float((rm1 - rm2).abs().max()) < 2e-5 * std and torch.allclose(rv1, rv2, rtol=1e-4, atol=1e-6) and int(n1) == int(n2) == 1


def test_strided_convolution_epilogue_sums_feed_the_box_batchnorm(monkeypatch):
    """The stacked stride-2 branches leave per-channel (sum, sum of squares) of their outputs from the kernels' epilogues
    (mvsb200_conv3d_s2_fwd_stats); the box BatchNorm that follows takes them instead of a statistics pass.  Against the sums of
    the stored bf16 outputs: the fp32 accumulators differ from their bf16 roundings by 2^-9 relative, unbiased."""
    from mvs_b200 import conv3d_sm100 as c
    torch.manual_seed(3)
    B, cin, D, h, w = 2, 32, 12, 20, 24
    x = torch.randn(B, cin, D, h, w, device=DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
    wcat = (torch.randn(112, cin, 3, 3, 3, device=DEV) / 30).requires_grad_(True)
    pads, out_dims, splits = (1, 2, 1), (6, 10, 12), (16, 32, 64)
    outs = c.Tcgen05ConvBackend.conv3d_s2_box(x, wcat, pads, out_dims, splits)
    assert len(outs) == 3
    for o, n in zip(outs, splits):
        s1, s2 = o._mvs_box_sums
        assert s1.shape == (n,) and s2.shape == (n,)
        of = o.detach().float()
        r1, r2 = of.sum((0, 2, 3, 4)), (of * of).sum((0, 2, 3, 4))
        scale = of.abs().amax((0, 2, 3, 4)) * (of[:, 0].numel() ** 0.5)
        assert ((s1 - r1).abs() <= 2e-2 * scale + 1e-3).all(), (s1 - r1).abs().max()
        assert torch.allclose(s2, r2, rtol=5e-3)
    monkeypatch.setenv("MVSB200_S2_STATS", "0")
    outs0 = c.Tcgen05ConvBackend.conv3d_s2_box(x, wcat, pads, out_dims, splits)
    assert all(torch.equal(a, b) for a, b in zip(outs, outs0)) and not hasattr(outs0[0], "_mvs_box_sums")
