"""Parity of the CUDA path (called through the C ABI via mvs_b200) against (a) golden vectors produced by the
unmodified reference and (b) the CPU oracle on seeded inputs.  Tolerances are BASELINE.json's:
warped / variance / probability volumes 1e-4 relative (max-norm) in fp32, 1e-2 with bf16 storage;
depth maps within 0.5 % of the depth interval.  Run on the B200 box with `-m gpu`."""
import os

import numpy as np
import pytest
import torch

import mvs_b200
import plane_sweep as ps

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CASES = ["tiny_b1v3", "b2v3", "v5", "v7_odd"]


def _load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name + ".npz")))


def _t(a, dev="cpu"):
    return torch.from_numpy(np.asarray(a)).to(dev)


def _relmax(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / np.abs(b).max())


def _sweep_inputs(g):
    return (_t(g["K"]), _t(g["R"]), _t(g["T"]), _t(g["d_min"]), _t(g["d_int"]))


def _cost_from_golden_inputs(g, feat, out_dtype=torch.float32):
    B, V, D = int(g["B"]), int(g["V"]), int(g["D"])
    warped, d_batch, ref_idx = mvs_b200.homography_warping(*_sweep_inputs(g), feat, B, V, D, int(g["d_scale"]))
    return mvs_b200.assemble_cost_volume(warped, V, out_dtype), warped, d_batch, ref_idx


# ---------------------------------------------------------------------------------------------- K1
def test_launches_go_through_our_library(golden_dir):
    g = _load(golden_dir, "tiny_b1v3")
    n0 = mvs_b200.launch_count()
    _cost_from_golden_inputs(g, _t(g["feat"], DEV))
    assert mvs_b200.launch_count() > n0
    assert os.path.basename(mvs_b200.LIB_PATH) in open("/proc/self/maps").read()


def test_warped_volumes_match_reference(golden_dir):
    g = _load(golden_dir, "tiny_b1v3")
    _, warped, d_batch, ref_idx = _cost_from_golden_inputs(g, _t(g["feat"], DEV))
    assert tuple(warped.shape) == g["warped"].shape
    w = warped.materialize()
    assert _relmax(w.cpu().numpy(), g["warped"]) < 1e-4
    assert np.array_equal(d_batch.cpu().numpy(), g["d_batch"]) and d_batch.is_cuda
    assert np.array_equal(ref_idx.numpy(), g["ref_idx"]) and not ref_idx.is_cuda
    # the handle behaves like the tensor when a caller treats it as one (reference: costvolume.py on a tensor)
    assert torch.allclose(torch.sum(warped, dim=(1, 2, 3, 4)), w.sum((1, 2, 3, 4)))


@pytest.mark.parametrize("name", CASES)
def test_cost_volume_fp32_matches_reference(golden_dir, name):
    g = _load(golden_dir, name)
    cost, *_ = _cost_from_golden_inputs(g, _t(g["feat"], DEV))
    assert tuple(cost.shape) == g["cost"].shape and cost.dtype == torch.float32
    assert cost.is_contiguous(memory_format=torch.channels_last_3d)
    assert _relmax(cost.cpu().numpy(), g["cost"]) < 1e-4


@pytest.mark.parametrize("name", CASES)
def test_cost_volume_bf16_storage(golden_dir, name):
    g = _load(golden_dir, name)
    cost, *_ = _cost_from_golden_inputs(g, _t(g["feat"], DEV), torch.bfloat16)
    assert cost.dtype == torch.bfloat16
    assert _relmax(cost.float().cpu().numpy(), g["cost"]) < 1e-2


def test_channels_last_features_are_zero_copy_and_equal(golden_dir):
    g = _load(golden_dir, "b2v3")
    f = _t(g["feat"], DEV)
    a, *_ = _cost_from_golden_inputs(g, f)
    b, *_ = _cost_from_golden_inputs(g, f.contiguous(memory_format=torch.channels_last))
    assert torch.equal(a, b)


def test_zero_depth_plane_is_nan_like_the_reference(golden_dir):
    g = _load(golden_dir, "val_dmin0")                       # validate.py:40 sweeps from d_min = 0
    cost, *_ = _cost_from_golden_inputs(g, _t(g["feat"], DEV))
    c = cost.cpu().numpy()
    assert np.isnan(c[:, :, 0]).all() and np.isnan(g["cost"][:, :, 0]).all()
    assert _relmax(c[:, :, 1:], g["cost"][:, :, 1:]) < 1e-4


def test_plain_tensor_cost_volume(golden_dir):
    g = _load(golden_dir, "tiny_b1v3")
    w = _t(g["warped"], DEV).requires_grad_(True)
    cost = mvs_b200.assemble_cost_volume(w, 3)
    assert _relmax(cost.detach().cpu().numpy(), g["cost"]) < 1e-5
    gw = _t(g["gcost"], DEV)
    (gx,) = torch.autograd.grad((cost * gw).sum(), w)
    wr = _t(g["warped"]).requires_grad_(True)
    (gr,) = torch.autograd.grad((ps.variance_cost(wr, 3) * _t(g["gcost"])).sum(), wr)
    assert _relmax(gx.cpu().numpy(), gr.numpy()) < 1e-5


# ---------------------------------------------------------------------------------------------- K2
@pytest.mark.parametrize("name", CASES)
def test_feature_gradient_matches_reference_autograd(golden_dir, name):
    g = _load(golden_dir, name)
    feat = _t(g["feat"], DEV).requires_grad_(True)
    cost, *_ = _cost_from_golden_inputs(g, feat)
    (gf,) = torch.autograd.grad((cost * _t(g["gcost"], DEV)).sum(), feat)
    assert gf.shape == feat.shape
    assert _relmax(gf.cpu().numpy(), g["gfeat"]) < 1e-4


def test_feature_gradient_bf16_upstream_and_nhwc_features(golden_dir):
    g = _load(golden_dir, "tiny_b1v3")
    feat = _t(g["feat"], DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    cost, *_ = _cost_from_golden_inputs(g, feat, torch.bfloat16)
    gc = _t(g["gcost"], DEV).to(torch.bfloat16)
    (gf,) = torch.autograd.grad((cost * gc).sum(), feat)
    # oracle with the same bf16-rounded upstream gradient
    fr = _t(g["feat"]).requires_grad_(True)
    prm = ps.view_params_closed64(g["K"], g["R"], g["T"], 1, 3, 16, 20)
    ix, iy = ps.sample_positions_closed64(prm, g["d_batch"].reshape(1, -1)[[0, 0, 0]], 16, 20)
    c = ps.variance_cost(ps.bilinear_grid_sample(fr, _t(ix), _t(iy)), 3)
    (gr,) = torch.autograd.grad((c * gc.float().cpu()).sum(), fr)
    assert _relmax(gf.cpu().numpy(), gr.numpy()) < 1e-4


# ---------------------------------------------------------------------------------------------- K4
@pytest.mark.parametrize("name", ["tiny_b1v3", "b2v3", "v7_odd"])
def test_softmax_and_depth_match_reference(golden_dir, name):
    g = _load(golden_dir, name)
    logits = _t(g["logits"], DEV)
    prob = mvs_b200.softmax_over_depth(logits)
    assert _relmax(prob.cpu().numpy(), g["prob"]) < 1e-4
    depth = mvs_b200.extract_depth_map(prob, _t(g["d_batch"], DEV))
    ok = ~ps.tie_pixels(g["prob"])
    step = float(g["d_scale"]) * float(g["d_int"].ravel()[0])
    assert np.abs(depth.cpu().numpy()[:, 0] - g["depth"][:, 0])[ok].max() < 0.005 * step
    # standalone extraction on a plain probability tensor (no stashed ranks), and the fully fused launch
    d2 = mvs_b200.extract_depth_map(_t(g["prob"], DEV), _t(g["d_batch"], DEV))
    assert np.abs(d2.cpu().numpy()[:, 0] - g["depth"][:, 0])[ok].max() < 0.005 * step
    p3, d3 = mvs_b200.softmax_depth(logits, _t(g["d_batch"], DEV))
    assert torch.equal(p3, prob) and torch.equal(d3, depth)


def test_depth_with_planted_ties(golden_dir):
    g = _load(golden_dir, "depth_ties")
    d = mvs_b200.extract_depth_map(_t(g["prob"], DEV), _t(g["d_batch"], DEV)).cpu().numpy()
    oracle, _ = ps.extract_depth(g["prob"], g["d_batch"])
    assert np.abs(d - oracle).max() < 1e-2                    # identical stable tie rule as the oracle, everywhere
    ties = ps.tie_pixels(g["prob"])
    assert np.abs(d[:, 0] - g["depth"][:, 0])[~ties].max() < 0.005 * 40


def test_depth_gradient_matches_reference(golden_dir):
    g = _load(golden_dir, "tiny_b1v3")
    logits = _t(g["logits"], DEV).requires_grad_(True)
    prob = mvs_b200.softmax_over_depth(logits)
    depth = mvs_b200.extract_depth_map(prob, _t(g["d_batch"], DEV))
    (gl,) = torch.autograd.grad((depth * _t(g["gdepth"], DEV)).sum(), logits)
    lr = _t(g["logits"]).requires_grad_(True)
    dr = ps.extract_depth_torch(torch.softmax(lr, 2), _t(g["d_batch"]))
    (gr,) = torch.autograd.grad((dr * _t(g["gdepth"])).sum(), lr)
    assert _relmax(gl.cpu().numpy(), gr.numpy()) < 1e-4


def test_few_planes_keep_everything():
    torch.manual_seed(0)
    logits = torch.randn(1, 1, 3, 5, 7, device=DEV)            # D < N_DEPTH_EST: idx < 5 is always true
    d_batch = ps.depth_table(torch.full((1, 1, 1, 1), 425.0), torch.ones(1, 1, 1, 1), 3, 100).to(DEV)
    prob, depth = mvs_b200.softmax_depth(logits, d_batch)
    ref = (prob[:, 0] * d_batch.view(1, 3, 1, 1)).sum(1, keepdim=True)
    assert torch.allclose(depth, ref, rtol=1e-5)


# ------------------------------------------------------------------------------------ regulariser
def _reg(golden_dir, **kw):
    reg = mvs_b200.CostVolumeReg(**kw)
    w0 = np.load(os.path.join(golden_dir, "reg_weights.npz"))
    reg.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in w0.items()})
    return reg.train()


def test_default_constructor_is_the_native_path(golden_dir):
    """`CostVolumeReg()` exactly as scripts/model.py:161 calls it: parameters on the CUDA device (reference default
    device=DEVICE, model.py:70), convolutions on this library's tcgen05 kernels -- not the library fallback -- and the
    probability volume inside the tolerance BASELINE.json states for that mode (1e-2, bf16 operands / fp32 accumulate),
    with torch's TF32 switches left alone."""
    from mvs_b200 import ops
    g = _load(golden_dir, "tiny_b1v3")
    torch.manual_seed(1234)
    reg = mvs_b200.CostVolumeReg()
    assert reg.precision == "bf16" and all(p.is_cuda for p in reg.parameters())
    w0 = np.load(os.path.join(golden_dir, "reg_weights.npz"))
    for k, v in reg.state_dict().items():                      # seed 1234 on the CUDA generator: same shapes and keys
        assert tuple(v.shape) == w0[k].shape, k
    reg.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in w0.items()})
    reg.train()
    ops.EVENTS = {}
    try:
        prob = reg(_t(g["cost"], DEV))                          # fp32 volume in, as assemble_cost_volume returns it
        torch.cuda.synchronize()
        ran = set(ops.EVENTS)
    finally:
        ops.EVENTS = None
    assert {"conv3d_s1_tc", "conv3d_s2_tc", "deconv3d_s2_tc", "conv_out_fwd", "bn_stats"} <= ran, ran
    assert prob.dtype == torch.float32 and tuple(prob.shape) == g["prob"].shape
    assert _relmax(prob.detach().cpu().numpy(), g["prob"]) < 1e-2


@pytest.mark.parametrize("allow_tf32", [True, False])
@pytest.mark.parametrize("name", ["tiny_b1v3", "b2v3", "v7_odd"])
def test_regulariser_fp32_matches_reference(golden_dir, name, allow_tf32, monkeypatch):
    """precision="fp32" owns its precision: 1e-4 whatever torch.backends.cudnn.allow_tf32 says."""
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", allow_tf32)
    g = _load(golden_dir, name)
    reg = _reg(golden_dir, device=DEV, precision="fp32")
    prob = reg(_t(g["cost"], DEV))
    assert _relmax(prob.detach().cpu().numpy(), g["prob"]) < 1e-4
    sd = reg.state_dict()
    for k, v in g.items():
        if k.startswith("bn_after/"):
            assert np.allclose(sd[k[len("bn_after/"):]].cpu().numpy(), v, rtol=1e-4, atol=1e-6), k


@pytest.mark.parametrize("name", ["tiny_b1v3", "b2v3", "v7_odd", "bquirk_b2v3"])
def test_regulariser_bf16(golden_dir, name):
    g = _load(golden_dir, name)
    reg = _reg(golden_dir, device=DEV)                          # default precision: bf16 operands on tcgen05
    prob = reg(_t(g["cost"], DEV).to(torch.bfloat16))
    assert _relmax(prob.detach().cpu().numpy(), g["prob"]) < 1e-2
    prob32 = _reg(golden_dir, device=DEV)(_t(g["cost"], DEV))  # fp32 volume in: converted inside
    assert _relmax(prob32.detach().cpu().numpy(), g["prob"]) < 1e-2


def test_cpu_parameters_are_refused(golden_dir):
    g = _load(golden_dir, "tiny_b1v3")
    reg = mvs_b200.CostVolumeReg(device="cpu")
    with pytest.raises(mvs_b200.MvsB200Error, match="CPU"):
        reg(_t(g["cost"], DEV))


def test_hot_path_end_to_end_like_mvsnet_forward(golden_dir):
    """The call sequence of MVSNet.forward (model.py:177-187) on the drop-in functions, incl. backward (fp32 mode, 1e-4 class)."""
    g = _load(golden_dir, "tiny_b1v3")
    reg = _reg(golden_dir, device=DEV, precision="fp32")
    feat = _t(g["feat"], DEV).requires_grad_(True)
    warped, d_batch, ref_views = mvs_b200.homography_warping(*_sweep_inputs(g), feat, 1, 3, 8, 60)
    cost = mvs_b200.assemble_cost_volume(warped, 3)
    prob = reg(cost)
    depth = mvs_b200.extract_depth_map(prob, d_batch)
    (depth * _t(g["gdepth"], DEV)).sum().backward()
    ok = ~ps.tie_pixels(g["prob"])
    assert np.abs(depth.detach().cpu().numpy()[:, 0] - g["depth"][:, 0])[ok].max() < 0.005 * 60
    assert feat.grad is not None and torch.isfinite(feat.grad).all()
    gw = dict(reg.named_parameters())["conv_out.weight"].grad.cpu().numpy()
    ref = g["gparam/conv_out.weight"]
    assert np.abs(gw - ref).max() < 1e-3 * np.abs(ref).max()


def test_batch_quirk_with_unequal_d_min(golden_dir):
    """scripts/homography.py:26 tiles the depth table V times along dim 0 while the matrices are ordered b*V+v: flat view i reads
    depth row i mod B.  Golden from the unmodified reference with d_min = (425, 520): the CUDA path reproduces the quirk."""
    g = _load(golden_dir, "bquirk_b2v3")
    assert float(g["d_min"].ravel()[0]) != float(g["d_min"].ravel()[1])
    feat = _t(g["feat"], DEV).requires_grad_(True)
    cost, warped, d_batch, _ = _cost_from_golden_inputs(g, feat)
    assert np.array_equal(d_batch.cpu().numpy(), g["d_batch"])
    assert _relmax(warped.materialize().cpu().numpy(), g["warped"]) < 1e-4
    assert _relmax(cost.detach().cpu().numpy(), g["cost"]) < 1e-4
    (gf,) = torch.autograd.grad((cost * _t(g["gcost"], DEV)).sum(), feat)
    assert _relmax(gf.cpu().numpy(), g["gfeat"]) < 1e-4
    # and the geometrically intended table gives something else (the quirk is really exercised)
    w2, _, _ = mvs_b200.api.homography_warping(*_sweep_inputs(g), feat.detach(), 2, 3, int(g["D"]), int(g["d_scale"]), bug_compatible=False)
    c2 = mvs_b200.assemble_cost_volume(w2, 3)
    assert _relmax(c2.cpu().numpy(), g["cost"]) > 1e-2


# ------------------------------------------------------------------- oracle on seeded inputs, errors
@pytest.mark.parametrize("B,V,D,h,w", [(1, 3, 48, 64, 80), (2, 4, 20, 33, 47), (1, 6, 16, 24, 40), (1, 2, 9, 8, 8),
                                       (1, 8, 5, 16, 24)])
def test_cost_volume_against_live_oracle(B, V, D, h, w):
    gen = torch.Generator().manual_seed(B * 1000 + V * 100 + D)
    K, R, T = ps.synthetic_cameras(B, V, h, w, seed=V)
    d_min, d_int = torch.full((B, 1, 1, 1), 425.0), torch.ones(B, 1, 1, 1)
    feat = torch.randn(B * V, 32, h, w, generator=gen)
    d_scale = 480.0 / D
    ref, _, _ = ps.plane_sweep_cost(feat, K, R, T, d_min, d_int, B, V, D, d_scale, sampler="torch")
    warped, _, _ = mvs_b200.homography_warping(K, R, T, d_min, d_int, feat.to(DEV), B, V, D, d_scale)
    cost = mvs_b200.assemble_cost_volume(warped, V)
    assert _relmax(cost.cpu().numpy(), ref.numpy()) < 1e-4


@pytest.mark.parametrize("V,D,h,w,d0,shift", [(3, 12, 13, 21, 150.0, 0.0), (4, 10, 9, 7, 120.0, 30.0), (2, 16, 2, 2, 425.0, 0.0),
                                              (5, 6, 31, 17, 80.0, -25.0)])
def test_cost_volume_with_footprints_hanging_over_every_edge(V, D, h, w, d0, shift):
    """Forward and backward kernels (one clamped base offset per footprint, padding and clamping folded into the weights) on
    sweeps whose sampling positions leave the image on every side: near planes (large disparity) and a shifted principal point;
    tiny and odd-sized maps.  Forward and the backward of the same sweep against the oracle (its autograd for the backward)."""
    gen = torch.Generator().manual_seed(V * 100 + D + h)
    K, R, T = ps.synthetic_cameras(1, V, h, w, seed=V)
    K = K.clone()
    K[1:, 0, 2] += shift * w / 160.0                                   # source principal points off-centre
    K[1:, 1, 2] -= shift * h / 128.0
    d_min, d_int = torch.full((1, 1, 1, 1), d0), torch.ones(1, 1, 1, 1)
    feat = torch.randn(V, 32, h, w, generator=gen)
    d_scale = 40.0
    fo = feat.clone().requires_grad_(True)
    ref, _, _ = ps.plane_sweep_cost(fo, K, R, T, d_min, d_int, 1, V, D, d_scale, sampler="torch")
    g = torch.randn(ref.shape, generator=gen)
    (ref * g).sum().backward()
    fg = feat.to(DEV).requires_grad_(True)
    warped, _, _ = mvs_b200.homography_warping(K, R, T, d_min, d_int, fg, 1, V, D, d_scale)
    cost = mvs_b200.assemble_cost_volume(warped, V)
    (cost * g.to(DEV)).sum().backward()
    refn = ref.detach().numpy()
    vols = warped.materialize()                                         # [V,32,D,h,w]: zero rows = samples outside the image
    outside = float((vols[1:].abs().amax(1) == 0).float().mean())
    assert outside > 0.02 or h * w <= 4, outside                        # the case really has out-of-image samples
    assert _relmax(cost.detach().cpu().numpy(), refn) < 1e-4
    assert _relmax(fg.grad.cpu().numpy(), fo.grad.numpy()) < 1e-4


def test_bad_arguments_raise():
    K, R, T = ps.synthetic_cameras(1, 3, 8, 8)
    d_min, d_int = torch.full((1, 1, 1, 1), 425.0), torch.ones(1, 1, 1, 1)
    with pytest.raises(ValueError):                                     # rank check, as kornia does
        mvs_b200.homography_warping(K, R, T, d_min, d_int, torch.zeros(3, 32, 8, device=DEV), 1, 3, 4, 10)
    with pytest.raises(ValueError):
        mvs_b200.homography_warping(K, R, T, d_min, d_int, torch.zeros(4, 32, 8, 8, device=DEV), 1, 3, 4, 10)
    wv, _, _ = mvs_b200.homography_warping(K, R, T, d_min, d_int, torch.zeros(3, 16, 8, 8, device=DEV), 1, 3, 4, 10)
    with pytest.raises(mvs_b200.MvsB200Error, match="C must be 32"):
        mvs_b200.assemble_cost_volume(wv, 3)


@pytest.mark.parametrize("B,h,w", [(4, 128, 160), (1, 7, 9), (3, 33, 20)])
def test_fused_masked_l1_loss_matches_the_reference_formula(B, h, w):
    """mvsb200_masked_l1_fwd / _bwd against scripts/loss.py:4-41 restated in torch (mask = gt != 0, per-sample
    normalisation, loss = sum_b l0 + l1, the two accuracies = batch means) -- values and gradients, with invalid (zero) pixels and
    exact ties (gt == estimate, where torch.abs' gradient is 0)."""
    from mvs_b200 import ops
    g = torch.Generator().manual_seed(B * 1000 + h)
    gt = (425 + 480 * torch.rand(B, 1, h, w, generator=g)).to(DEV)
    gt[torch.rand(B, 1, h, w, generator=g).to(DEV) < 0.3] = 0.0                 # invalid pixels
    init = (gt + 5 * torch.randn(B, 1, h, w, generator=g).to(DEV))
    ref = (gt + 3 * torch.randn(B, 1, h, w, generator=g).to(DEV))
    init[:, :, 0, :2] = gt[:, :, 0, :2]                                        # exact ties
    i1, r1 = init.clone().requires_grad_(True), ref.clone().requires_grad_(True)
    n0 = mvs_b200.launch_count()
    loss, a0, a1 = ops.masked_l1_loss(gt, i1, r1)
    (loss + 0.5 * a0 - 0.25 * a1).backward()
    assert mvs_b200.launch_count() - n0 == 2                                   # one launch forward, one backward
    i2, r2 = init.clone().requires_grad_(True), ref.clone().requires_grad_(True)
    mask = (gt != 0).float()
    nv = mask.sum((1, 2, 3))
    l0 = (mask * (gt - i2).abs()).sum((1, 2, 3)) / nv
    l1 = (mask * (gt - r2).abs()).sum((1, 2, 3)) / nv
    lr, b0, b1 = (l0 + l1).sum(), l0.mean(), l1.mean()
    (lr + 0.5 * b0 - 0.25 * b1).backward()
    for got, want in ((loss, lr), (a0, b0), (a1, b1)):
        assert abs(float(got) - float(want)) <= 1e-5 * abs(float(want))
    assert torch.allclose(i1.grad, i2.grad, rtol=1e-5, atol=1e-9)
    assert torch.allclose(r1.grad, r2.grad, rtol=1e-5, atol=1e-9)
