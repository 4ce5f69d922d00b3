"""Full-size (BASELINE.json config 1/2 shapes) checks through size-independent properties; the CPU oracle is
only used on a strided sample of planes so the test stays in seconds."""
import os

import numpy as np
import pytest
import torch

import mvs_b200
import plane_sweep as ps

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
B, V, C, D, H, W = 1, 3, 32, 192, 128, 160


def _inputs(seed=0, b=B, v=V):
    gen = torch.Generator().manual_seed(seed)
    K, R, T = ps.synthetic_cameras(b, v, H, W, seed=seed)
    feat = torch.randn(b * v, C, H, W, generator=gen)
    return K, R, T, torch.full((b, 1, 1, 1), 425.0), torch.ones(b, 1, 1, 1), feat


def _cost(K, R, T, d_min, d_int, feat, b=B, v=V, dtype=torch.float32):
    warped, d_batch, _ = mvs_b200.homography_warping(K, R, T, d_min, d_int, feat.to(DEV), b, v, D, 480.0 / D)
    return mvs_b200.assemble_cost_volume(warped, v, dtype), warped, d_batch


def test_fullsize_against_oracle_on_sampled_planes():
    K, R, T, d_min, d_int, feat = _inputs()
    cost, _, _ = _cost(K, R, T, d_min, d_int, feat)
    prm = ps.view_params_closed64(K, R, T, B, V, H, W)
    d0 = ps.depth_table(d_min, d_int, D, 480.0 / D).reshape(B, D).numpy()
    planes = [0, 1, 47, 48, 95, 96, 143, 144, 190, 191]      # incl. both sides of the depth-run boundaries
    ix, iy = ps.sample_positions_closed64(prm, d0[ps.view_depth_rows(B, V)][:, planes], H, W)
    ref = ps.variance_cost(ps.bilinear_grid_sample(feat, torch.from_numpy(ix), torch.from_numpy(iy)), V).numpy()
    got = cost[:, :, planes].cpu().numpy()
    assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-4


def test_fused_kernel_equals_materialise_then_variance():
    K, R, T, d_min, d_int, feat = _inputs(1)
    cost, warped, _ = _cost(K, R, T, d_min, d_int, feat)
    two_step = mvs_b200.assemble_cost_volume(warped.materialize(), V)
    assert (cost - two_step).abs().max().item() <= 2e-6 * two_step.abs().max().item()


def test_depth_run_length_does_not_change_results():
    K, R, T, d_min, d_int, feat = _inputs(2)
    a, _, _ = _cost(K, R, T, d_min, d_int, feat)
    try:
        os.environ["MVSB200_DCHUNK"] = "17"
        b, _, _ = _cost(K, R, T, d_min, d_int, feat)
    finally:
        os.environ.pop("MVSB200_DCHUNK", None)
    assert torch.equal(a, b)


def test_identical_views_give_zero_variance_and_scaling_is_quadratic():
    K, R, T, d_min, d_int, feat = _inputs(3)
    Ks, Rs, Ts = K[:1].repeat(V, 1, 1), R[:1].repeat(V, 1, 1), T[:1].repeat(V, 1, 1)
    same = feat[:1].repeat(V, 1, 1, 1)
    c0, _, _ = _cost(Ks, Rs, Ts, d_min, d_int, same)
    # not exactly 0: (x+x+x)/3 need not round back to x (the reference's mean has the same rounding)
    assert c0.abs().max().item() <= 1e-12 * same.abs().max().item() ** 2
    c1, _, _ = _cost(K, R, T, d_min, d_int, feat)
    c4, _, _ = _cost(K, R, T, d_min, d_int, feat * 4.0)      # power of two: exact in fp32
    assert torch.equal(c4, c1 * 16.0)


def test_source_view_order_is_irrelevant():
    K, R, T, d_min, d_int, feat = _inputs(4)
    a, _, _ = _cost(K, R, T, d_min, d_int, feat)
    p = [0, 2, 1]
    b, _, _ = _cost(K[p], R[p], T[p], d_min, d_int, feat[p])
    assert (a - b).abs().max().item() <= 1e-5 * a.abs().max().item()


def test_backward_fullsize_linearity_and_finite_difference():
    K, R, T, d_min, d_int, feat = _inputs(5)
    f = feat.to(DEV).requires_grad_(True)
    warped, _, _ = mvs_b200.homography_warping(K, R, T, d_min, d_int, f, B, V, D, 480.0 / D)
    cost = mvs_b200.assemble_cost_volume(warped, V)
    g = torch.randn(cost.shape, device=DEV, generator=torch.Generator(DEV).manual_seed(0))
    (gf,) = torch.autograd.grad((cost * g).sum(), f, retain_graph=True)
    (gf2,) = torch.autograd.grad((cost * (2 * g)).sum(), f)
    assert torch.allclose(gf2, 2 * gf, rtol=1e-4, atol=1e-4 * gf.abs().max().item())      # atomics reorder sums
    # cost is quadratic in the features: <grad, df> == cost(f+df/2) - cost(f-df/2) contracted with g, exactly
    df = torch.randn_like(f)
    with torch.no_grad():
        cp, _, _ = _cost(K, R, T, d_min, d_int, (f + 0.5 * df).cpu())
        cm, _, _ = _cost(K, R, T, d_min, d_int, (f - 0.5 * df).cpu())
        lhs = (gf.double() * df.double()).sum().item()
        rhs = ((cp.double() - cm.double()) * g.double()).sum().item()
    assert abs(lhs - rhs) <= 1e-3 * max(abs(lhs), abs(rhs))


def test_softmax_depth_fullsize_properties():
    torch.manual_seed(0)
    logits = torch.randn(2, 1, D, H, W, device=DEV)
    d_batch = ps.depth_table(torch.full((2, 1, 1, 1), 425.0), torch.ones(2, 1, 1, 1), D, 2.5).to(DEV)
    prob, depth = mvs_b200.softmax_depth(logits, d_batch)
    assert torch.allclose(prob.sum(2), torch.ones_like(prob.sum(2)), atol=1e-5)
    assert torch.allclose(prob, torch.softmax(logits, 2), rtol=1e-5, atol=1e-8)
    assert (depth >= 425.0).all() and (depth <= 425.0 + 2.5 * (D - 1)).all()
    sub = slice(0, 16)
    oracle, _ = ps.extract_depth(prob[:, :, :, sub].cpu().numpy(), d_batch.cpu().numpy())
    ties = ps.tie_pixels(prob[:, :, :, sub].cpu().numpy())
    err = np.abs(depth[:, :, sub].cpu().numpy() - oracle)[:, 0]
    assert err[~ties].max() < 0.005 * 2.5
    # shifting the logits per pixel changes nothing
    p2, d2 = mvs_b200.softmax_depth(logits + 3.0, d_batch)
    assert torch.allclose(p2, prob, rtol=1e-4, atol=1e-9)


def test_batch4_bf16_config2_shape_runs_and_matches_fp32():
    K, R, T, d_min, d_int, feat = _inputs(6, b=4)
    a, _, _ = _cost(K, R, T, d_min, d_int, feat, b=4)
    b, _, _ = _cost(K, R, T, d_min, d_int, feat, b=4, dtype=torch.bfloat16)
    assert a.shape == (4, C, D, H, W)
    assert (a - b.float()).abs().max().item() <= 1e-2 * a.abs().max().item()
