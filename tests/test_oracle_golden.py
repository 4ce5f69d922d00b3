"""Pin the CPU oracle (oracle/plane_sweep.py) against vectors produced by the unmodified reference
(oracle/make_golden.py -> tests/golden/*.npz).  CPU only."""
import os

import numpy as np
import pytest
import torch

import plane_sweep as ps

CASES = ["tiny_b1v3", "b2v3", "v5", "v7_odd"]


def _load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name + ".npz")))


def _t(a):
    return torch.from_numpy(np.asarray(a))


def _relmax(a, b):
    return float(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)).max() / np.abs(b).max())


@pytest.mark.parametrize("name", CASES)
def test_depth_table_and_chain32_homographies(golden_dir, name):
    g = _load(golden_dir, name)
    B, V, D = int(g["B"]), int(g["V"]), int(g["D"])
    d0 = ps.depth_table(_t(g["d_min"]), _t(g["d_int"]), D, int(g["d_scale"]))
    assert np.array_equal(d0.numpy(), g["d_batch"])
    H = ps.homographies_chain32(_t(g["K"]), _t(g["R"]), _t(g["T"]), d0, B, V)
    assert _relmax(H.numpy(), g["H"]) < 1e-6
    assert np.array_equal(g["ref_idx"], np.arange(0, B * V, V))


@pytest.mark.parametrize("name", CASES)
def test_closed_form_positions_match_reference_homographies(golden_dir, name):
    """fp64 closed form (what the CUDA kernel evaluates) vs positions implied by the reference's fp32 H."""
    g = _load(golden_dir, name)
    B, V, D, h, w = (int(g[k]) for k in "BVDhw")
    prm = ps.view_params_closed64(g["K"], g["R"], g["T"], B, V, h, w)
    d0 = g["d_batch"].reshape(B, D)
    ix, iy = ps.sample_positions_closed64(prm, d0[ps.view_depth_rows(B, V)], h, w)
    jx, jy = ps.sample_positions_from_H(g["H"], h, w)
    assert np.abs(ix - jx).max() < 2e-3 and np.abs(iy - jy).max() < 2e-3   # fp32 chain noise, px
    for b in range(B):                       # reference view: identity => fixed half-pixel resample
        assert np.allclose(ix[b * V, :, 0, 0], -0.5, atol=1e-9)
        assert np.allclose(ix[b * V, :, -1, -1], (w - 1) * w / (w - 1.0) - 0.5, atol=1e-9)


def test_warped_volume_tiny(golden_dir):
    g = _load(golden_dir, "tiny_b1v3")
    cost, warped, _ = ps.plane_sweep_cost(_t(g["feat"]), g["K"], g["R"], g["T"], _t(g["d_min"]), _t(g["d_int"]),
                                          1, 3, int(g["D"]), int(g["d_scale"]))
    assert _relmax(warped, g["warped"]) < 1e-4
    assert _relmax(cost, g["cost"]) < 1e-4
    # the torch sampler agrees with the numpy gather
    _, warped_t, _ = ps.plane_sweep_cost(_t(g["feat"]), g["K"], g["R"], g["T"], _t(g["d_min"]), _t(g["d_int"]),
                                         1, 3, int(g["D"]), int(g["d_scale"]), sampler="torch")
    assert _relmax(warped_t.numpy(), g["warped"]) < 1e-4


@pytest.mark.parametrize("name", CASES)
def test_cost_volume(golden_dir, name):
    g = _load(golden_dir, name)
    B, V = int(g["B"]), int(g["V"])
    cost, _, d0 = ps.plane_sweep_cost(_t(g["feat"]), g["K"], g["R"], g["T"], _t(g["d_min"]), _t(g["d_int"]),
                                      B, V, int(g["D"]), int(g["d_scale"]))
    assert cost.shape == g["cost"].shape
    assert _relmax(cost, g["cost"]) < 1e-4


def test_reference_chain_restatement_is_bit_faithful(golden_dir):
    g = _load(golden_dir, "tiny_b1v3")
    warped, _ = ps.warp_reference_chain(_t(g["feat"]), _t(g["K"]), _t(g["R"]), _t(g["T"]), _t(g["d_min"]),
                                        _t(g["d_int"]), 1, 3, int(g["D"]), int(g["d_scale"]))
    assert _relmax(warped.numpy(), g["warped"]) < 2e-6


def test_d_equal_zero_plane_is_nan_in_reference(golden_dir):
    """validate.py:40 sweeps from d_min = 0: the reference divides by d = 0 => plane 0 is NaN everywhere."""
    g = _load(golden_dir, "val_dmin0")
    assert np.isnan(g["cost"][:, :, 0]).all() and not np.isnan(g["cost"][:, :, 1:]).any()
    cost, _, _ = ps.plane_sweep_cost(_t(g["feat"]), g["K"], g["R"], g["T"], _t(g["d_min"]), _t(g["d_int"]),
                                     1, 3, int(g["D"]), int(g["d_scale"]))
    assert np.isnan(cost[:, :, 0]).all()
    assert _relmax(cost[:, :, 1:], g["cost"][:, :, 1:]) < 1e-4


def test_feature_gradient_of_cost(golden_dir):
    g = _load(golden_dir, "tiny_b1v3")
    feat = _t(g["feat"]).clone().requires_grad_(True)
    B, V, D, h, w = (int(g[k]) for k in "BVDhw")
    prm = ps.view_params_closed64(g["K"], g["R"], g["T"], B, V, h, w)
    ix, iy = ps.sample_positions_closed64(prm, g["d_batch"].reshape(B, D)[ps.view_depth_rows(B, V)], h, w)
    warped = ps.bilinear_grid_sample(feat, _t(ix), _t(iy))
    cost = ps.variance_cost(warped, V)
    (gf,) = torch.autograd.grad((cost * _t(g["gcost"])).sum(), feat)
    assert _relmax(gf.numpy(), g["gfeat"]) < 1e-4


@pytest.mark.parametrize("name", ["tiny_b1v3", "b2v3", "v7_odd"])
def test_regulariser_and_depth(golden_dir, name):
    g = _load(golden_dir, name)
    w0 = {k: _t(v).clone() for k, v in np.load(os.path.join(golden_dir, "reg_weights.npz")).items()}
    prob, logits = ps.reg_forward(w0, _t(g["cost"]), train_bn=True, update_running=True, return_logits=True)
    assert _relmax(logits.numpy(), g["logits"]) < 1e-4
    assert _relmax(prob.numpy(), g["prob"]) < 1e-4
    for k, v in g.items():
        if k.startswith("bn_after/"):
            assert np.allclose(w0[k[len("bn_after/"):]].numpy(), v, rtol=1e-5, atol=1e-6), k
    depth, ranks = ps.extract_depth(g["prob"], g["d_batch"])
    ok = ~ps.tie_pixels(g["prob"])
    step = float(g["d_scale"]) * float(g["d_int"].ravel()[0])
    assert np.abs(depth[:, 0] - g["depth"][:, 0])[ok].max() < 0.005 * step
    assert ok.mean() > 0.99


def test_depth_ties(golden_dir):
    g = _load(golden_dir, "depth_ties")
    depth, _ = ps.extract_depth(g["prob"], g["d_batch"])
    ties = ps.tie_pixels(g["prob"])
    assert ties.sum() > 0
    err = np.abs(depth[:, 0] - g["depth"][:, 0])
    assert err[~ties].max() < 0.005 * 40
    # stable-descending tie rule reproduces the reference's CPU sort on the planted ties too
    assert (err[ties] < 0.005 * 40).mean() > 0.5
    lit = ps.extract_depth_torch(_t(g["prob"]), _t(g["d_batch"])).numpy()
    assert np.abs(lit - depth)[~np.isnan(lit)].max() < 1e-3


def test_batch_quirk_golden_with_unequal_d_min(golden_dir):
    """scripts/homography.py:26: the depth table is tiled V times along dim 0 while views are ordered b*V+v, so flat view i reads
    depth row i mod B.  Golden with d_min = (425, 520) from the unmodified reference: the oracle reproduces it in bug-compatible
    mode and does NOT with the geometrically intended rows."""
    g = _load(golden_dir, "bquirk_b2v3")
    assert float(g["d_min"].ravel()[0]) != float(g["d_min"].ravel()[1])
    args = (_t(g["feat"]), g["K"], g["R"], g["T"], _t(g["d_min"]), _t(g["d_int"]), 2, 3, int(g["D"]), int(g["d_scale"]))
    cost, warped, _ = ps.plane_sweep_cost(*args)
    assert _relmax(warped, g["warped"]) < 1e-4 and _relmax(cost, g["cost"]) < 1e-4
    cost_fixed, _, _ = ps.plane_sweep_cost(*args, bug_compatible=False)
    assert _relmax(cost_fixed, g["cost"]) > 1e-2
    w0 = {k: _t(v).clone() for k, v in np.load(os.path.join(golden_dir, "reg_weights.npz")).items()}
    prob = ps.reg_forward(w0, _t(g["cost"]), train_bn=True)
    assert _relmax(prob.numpy(), g["prob"]) < 1e-4


def test_fullsize_digest_inputs_regenerate_and_oracle_was_pinned_there(golden_dir):
    """tests/golden/cfg1_digest.npz (the reference at BASELINE size, oracle/make_golden.py --fullsize): its seeded inputs
    regenerate bit-identically, the kept ranks stored are consistent with the stored depth map, and the oracle's restatement
    agreed with the reference at that size when the digest was made."""
    g = _load(golden_dir, "cfg1_digest")
    feat = ps.smooth_features_exact(3, 32, 128, 160, int(g["seed"]))
    assert np.array_equal(feat[:, ::11, ::37, ::41].numpy(), g["feat_probe"])
    assert abs(feat.double().sum().item() - float(g["feat_sum"])) <= 1e-9 * float(g["feat_abs_sum"])
    K, R, T = ps.synthetic_cameras(1, 3, 128, 160, seed=int(g["seed"]))
    assert np.array_equal(K.numpy(), g["K"]) and np.array_equal(R.numpy(), g["R"]) and np.array_equal(T.numpy(), g["T"])
    assert float(g["oracle_prob_relerr"]) < 1e-6 and float(g["oracle_depth_abserr"]) < 0.005 * float(g["d_scale"])
    assert g["ranks"].shape == (1, 5, 128, 160) and int(g["ties"].sum()) < 20
    lo, hi = float(g["d_batch"].min()), float(g["d_batch"].max())
    assert lo <= float(g["depth"].min()) and float(g["depth"].max()) <= hi
    # cost samples against the oracle's sampler on the same planes (the CPU oracle at full size, sampled)
    planes = g["planes"].tolist()
    prm = ps.view_params_closed64(g["K"], g["R"], g["T"], 1, 3, 128, 160)
    ix, iy = ps.sample_positions_closed64(prm, g["d_batch"].reshape(1, -1)[[0, 0, 0]][:, planes], 128, 160)
    ref = ps.variance_cost(ps.bilinear_grid_sample(feat, _t(ix), _t(iy)), 3).numpy()
    assert np.abs(ref[:, g["cost_ch"].tolist()] - g["cost"]).max() / float(g["cost_absmax"]) < 1e-4
