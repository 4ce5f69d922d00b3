"""Host fp64 geometry of the product (mvs_b200.geometry) vs the oracle and the reference's own homographies."""
import os

import numpy as np
import pytest
import torch

import plane_sweep as ps
from mvs_b200 import geometry

CASES = ["tiny_b1v3", "b2v3", "v5", "v7_odd"]


def _positions(params, tinv, h, w):
    N, D = tinv.shape
    ys, xs = np.meshgrid(np.arange(h, dtype=np.float64), np.arange(w, dtype=np.float64), indexing="ij")
    p = np.stack([xs, ys, np.ones_like(xs)], 0).reshape(3, -1)
    P = params.astype(np.float64)
    ix = np.empty((N, D, h, w)); iy = np.empty((N, D, h, w))
    for i in range(N):
        a = P[i, :9].reshape(3, 3) @ p
        c = P[i, 12:15] @ p
        q = a[None] + P[i, 9:12][None, :, None] * (c[None, None] * tinv[i].astype(np.float64)[:, None, None])
        ix[i] = (q[:, 0] / q[:, 2] - 0.5).reshape(D, h, w)
        iy[i] = (q[:, 1] / q[:, 2] - 0.5).reshape(D, h, w)
    return ix, iy


@pytest.mark.parametrize("name", CASES)
def test_view_tables_reproduce_reference_sampling_positions(golden_dir, name):
    g = dict(np.load(os.path.join(golden_dir, name + ".npz")))
    B, V, D, h, w = (int(g[k]) for k in "BVDhw")
    d0 = geometry.depth_table(torch.from_numpy(g["d_min"]), torch.from_numpy(g["d_int"]), D, int(g["d_scale"]))
    assert np.array_equal(d0.numpy(), g["d_batch"])
    params, tinv = geometry.view_tables(g["K"], g["R"], g["T"], d0, B, V, h, w)
    assert params.shape == (B * V, 16) and tinv.shape == (B * V, D) and params.dtype == np.float32
    ix, iy = _positions(params, tinv, h, w)
    jx, jy = ps.sample_positions_from_H(g["H"], h, w)           # what the reference's fp32 matrices imply
    assert np.abs(ix - jx).max() < 2e-3 and np.abs(iy - jy).max() < 2e-3
    prm = ps.view_params_closed64(g["K"], g["R"], g["T"], B, V, h, w)
    kx, ky = ps.sample_positions_closed64(prm, g["d_batch"].reshape(B, D)[ps.view_depth_rows(B, V)], h, w)
    assert np.abs(ix - kx).max() < 2e-4 and np.abs(iy - ky).max() < 2e-4   # fp32 rounding of the tables only


def test_batch_quirk_row_selection():
    K, R, T = ps.synthetic_cameras(2, 3, 16, 20, seed=5)
    d_min = torch.tensor([400.0, 500.0]).view(2, 1, 1, 1)
    d0 = geometry.depth_table(d_min, torch.ones(2, 1, 1, 1), 4, 10)
    _, t_bug = geometry.view_tables(K, R, T, d0, 2, 3, 16, 20, bug_compatible=True)
    _, t_fix = geometry.view_tables(K, R, T, d0, 2, 3, 16, 20, bug_compatible=False)
    H = ps.homographies_chain32(K, R, T, d0, 2, 3).numpy().astype(np.float64)       # reference op chain
    # recover the depth each flat view actually used: view i reads row i mod B (homography.py:26)
    rows = ps.view_depth_rows(2, 3)
    assert list(rows) == [0, 1, 0, 1, 0, 1]
    assert not np.allclose(t_bug, t_fix)
    params, _ = geometry.view_tables(K, R, T, d0, 2, 3, 16, 20)
    ix, _ = _positions(params, t_bug, 16, 20)
    jx, _ = ps.sample_positions_from_H(H, 16, 20)
    assert np.abs(ix - jx).max() < 2e-3


def test_zero_depth_plane_is_flagged_nan():
    K, R, T = ps.synthetic_cameras(1, 3, 16, 20)
    d0 = geometry.depth_table(torch.zeros(1, 1, 1, 1), torch.ones(1, 1, 1, 1), 4, 10)
    _, tinv = geometry.view_tables(K, R, T, d0, 1, 3, 16, 20)
    assert np.isnan(tinv[:, 0]).all() and np.isfinite(tinv[:, 1:]).all()


def test_plane_sweep_update_rewrites_the_same_buffers():
    """ops.PlaneSweep.update (what a captured CUDA graph relies on): new cameras / depth range land in the SAME tensors and
    equal the tables of a freshly built sweep."""
    import torch
    import plane_sweep as ps
    from mvs_b200 import ops
    B, V, D, h, w = 2, 3, 12, 16, 20
    K0, R0, T0 = ps.synthetic_cameras(B, V, h, w, seed=1)
    K1, R1, T1 = ps.synthetic_cameras(B, V, h, w, seed=2)
    d0, i0 = torch.full((B, 1, 1, 1), 425.0), torch.ones(B, 1, 1, 1)
    d1, i1 = torch.full((B, 1, 1, 1), 500.0), torch.full((B, 1, 1, 1), 1.5)
    sweep = ops.PlaneSweep(K0, R0, T0, d0, i0, B, V, D, 40.0, h, w, torch.device("cpu"))
    ptrs = (sweep.view_params.data_ptr(), sweep.tinv.data_ptr(), sweep.d_batch_dev.data_ptr())
    sweep.update(K1, R1, T1, d1, i1)
    fresh = ops.PlaneSweep(K1, R1, T1, d1, i1, B, V, D, 40.0, h, w, torch.device("cpu"))
    assert ptrs == (sweep.view_params.data_ptr(), sweep.tinv.data_ptr(), sweep.d_batch_dev.data_ptr())
    assert torch.equal(sweep.view_params, fresh.view_params) and torch.equal(sweep.tinv, fresh.tinv)
    assert torch.equal(sweep.d_batch_dev, fresh.d_batch_dev) and torch.equal(sweep.d_batch_0, fresh.d_batch_0)
    sweep.update(K0, R0, T0, d0, i0)                                    # and back
    first = ops.PlaneSweep(K0, R0, T0, d0, i0, B, V, D, 40.0, h, w, torch.device("cpu"))
    assert torch.equal(sweep.tinv, first.tinv)
