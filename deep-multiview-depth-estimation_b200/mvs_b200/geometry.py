"""Host-side (fp64) plane-sweep geometry: cameras -> the 16 floats per view + 1/(d - s) table the kernels read.

Follows /root/reference/scripts/homography.py:23-75 including its quirks:
  * plane normal n = 3rd COLUMN of R_ref (:49)
  * the depth table is tiled V times along dim 0 (:26) while views are ordered b*V+v, so flat view i
    reads depth row i mod B ("batch quirk"; harmless when all d_min are equal, as on DTU)
  * kornia's warp_perspective inverts the matrix it is given and resamples by
    ix = px*w/(w-1) - 0.5 (align_corners mismatch)
  * a plane at d == 0 (validate.py:40 sweeps from 0) divides by zero => NaN plane; marked by tinv = NaN
"""
from __future__ import annotations

import numpy as np
import torch

from ._lib import VIEW_PARAM_FLOATS


def depth_table(d_min: torch.Tensor, d_int: torch.Tensor, d_num: int, d_scale) -> torch.Tensor:
    """d_batch_0 [B,D,1,1] fp32, computed with the reference's own expression (homography.py:24-25)."""
    k = torch.arange(d_num).reshape(1, d_num, 1, 1)
    return d_min.detach().cpu() + d_scale * d_int.detach().cpu() * k


def view_tables(K, R, T, d_batch_0, batch_size, n_views, h, w, bug_compatible=True):
    """-> (view_params [N,16] float32, tinv [N,D] float32) as numpy arrays."""
    N = batch_size * n_views
    K = np.asarray(K.detach().cpu().numpy() if isinstance(K, torch.Tensor) else K, dtype=np.float64).reshape(N, 3, 3)
    R = np.asarray(R.detach().cpu().numpy() if isinstance(R, torch.Tensor) else R, dtype=np.float64).reshape(N, 3, 3)
    T = np.asarray(T.detach().cpu().numpy() if isinstance(T, torch.Tensor) else T, dtype=np.float64).reshape(N, 3, 1)
    depths = np.asarray(d_batch_0.detach().cpu().numpy(), dtype=np.float64).reshape(batch_size, -1)
    ref = (np.arange(N) // n_views) * n_views
    Rt = np.transpose(R, (0, 2, 1))
    C = -Rt @ T                                               # camera centres           (:48,:58)
    RrKr = Rt[ref] @ np.linalg.inv(K[ref])                    # R_ref^T K_ref^-1          (:64-65)
    KR = K @ R                                                # K_i R_i                   (:61)
    A = KR @ RrKr
    u = KR @ (C - C[ref])                                     # N,3,1
    wT = np.transpose(R[ref][:, :, 2:3], (0, 2, 1)) @ RrKr    # N,1,3  (n = R_ref[:,2] as a row, :49)
    Ainv = np.linalg.inv(A)
    g = Ainv @ u                                              # N,3,1
    r = wT @ Ainv                                             # N,1,3
    s = (wT @ g).reshape(N)
    S = np.diag([w / (w - 1.0) if w > 1 else 1.0, h / (h - 1.0) if h > 1 else 1.0, 1.0])
    params = np.zeros((N, VIEW_PARAM_FLOATS), dtype=np.float64)
    params[:, 0:9] = (S @ Ainv).reshape(N, 9)
    params[:, 9:12] = (S @ g).reshape(N, 3)
    params[:, 12:15] = r.reshape(N, 3)
    rows = (np.arange(N) % batch_size) if bug_compatible else (np.arange(N) // n_views)
    dv = depths[rows]                                         # N,D
    with np.errstate(divide="ignore", invalid="ignore"):
        tinv = 1.0 / (dv - s[:, None])
    tinv = np.where(dv == 0.0, np.nan, tinv)
    tinv = np.where(np.isinf(tinv), np.float64(3.0e38), tinv)   # d == s: position at infinity -> out of bounds
    return params.astype(np.float32), tinv.astype(np.float32)
