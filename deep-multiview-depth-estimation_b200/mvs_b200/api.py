"""The reference's three hot-path functions with explicit configuration (no import-time globals).

dropin/{homography,costvolume,depthmap}.py adapt these to the reference's exact signatures and to its
`config` module.
"""
from __future__ import annotations

import torch

from . import ops
from .handle import WarpedFeatureVolumes


def homography_warping(K_batch, R_batch, T_batch, d_min, d_int, feature_maps, batch_size, n_views, d_num,
                       d_scale, bug_compatible=True, sweep=None):
    """scripts/homography.py:6-92.  Returns (warped handle, d_batch_0 [B,D,1,1] on the feature device,
    ref_idx_0 [B] CPU int64)."""
    if feature_maps.dim() != 4:
        raise ValueError(f"Input src must be a BxCxHxW tensor. Got {tuple(feature_maps.shape)}")   # kornia's check
    n, _, h, w = feature_maps.shape
    if n != batch_size * n_views:
        raise ValueError(f"feature_maps has {n} maps, expected batch_size*n_views = {batch_size * n_views}")
    if sweep is None:                                      # else: geometry already resident (ops.PlaneSweep.update)
        sweep = ops.PlaneSweep(K_batch, R_batch, T_batch, d_min, d_int, batch_size, n_views, int(d_num), d_scale,
                               h, w, feature_maps.device, bug_compatible)
    return WarpedFeatureVolumes(feature_maps, sweep), sweep.d_batch_dev, torch.arange(0, n, n_views)


def assemble_cost_volume(warped_feature_maps, n_views, out_dtype=torch.float32):
    """scripts/costvolume.py:3-16."""
    if isinstance(warped_feature_maps, WarpedFeatureVolumes):
        if n_views != warped_feature_maps.sweep.V:
            raise ValueError("n_views differs from the homography_warping call that produced this volume")
        return ops.warp_variance(warped_feature_maps.feature_maps, warped_feature_maps.sweep, out_dtype)
    return ops.variance_views(warped_feature_maps, n_views)


def extract_depth_map(prob_volume, d_batch, n_depth_est=5):
    """scripts/depthmap.py:4-22."""
    return ops.depth_from_prob(prob_volume, d_batch, int(n_depth_est))
