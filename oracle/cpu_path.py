"""CPU timing leg of the oracle port  --  TEST / BASELINE INFRASTRUCTURE, NOT PRODUCT.

Runs the reference's algorithm for the hot path (as restated in oracle/plane_sweep.py: per-plane
grid_sample warp of all views, variance, dense-canvas CostVolumeReg with train-mode BN, sort-based depth
extraction; scripts/homography.py:23-92, costvolume.py:7-16, model.py:100-126, depthmap.py:11-19) on the
host cores with torch CPU ops, forward and -- for the train-step workload -- backward through autograd.
Used by bench.py for `cpu_baseline` and for `--impl reference` (kind = "port": the reference itself is
Python and cannot travel to the GPU box, /root/reference does not exist there).
"""
from __future__ import annotations

import time

import torch

import plane_sweep as ps


def make_sample(V=3, D=48, h=128, w=160, d_total=192, seed=0):
    """One batch item of the DTU-shaped workload restricted to the first D of d_total planes."""
    gen = torch.Generator().manual_seed(seed)
    K, R, T = ps.synthetic_cameras(1, V, h, w, seed=seed)
    return dict(V=V, D=D, h=h, w=w, d_scale=480.0 / d_total, K=K, R=R, T=T,
                d_min=torch.full((1, 1, 1, 1), 425.0), d_int=torch.ones(1, 1, 1, 1),
                feat=torch.randn(V, 32, h, w, generator=gen), sd=ps.reg_init(seed),
                gdepth=torch.randn(1, 1, h, w, generator=gen))


def hot_path_step(s, backward=True):
    """One pass of the hot path on the sample; returns (seconds, depth map)."""
    t0 = time.perf_counter()
    feat = s["feat"].clone().requires_grad_(backward)
    sd = {k: (v.clone().requires_grad_(backward) if v.is_floating_point() and "running" not in k else v.clone())
          for k, v in s["sd"].items()}
    with torch.set_grad_enabled(backward):
        warped, d0 = ps.warp_reference_chain(feat, s["K"], s["R"], s["T"], s["d_min"], s["d_int"], 1, s["V"], s["D"],
                                             s["d_scale"])
        cost = ps.variance_cost(warped, s["V"])
        prob = ps.reg_forward(sd, cost, train_bn=True)
        depth = ps.extract_depth_torch(prob, d0)
        if backward:
            (depth * s["gdepth"]).sum().backward()
    return time.perf_counter() - t0, depth.detach()
