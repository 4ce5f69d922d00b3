"""Lazy stand-in for the tensor homography_warping returns (scripts/homography.py:92).

The reference materialises N = B*V warped volumes [N,C,D,h,w] (1.5 GB per sample at D = 192) only to
reduce them to a variance right away (scripts/model.py:177-181).  The drop-in returns this handle instead;
`assemble_cost_volume` recognises it and launches the fused kernel.  Anything else that touches it as a
tensor gets the real volumes (materialised by the parity/debug kernel, no autograd).
"""
from __future__ import annotations

import torch

from . import ops


class WarpedFeatureVolumes:
    def __init__(self, feature_maps: torch.Tensor, sweep: ops.PlaneSweep):
        self.feature_maps = feature_maps
        self.sweep = sweep
        n, c, h, w = feature_maps.shape
        self.shape = torch.Size((n, c, sweep.D, h, w))
        self.device, self.dtype = feature_maps.device, torch.float32

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def dim(self):
        return 5

    def materialize(self) -> torch.Tensor:
        return ops.warp_materialize(self.feature_maps, self.sweep)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        conv = lambda a: a.materialize() if isinstance(a, WarpedFeatureVolumes) else a
        return func(*[conv(a) for a in args], **{k: conv(v) for k, v in kwargs.items()})

    def __getattr__(self, name):          # .reshape / .sum / ... -> behave like the materialised tensor
        if name.startswith("__"):
            raise AttributeError(name)
        return getattr(self.materialize(), name)

    def __repr__(self):
        return f"WarpedFeatureVolumes(shape={tuple(self.shape)}, lazy, device={self.device})"
