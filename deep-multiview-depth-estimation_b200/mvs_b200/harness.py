"""Measurement harness around the hot path: the pieces of the reference that are OUT OF SCOPE for the
native build (2D feature encoder, depth refinement, loss, optimiser step) written with
stock torch.nn so that bench.py can time a whole MVSNet forward / train step on synthetic DTU-shaped data.
Layer shapes follow /root/reference/scripts/model.py:22-65 (encoder), :129-152 (refinement), :168-207 (glue)
and scripts/loss.py:4-41.  Only the four hot-path calls in `forward` are mvs_b200 code.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import api
from .regulariser import CostVolumeReg


def _conv_bn_relu(i, o, k, s):
    return [nn.Conv2d(i, o, k, stride=s, padding=k // 2, bias=False), nn.BatchNorm2d(o), nn.ReLU()]


class FeatureEncoder(nn.Module):
    """[N,3,H,W] -> [N,32,H/4,W/4]: 3-3-5(s2)-3-3-5(s2)-3-3 convs, 8/16/32 channels, no BN/ReLU on the last."""

    def __init__(self, in_ch=3, base=8):
        super().__init__()
        c1, c2, c3 = base, base * 2, base * 4
        layers = (_conv_bn_relu(in_ch, c1, 3, 1) + _conv_bn_relu(c1, c1, 3, 1) + _conv_bn_relu(c1, c2, 5, 2)
                  + _conv_bn_relu(c2, c2, 3, 1) + _conv_bn_relu(c2, c2, 3, 1) + _conv_bn_relu(c2, c3, 5, 2)
                  + _conv_bn_relu(c3, c3, 3, 1) + [nn.Conv2d(c3, c3, 3, padding=1, bias=False)])
        self.model = nn.Sequential(*layers)

    def forward(self, x):
        return self.model(x)


class DepthRefinement(nn.Module):
    """[B,4,h,w] (normalised depth + resized reference image) -> residual-refined normalised depth."""

    def __init__(self, in_ch=4, base=32):
        super().__init__()
        self.model = nn.Sequential(*(_conv_bn_relu(in_ch, base, 3, 1) + _conv_bn_relu(base, base, 3, 1)
                                     + _conv_bn_relu(base, base, 3, 1) + [nn.Conv2d(base, 1, 3, padding=1, bias=False)]))

    def forward(self, x):
        return self.model(x) + x[:, :1]


class MVSNet(nn.Module):
    """The reference's MVSNet wiring (model.py:168-207) around the mvs_b200 hot path."""

    def __init__(self, d_num, d_scale, precision="bf16", n_depth_est=5, conv_backend="auto"):
        super().__init__()
        self.d_num, self.d_scale, self.precision, self.n_depth_est = int(d_num), d_scale, precision, int(n_depth_est)
        self.feature_encoder = FeatureEncoder()
        self.cost_volume_reg = CostVolumeReg(precision=precision, conv_backend=conv_backend, n_depth_est=n_depth_est)
        self.depthmap_refine = DepthRefinement()

    def forward(self, nn_input, K_batch, R_batch, T_batch, d_min, d_int, batch_size, n_views):
        amp = self.precision == "bf16" and nn_input.is_cuda
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):        # out-of-scope 2D net: stock torch AMP
            feats = self.feature_encoder(nn_input.contiguous(memory_format=torch.channels_last))
        feats = feats.float()
        warped, d_batch, ref_views = api.homography_warping(K_batch, R_batch, T_batch, d_min, d_int, feats,
                                                            batch_size, n_views, self.d_num, self.d_scale)
        vol_dtype = torch.bfloat16 if self.precision == "bf16" else torch.float32
        cost = api.assemble_cost_volume(warped, n_views, vol_dtype)
        prob = self.cost_volume_reg(cost)
        initial = api.extract_depth_map(prob, d_batch, self.n_depth_est)
        dev = initial.device
        d_trans = d_min.to(dev)
        d_span = d_int.to(dev) * self.d_num * self.d_scale
        norm = (initial - d_trans) / d_span
        h, w = initial.shape[-2:]
        ref_img = F.interpolate(nn_input[ref_views.to(dev)], (h, w), mode="bilinear", align_corners=False)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
            refined = self.depthmap_refine(torch.cat((norm, ref_img), 1))
        refined = refined.float() * d_span + d_trans
        return initial, refined


def loss_fcn(gt, initial, refined):
    """Masked L1 on both depth maps (scripts/loss.py:4-41): returns (loss, initial MAE, refined MAE)."""
    mask = (gt != 0).float()
    n_valid = mask.sum((1, 2, 3))
    l0 = (mask * (gt - initial).abs()).sum((1, 2, 3)) / n_valid
    l1 = (mask * (gt - refined).abs()).sum((1, 2, 3)) / n_valid
    return (l0 + l1).sum(), l0.mean(), l1.mean()


class FlatGradAllReduce:
    """Data-parallel gradient reduction for scene/batch sharding: one flat fp32 bucket holding
    every gradient (1.53 MB for MVSNet), one all-reduce per step, averaged.  Works on any process group
    (NCCL on the GPUs, gloo in the CPU tests).  The reference has no distributed code; its parameters live in
    a plain list (model.py:164-166), so this takes a list, not a DDP-wrapped module."""

    def __init__(self, params, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.params = [p for p in params if p.requires_grad]
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.bucket = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views, o = [], 0
        for p in self.params:
            self.views.append(self.bucket[o:o + p.numel()].view_as(p))
            o += p.numel()

    def broadcast_parameters(self, buffers=()):
        if self.world == 1:
            return
        for t in list(self.params) + list(buffers):
            self.dist.broadcast(t.data if isinstance(t, nn.Parameter) else t, 0, group=self.group)

    def reduce(self):
        """Call after backward(): grads <- mean over ranks."""
        if self.world == 1:
            return
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            else:
                v.copy_(p.grad)
        self.dist.all_reduce(self.bucket, group=self.group)
        self.bucket.div_(self.world)
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                p.grad = v.clone()
            else:
                p.grad.copy_(v)
