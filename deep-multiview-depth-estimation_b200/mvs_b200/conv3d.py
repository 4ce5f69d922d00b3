"""Convolution backends for the regulariser (k = 3, no bias; scripts/model.py:223-234).

  "cudnn"   -- torch.nn.functional (cuDNN on the GPU).  Library path; also what the CPU-side unit tests
               of the canvas algebra use.
  "tcgen05" -- sm_100a implicit-GEMM kernels of libmvs_b200.so (registered by mvs_b200.conv3d_sm100 when
               the library exports them).
  "auto"    -- tcgen05 where available, else cudnn.
  "fp32"    -- the library with TF32 switched off inside the call, forward and backward (ExactTorchConvBackend): what
               CostVolumeReg(precision="fp32") uses, so that its 1e-4 parity does not hang on a process-wide torch flag.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


class TorchConvBackend:
    name = "cudnn"

    @staticmethod
    def conv3d(x, w, stride, padding):
        return F.conv3d(x, w, None, stride, padding)

    @staticmethod
    def conv_transpose3d_alloc(x, w, stride, padding, out_dims):
        """Stride-2 transposed conv from the central box; the result holds the canvas `out_dims` at its origin and may be
        up to one plane/line/column larger (callers that can address the canvas inside it avoid a crop copy)."""
        size = [stride * (m - 1) - 2 * p + 3 for m, p in zip(x.shape[-3:], padding)]
        opad = tuple(max(0, n - s) for n, s in zip(out_dims, size))
        return F.conv_transpose3d(x, w, None, stride, tuple(padding), opad)

    @classmethod
    def conv_transpose3d(cls, x, w, stride, padding, out_dims):
        """Stride-2 transposed conv from the central box to the full canvas `out_dims` (cropped to it)."""
        D, h, w_ = out_dims
        return cls.conv_transpose3d_alloc(x, w, stride, padding, out_dims)[..., :D, :h, :w_]


def _no_tf32():
    c = torch.backends.cudnn
    return c.flags(enabled=c.enabled, benchmark=c.benchmark, deterministic=c.deterministic, allow_tf32=False)


class _ExactConvolution(torch.autograd.Function):
    """aten.convolution / aten.convolution_backward with TF32 off in BOTH directions (autograd would run the backward of a
    plain F.conv3d outside any flag context the forward was called under)."""

    @staticmethod
    def forward(ctx, x, w, stride, padding, transposed, output_padding):
        ctx.save_for_backward(x, w)
        ctx.cfg = (list(stride), list(padding), bool(transposed), list(output_padding))
        with _no_tf32():
            return torch.ops.aten.convolution(x, w, None, ctx.cfg[0], ctx.cfg[1], [1, 1, 1], ctx.cfg[2], ctx.cfg[3], 1)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        stride, padding, transposed, output_padding = ctx.cfg
        with _no_tf32():
            gx, gw, _ = torch.ops.aten.convolution_backward(gy.contiguous(), x, w, None, stride, padding, [1, 1, 1], transposed,
                                                            output_padding, 1, [ctx.needs_input_grad[0], ctx.needs_input_grad[1], False])
        return gx, gw, None, None, None, None


class ExactTorchConvBackend(TorchConvBackend):
    name = "fp32"

    @staticmethod
    def conv3d(x, w, stride, padding):
        pad = tuple(padding) if isinstance(padding, (tuple, list)) else (padding,) * 3
        return _ExactConvolution.apply(x, w, (stride,) * 3, pad, False, (0, 0, 0))

    @staticmethod
    def conv_transpose3d_alloc(x, w, stride, padding, out_dims):
        size = [stride * (m - 1) - 2 * p + 3 for m, p in zip(x.shape[-3:], padding)]
        opad = tuple(max(0, n - s) for n, s in zip(out_dims, size))
        return _ExactConvolution.apply(x, w, (stride,) * 3, tuple(padding), True, opad)


_BACKENDS = {"cudnn": TorchConvBackend, "fp32": ExactTorchConvBackend}


def register(name, backend):
    _BACKENDS[name] = backend


def get(name="auto"):
    if name == "auto":
        return _BACKENDS.get("tcgen05", TorchConvBackend)
    if hasattr(name, "conv3d"):
        return name
    return _BACKENDS[name]
