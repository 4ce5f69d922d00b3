"""tcgen05 convolution backend of the regulariser (K3): stride-1 3x3x3 convolutions of bf16 channel-last volumes run in
libmvs_b200.so's implicit-GEMM kernel (csrc/conv3d_tc.cu) -- forward AND the data gradient, which for a stride-1
convolution is the same convolution with the flipped, transposed filter.  Everything the kernel does not cover yet
(stride-2 branches, transposed convolutions, 8- and 1-channel operands, the weight gradient, fp32 volumes) goes to
the cuDNN backend; DESIGN.md §K3 lists which layer runs where.

Reference layers: scripts/model.py:223-234 (Conv3d factory), :101-113 (their use).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import _lib, conv3d as conv_backends
from .ops import _stream, _timed

_CIN_OK = (16, 32, 64)


def _n_rows(c):
    return 16 if c <= 16 else (32 if c <= 32 else 64)


def pack_filter(w: torch.Tensor) -> torch.Tensor:
    """[Cout, Cin, 3, 3, 3] -> bf16 [27, n_rows, Cin]: tap-major (kd, kh, kw), output channel (zero rows up to n_rows),
    input channel -- the K-major B operand the kernel keeps resident in shared memory."""
    co, ci = w.shape[:2]
    wp = torch.zeros(27, _n_rows(co), ci, dtype=torch.bfloat16, device=w.device)
    wp[:, :co] = w.detach().permute(2, 3, 4, 0, 1).reshape(27, co, ci).to(torch.bfloat16)
    return wp


def pack_filter_dgrad(w: torch.Tensor) -> torch.Tensor:
    """Filter of the data gradient: taps flipped, channel roles swapped ([Cin, Cout] per tap)."""
    return pack_filter(w.detach().flip(2, 3, 4).transpose(0, 1))


def _supported(cin, cout):
    return cin in _CIN_OK and cout % 8 == 0 and 8 <= cout <= 64


def _launch(x_cl, wp, cout, out_dims, off):
    B, cin, Di, Hi, Wi = x_cl.shape
    Do, Ho, Wo = out_dims
    y = torch.empty((B, cout, Do, Ho, Wo), dtype=torch.bfloat16, device=x_cl.device, memory_format=torch.channels_last_3d)
    with _timed("conv3d_s1_tc", 2.0 * 27 * cin * cout * B * Do * Ho * Wo):
        _lib.call("mvsb200_conv3d_s1_fwd", x_cl.data_ptr(), wp.data_ptr(), y.data_ptr(), B, Di, Hi, Wi, cin, Do, Ho, Wo,
                  cout, cout, wp.shape[1], off, off, off, _stream())
    return y


class _Conv3dS1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, pad):
        x_cl = x.detach().contiguous(memory_format=torch.channels_last_3d)
        D, H, W = x.shape[2:]
        out_dims = (D, H, W) if pad == 1 else (D - 2, H - 2, W - 2)
        y = _launch(x_cl, pack_filter(w), w.shape[0], out_dims, -pad)
        ctx.save_for_backward(x_cl, w)
        ctx.pad = pad
        return y

    @staticmethod
    def backward(ctx, gy):
        x_cl, w = ctx.saved_tensors
        pad = ctx.pad
        gx = gw = None
        gy = gy.to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
        if ctx.needs_input_grad[0]:
            cout, cin = w.shape[:2]
            off = -1 if pad == 1 else -2
            if _supported(cout, cin):          # roles swap: the gradient volume has Cout channels, the result Cin
                gx = _launch(gy, pack_filter_dgrad(w), cin, tuple(x_cl.shape[2:]), off)
            elif cout == 8 and _supported(16, cin):
                # 8-channel gradient rows (K = 8 < UMMA K): widen to 16 channels with zeros, zero filter rows to match
                B, _, Do, Ho, Wo = gy.shape
                gy16 = torch.empty((B, 16, Do, Ho, Wo), dtype=torch.bfloat16, device=gy.device,
                                   memory_format=torch.channels_last_3d)
                gy16[:, :8] = gy
                gy16[:, 8:] = 0
                w16 = torch.zeros((16,) + tuple(w.shape[1:]), dtype=w.dtype, device=w.device)
                w16[:8] = w.detach()
                gx = _launch(gy16, pack_filter_dgrad(w16), cin, tuple(x_cl.shape[2:]), off)
            else:
                gx = torch.nn.grad.conv3d_input(x_cl.shape, w.to(gy.dtype), gy, padding=pad)
        if ctx.needs_input_grad[1]:
            cout, cin = w.shape[:2]
            if cin in _CIN_OK and cout in (8, 16, 32, 64):
                B, _, Di, Hi, Wi = x_cl.shape
                Do, Ho, Wo = gy.shape[2:]
                gw27 = torch.empty((27, cin, cout), dtype=torch.float32, device=gy.device)
                with _timed("conv3d_s1_wgrad_tc", 2.0 * 27 * cin * cout * B * Do * Ho * Wo):
                    _lib.call("mvsb200_conv3d_s1_wgrad", x_cl.data_ptr(), gy.data_ptr(), gw27.data_ptr(), B, Di, Hi, Wi, cin,
                              Do, Ho, Wo, cout, -pad, -pad, -pad, _stream())
                gw = gw27.reshape(3, 3, 3, cin, cout).permute(4, 3, 0, 1, 2).to(w.dtype)
            else:
                gw = torch.nn.grad.conv3d_weight(x_cl, w.shape, gy, padding=pad).to(w.dtype)
        return gx, gw, None


class Tcgen05ConvBackend:
    name = "tcgen05"

    @staticmethod
    def conv3d(x, w, stride, padding):
        pad = tuple(padding) if isinstance(padding, (tuple, list)) else (padding,) * 3
        if (stride == 1 and x.is_cuda and x.dtype == torch.bfloat16 and pad in ((0, 0, 0), (1, 1, 1))
                and _supported(x.shape[1], w.shape[0]) and min(x.shape[2:]) >= 3):
            return _Conv3dS1.apply(x, w, pad[0])
        return conv_backends.TorchConvBackend.conv3d(x, w, stride, padding)

    conv_transpose3d = conv_backends.TorchConvBackend.conv_transpose3d
    conv_transpose3d_alloc = conv_backends.TorchConvBackend.conv_transpose3d_alloc


def available() -> bool:
    try:
        return hasattr(_lib.load(), "mvsb200_conv3d_s1_fwd")
    except _lib.MvsB200Error:
        return False


conv_backends.register("tcgen05", Tcgen05ConvBackend)
