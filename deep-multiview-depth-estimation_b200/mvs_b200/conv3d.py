"""Convolution backends for the regulariser (k = 3, no bias; scripts/model.py:223-234).

  "cudnn"   -- torch.nn.functional (cuDNN on the GPU).  Library path; also what the CPU-side unit tests
               of the canvas algebra use.
  "tcgen05" -- sm_100a implicit-GEMM kernels of libmvs_b200.so (registered by mvs_b200.conv3d_sm100 when
               the library exports them).
  "auto"    -- tcgen05 where available, else cudnn.
"""
from __future__ import annotations

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


_BACKENDS = {"cudnn": TorchConvBackend}


def register(name, backend):
    _BACKENDS[name] = backend


def get(name="auto"):
    if name == "auto":
        return _BACKENDS.get("tcgen05", TorchConvBackend)
    if hasattr(name, "conv3d"):
        return name
    return _BACKENDS[name]
