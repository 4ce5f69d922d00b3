#!/usr/bin/env python
"""Bring-up check of the tcgen05 conv3d kernel against torch's fp32 conv on bf16-rounded operands."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deep-multiview-depth-estimation_b200"))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from mvs_b200 import _lib  # noqa: E402

DEV = "cuda:0"
torch.backends.cudnn.allow_tf32 = False


def pack(w, n_rows):
    co, ci = w.shape[:2]
    wp = torch.zeros(27, n_rows, ci, dtype=torch.bfloat16, device=w.device)
    wp[:, :co] = w.permute(2, 3, 4, 0, 1).reshape(27, co, ci).to(torch.bfloat16)
    return wp.contiguous()


def run(B, Cin, Cout, D, h, w, pad, mode):
    g = torch.Generator().manual_seed(Cin * 1000 + Cout * 10 + D)
    x = torch.randn(B, Cin, D, h, w, generator=g).to(DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
    wt = (torch.randn(Cout, Cin, 3, 3, 3, generator=g) / (27 * Cin) ** 0.5).to(DEV).to(torch.bfloat16)
    n_rows = 16 if Cout <= 16 else (32 if Cout <= 32 else 64)
    wp = pack(wt, n_rows)
    Do, Ho, Wo = (D, h, w) if pad else (D - 2, h - 2, w - 2)
    y = torch.full((B, Cout, Do, Ho, Wo), float("nan"), dtype=torch.bfloat16, device=DEV).contiguous(memory_format=torch.channels_last_3d)
    off = -1 if pad else 0
    _lib.call("mvsb200_conv3d_s1_fwd", x.data_ptr(), wp.data_ptr(), y.data_ptr(), B, D, h, w, Cin, Do, Ho, Wo, Cout, Cout,
              n_rows, off, off, off, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ref = F.conv3d(x.float(), wt.float(), padding=1 if pad else 0)
    err = (y.float() - ref).abs().max().item() / ref.abs().max().item()
    nan = int(torch.isnan(y.float()).sum())
    print(f"B={B} Cin={Cin} Cout={Cout} D={D} h={h} w={w} pad={pad} mode={mode}: rel err {err:.3e} nan {nan}", flush=True)
    return err


if __name__ == "__main__":
    cases = [(1, 32, 32, 4, 6, 12, 1), (1, 64, 32, 4, 6, 12, 1), (1, 16, 16, 4, 6, 12, 1), (1, 32, 8, 5, 9, 50, 1),
             (2, 32, 32, 7, 33, 47, 1), (1, 64, 64, 6, 20, 40, 0), (1, 16, 16, 9, 64, 80, 1)]
    for mode in (0, 1):
        for c in cases:
            try:
                run(*c, mode)
            except Exception as e:  # noqa: BLE001
                print("FAILED", c, mode, e, flush=True)
