#!/usr/bin/env python
"""Bring-up check of the tcgen05 weight-gradient kernel against torch.nn.grad.conv3d_weight (fp32)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deep-multiview-depth-estimation_b200"))
import torch  # noqa: E402
from mvs_b200 import _lib  # noqa: E402

DEV = "cuda:0"
torch.backends.cudnn.allow_tf32 = False


def run(B, Cin, Cout, D, h, w, pad):
    g = torch.Generator().manual_seed(Cin * 1000 + Cout * 10 + D)
    x = torch.randn(B, Cin, D, h, w, generator=g).to(DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
    Do, Ho, Wo = (D, h, w) if pad else (D - 2, h - 2, w - 2)
    gy = torch.randn(B, Cout, Do, Ho, Wo, generator=g).to(DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
    gw = torch.full((27, Cin, Cout), float("nan"), device=DEV)
    off = -1 if pad else 0
    _lib.call("mvsb200_conv3d_s1_wgrad", x.data_ptr(), gy.data_ptr(), gw.data_ptr(), B, D, h, w, Cin, Do, Ho, Wo, Cout, off, off, off,
              torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv3d_weight(x.float(), (Cout, Cin, 3, 3, 3), gy.float(), padding=1 if pad else 0)
    ours = gw.reshape(3, 3, 3, Cin, Cout).permute(4, 3, 0, 1, 2)
    err = (ours - ref).abs().max().item() / ref.abs().max().item()
    print(f"B={B} Cin={Cin} Cout={Cout} D={D} h={h} w={w} pad={pad}: rel err {err:.3e} nan {int(torch.isnan(gw).sum())}", flush=True)


if __name__ == "__main__":
    for c in [(1, 32, 8, 4, 6, 12, 1), (1, 32, 32, 4, 6, 12, 1), (1, 16, 16, 4, 6, 12, 1), (1, 64, 64, 5, 9, 20, 0),
              (2, 32, 8, 7, 33, 47, 1), (1, 64, 32, 6, 20, 40, 1), (1, 16, 16, 9, 64, 80, 0), (1, 32, 16, 5, 17, 23, 1)]:
        try:
            run(*c)
        except Exception as e:  # noqa: BLE001
            print("FAILED", c, e, flush=True)
