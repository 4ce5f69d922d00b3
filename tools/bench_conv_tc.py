#!/usr/bin/env python
"""Isolated timing of the tcgen05 conv3d kernel at the regulariser's real shapes, next to cuDNN (bf16, channels_last_3d).
One JSON line per layer: ms, algorithmic TFLOP/s, fraction of the measured bf16 tensor peak, activation GB/s."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deep-multiview-depth-estimation_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from mvs_b200 import _lib  # noqa: E402
from tools.test_conv_tc import pack  # noqa: E402

DEV = "cuda:0"
PEAK_TF, PEAK_GB = 1654.0, 6533.5
try:
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    PEAK_TF, PEAK_GB = float(pk["bf16_tflops"]), float(pk["hbm_gbs"])
except Exception:
    pass


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    ts = []
    flush = torch.empty(64 * 1024 * 1024, device=DEV)
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def layer(name, B, Cin, Cout, D, h, w, pad):
    x = torch.randn(B, Cin, D, h, w, device=DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
    wt = (torch.randn(Cout, Cin, 3, 3, 3, device=DEV) / (27 * Cin) ** 0.5).to(torch.bfloat16)
    n_rows = 16 if Cout <= 16 else (32 if Cout <= 32 else 64)
    wp = pack(wt, n_rows)
    Do, Ho, Wo = (D, h, w) if pad else (D - 2, h - 2, w - 2)
    y = torch.empty((B, Cout, Do, Ho, Wo), dtype=torch.bfloat16, device=DEV).contiguous(memory_format=torch.channels_last_3d)
    off = -1 if pad else 0
    st = torch.cuda.current_stream().cuda_stream
    ours = lambda: _lib.call("mvsb200_conv3d_s1_fwd", x.data_ptr(), wp.data_ptr(), y.data_ptr(), B, D, h, w, Cin, Do, Ho, Wo,
                             Cout, Cout, n_rows, off, off, off, st)
    wk = wp.view(3, 3, 3, n_rows, Cin).permute(1, 2, 0, 3, 4).contiguous()
    y2 = torch.empty_like(y)
    kdn = lambda: _lib.call("mvsb200_conv3d_s1_fwd_kdn", x.data_ptr(), wk.data_ptr(), y2.data_ptr(), B, D, h, w, Cin, Do, Ho, Wo,
                            Cout, Cout, n_rows, off, off, off, st)
    t_kdn = timeit(kdn)
    wt_cl = wt.contiguous(memory_format=torch.channels_last_3d)
    cudnn = lambda: F.conv3d(x, wt_cl, padding=1 if pad else 0)
    t_ours, t_cudnn = timeit(ours), timeit(cudnn)
    ref = cudnn()
    err = (y.float() - ref.float()).abs().max().item() / ref.float().abs().max().item()
    flops = 2.0 * 27 * Cin * Cout * B * Do * Ho * Wo
    byts = 2.0 * B * (Cin * D * h * w + Cout * Do * Ho * Wo)
    # weight gradient
    gy = torch.randn_like(y)
    gw = torch.empty(27, Cin, Cout, device=DEV)
    wg = lambda: _lib.call("mvsb200_conv3d_s1_wgrad", x.data_ptr(), gy.data_ptr(), gw.data_ptr(), B, D, h, w, Cin, Do, Ho, Wo, Cout,
                           off, off, off, st)
    wg_cudnn = lambda: torch.nn.grad.conv3d_weight(x, wt.shape, gy, padding=1 if pad else 0)
    t_wg, t_wg_cudnn = (timeit(wg), timeit(wg_cudnn)) if Cout in (8, 16, 32, 64) else (None, None)
    err_kdn = (y2.float() - ref.float()).abs().max().item() / ref.float().abs().max().item()
    print(json.dumps(dict(layer=name, ms_kdn=t_kdn, TFLOPs_kdn=flops / t_kdn / 1e9, frac_tensor_peak_kdn=flops / t_kdn / 1e9 / PEAK_TF,
                          frac_hbm_kdn=byts / t_kdn / 1e6 / PEAK_GB, rel_err_kdn=err_kdn, wgrad_ms=t_wg, wgrad_ms_cudnn=t_wg_cudnn, wgrad_TFLOPs=(flops / t_wg / 1e9) if t_wg else None, Cin=Cin, Cout=Cout, out=[B, Do, Ho, Wo], ms=t_ours, ms_cudnn=t_cudnn, rel_err_vs_cudnn=err,
                          TFLOPs=flops / t_ours / 1e9, frac_tensor_peak=flops / t_ours / 1e9 / PEAK_TF,
                          act_GBps=byts / t_ours / 1e6, frac_hbm=byts / t_ours / 1e6 / PEAK_GB)), flush=True)


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    only = sys.argv[2] if len(sys.argv) > 2 else None
    LAYERS = [("conv_0_0", 32, 8, 192, 128, 160, 1), ("conv_1_1", 16, 16, 101, 69, 85, 0), ("conv_2_1", 32, 32, 101, 69, 85, 0),
              ("conv_3_1", 64, 64, 101, 69, 85, 0), ("dense32x32", 32, 32, 192, 128, 160, 1), ("dense64x64", 64, 64, 96, 128, 160, 1)]
    for name, ci, co, D, h, w, pad in LAYERS:
        if only is None or only == name:
            layer(name, B, ci, co, D, h, w, pad)
