#!/usr/bin/env python
"""The 2D networks' convolution layers on the tcgen05 kernels (mvs_b200/nets2d.py), layer by layer at cfg2's shapes: forward, data
gradient and weight gradient in microseconds next to cuDNN's bf16 channels-last kernels (CUDA events, GPU kept busy while the host
enqueues).  python tools/bench_nets2d.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deep-multiview-depth-estimation_b200"))
import torch
import torch.nn.functional as F
from mvs_b200 import nets2d, _lib
DEV = "cuda:0"
torch.backends.cudnn.benchmark = True


def timeit(fn, reps=7):
    for _ in range(3): fn()
    ts = []
    flush = torch.empty(64 * 1024 * 1024, device=DEV)
    for _ in range(reps):
        for _ in range(20):
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return 1e3 * sorted(ts)[len(ts) // 2]


LAYERS = [("enc L0 3->8 @512x640", 12, 512, 640, 3, 8, 8, 8), ("enc L1 8->8 @512x640", 12, 512, 640, 8, 8, 8, 8),
          ("enc L2 (5x5 s2 as 32->16) @256x320", 12, 256, 320, 32, 16, 32, 16), ("enc L3 16->16 @256x320", 12, 256, 320, 16, 16, 16, 16),
          ("enc L5 (5x5 s2 as 64->32) @128x160", 12, 128, 160, 64, 32, 64, 32), ("enc L6 32->32 @128x160", 12, 128, 160, 32, 32, 32, 32),
          ("refine 4->32 @128x160 x4", 4, 128, 160, 4, 32, 16, 32), ("refine 32->32 @128x160 x4", 4, 128, 160, 32, 32, 32, 32),
          ("refine 32->1 @128x160 x4", 4, 128, 160, 32, 1, 32, 8)]
for name, N, H, W, ci, co, cx, cy in LAYERS:
    x = torch.zeros(1, cx, N, H, W, device=DEV, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
    x[:, :ci] = torch.randn(1, ci, N, H, W, device=DEV).to(torch.bfloat16)
    gy = torch.zeros(1, cy, N, H, W, device=DEV, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
    gy[:, :co] = torch.randn(1, co, N, H, W, device=DEV).to(torch.bfloat16)
    w = torch.randn(co, ci, 3, 3, device=DEV) / (3 * ci ** 0.5)
    wk_f = nets2d._pack2d(w, "fwd", nets2d._n_rows(cy), max(cx, 16))
    wk_d = nets2d._pack2d(w, "dgrad", nets2d._n_rows(cx), max(cy, 16))
    gw27 = torch.empty((27, max(cx, 16), cy), dtype=torch.float32, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    fwd = lambda: nets2d._conv_rows(x, wk_f, cy, None)
    dgr = lambda: nets2d._conv_rows(gy, wk_d, cx, None)
    wgr = lambda: _lib.call("mvsb200_conv3d_s1_wgrad_ex", x.data_ptr(), gy.data_ptr(), gw27.data_ptr(), 1, N, H, W, cx, N, H, W, cy, -1, -1, -1, 2, st)
    # cuDNN on the same logical layer (channels-last bf16)
    x4 = x[0, :ci].permute(1, 0, 2, 3).contiguous(memory_format=torch.channels_last)
    g4 = gy[0, :co].permute(1, 0, 2, 3).contiguous(memory_format=torch.channels_last)
    w4 = w.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    lib_f = lambda: F.conv2d(x4, w4, padding=1)
    lib_b = lambda: torch.ops.aten.convolution_backward(g4, x4, w4, None, (1, 1), (1, 1), (1, 1), False, (0, 0), 1, (True, True, False))
    print(json.dumps(dict(layer=name, fwd_us=round(timeit(fwd), 1), dgrad_us=round(timeit(dgr), 1), wgrad_us=round(timeit(wgr), 1),
                          cudnn_fwd_us=round(timeit(lib_f), 1), cudnn_bwd_us=round(timeit(lib_b), 1),
                          MB_in_out=round((cx + cy) * N * H * W * 2 / 1e6, 1))), flush=True)
