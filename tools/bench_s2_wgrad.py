#!/usr/bin/env python
"""Strided weight gradients of the regulariser (B = 4, cfg2 shapes): tcgen05 parity-class kernel next to cuDNN."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deep-multiview-depth-estimation_b200"))
import torch
from mvs_b200 import conv3d_sm100 as c
from mvs_b200.regulariser import central_region
DEV = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dims = (192, 128, 160)
reg = [central_region(n) for n in dims]
box = [hi - lo + 1 for lo, hi, _ in reg]
pads = tuple(L for _, _, L in reg)

def timeit(fn, reps=5):
    for _ in range(2): fn()
    ts = []
    flush = torch.empty(64 * 1024 * 1024, device=DEV)
    for _ in range(reps):
        for _ in range(40):                  # ~2 ms of GPU work: the host enqueues fn() while the GPU is busy, so the events
            flush.zero_()                    # bracket device time only (fn's host side -- tensor maps, memset -- is ~0.9 ms)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]

cl = lambda t: t.to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
for name, cb, cs, big_is_canvas_in in (("conv_123_0 (x canvas 32 ch, gy box 112 ch)", 32, 112, True),
                                      ("deconv_3_0 (gy canvas 32 ch, x box 64 ch)", 32, 64, False),
                                      ("deconv_2_0 (gy canvas 16 ch, x box 32 ch)", 16, 32, False),
                                      ("deconv_1_0 (gy canvas 8 ch, x box 16 ch)", 8, 16, False)):
    big = cl(torch.randn(B, cb, *dims, device=DEV))
    small = cl(torch.randn(B, cs, *box, device=DEV))
    lines = lambda: c.s2_wgrad(big, small, pads, "lines")
    t_lines = timeit(lines)
    g_lines = lines()
    if cb >= 16:
        ours = lambda: c.s2_wgrad(big, small, pads, "classes")
        t = timeit(ours)
        g = ours()
    else:
        t, g = None, g_lines
    P = tuple(q if q >= 2 else q + 2 for q in pads)
    off = tuple((a - b) // 2 for a, b in zip(P, pads))
    nat = tuple((n + 2 * a - 3) // 2 + 1 for n, a in zip(dims, P))
    padding = []
    for ax in (2, 1, 0):
        padding += [off[ax], nat[ax] - off[ax] - box[ax]]
    gfull = torch.nn.functional.pad(small, padding).contiguous(memory_format=torch.channels_last_3d)
    lib = lambda: torch.nn.grad.conv3d_weight(big, (cs, cb, 3, 3, 3), gfull, stride=2, padding=P)
    t_lib = timeit(lib)
    ref = lib().float()                                       # [cs, cb, 3,3,3]
    mine = g.reshape(3, 3, 3, cb, cs).permute(4, 3, 0, 1, 2)
    err = float((mine - ref).abs().max() / ref.abs().max())
    mine_l = g_lines.reshape(3, 3, 3, cb, cs).permute(4, 3, 0, 1, 2)
    err_l = float((mine_l - ref).abs().max() / ref.abs().max())
    print(json.dumps(dict(layer=name, B=B, ms_lines=t_lines, ms_classes=t, ms_cudnn=t_lib, rel_err_lines=err_l, rel_err_classes=err)), flush=True)
