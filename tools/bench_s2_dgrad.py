#!/usr/bin/env python
"""Data gradient of the stacked stride-2 branches (112 -> 32, B = 4, cfg2 shapes): deconv3d_s2_kc_kernel next to cuDNN."""
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
        for _ in range(40):
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


cl = lambda t: t.to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
gy = cl(torch.randn(B, 112, *box, device=DEV))
w = (torch.randn(112, 32, 3, 3, 3, device=DEV) / 30).to(torch.bfloat16)
acc = cl(torch.randn(B, 32, *dims, device=DEV))
for name, fn in (("write", lambda: c.conv_transpose3d_s2_kc(gy, (16, 32, 64), w, pads, dims)),
                 ("accumulate", lambda: c.conv_transpose3d_s2_kc(gy, (16, 32, 64), w, pads, dims, out=acc, accumulate=True))):
    print(json.dumps(dict(op="deconv3d_s2_kc 112->32 " + name, B=B, ms=timeit(fn))), flush=True)
P = tuple(q if q >= 2 else q + 2 for q in pads)
off = tuple((a - b) // 2 for a, b in zip(P, pads))
nat = tuple((n + 2 * a - 3) // 2 + 1 for n, a in zip(dims, P))
padding = []
for ax in (2, 1, 0):
    padding += [off[ax], nat[ax] - off[ax] - box[ax]]
gfull = torch.nn.functional.pad(gy, padding).contiguous(memory_format=torch.channels_last_3d)
x = cl(torch.randn(B, 32, *dims, device=DEV))
torch.backends.cudnn.benchmark = True
lib = lambda: torch.ops.aten.convolution_backward(gfull, x, w, None, [2, 2, 2], list(P), [1, 1, 1], False, [0, 0, 0], 1, [True, False, False])[0]
print(json.dumps(dict(op="cuDNN dgrad 112->32", B=B, ms=timeit(lib))), flush=True)
mine = c.conv_transpose3d_s2_kc(gy, (16, 32, 64), w, pads, dims).float()
ref = lib().float()
print(json.dumps(dict(rel_err=float((mine - ref).abs().max() / ref.abs().max()))))
