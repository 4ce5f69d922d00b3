#!/usr/bin/env python
"""Achieved HBM bandwidth of the streaming kernels (K3b BatchNorm / K3d box ops) at the regulariser's cfg2 shapes (B = 4).
One JSON line per op: ms, algorithmic GB moved, GB/s, fraction of the measured HBM copy peak."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deep-multiview-depth-estimation_b200"))
import torch  # noqa: E402
from mvs_b200 import ops  # noqa: E402

DEV = "cuda:0"
PEAK = 6533.5
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
B, D, H, W = 4, 192, 128, 160
BOX = (97, 65, 81)
flush = torch.empty(160 * 1024 * 1024, device=DEV)


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def vol(c, dims):
    return torch.randn((B, c) + tuple(dims), device=DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)


def report(name, ms, nbytes):
    print(json.dumps(dict(op=name, ms=ms, GB=nbytes / 1e9, GBps=nbytes / ms / 1e6, frac_hbm=nbytes / ms / 1e6 / PEAK)), flush=True)


for c in (8, 32):
    x = vol(c, (D, H, W)).requires_grad_(True)
    w_, b_ = torch.ones(c, device=DEV, requires_grad=True), torch.zeros(c, device=DEV, requires_grad=True)
    nb = x.numel() * 2
    ops.EVENTS = {}
    y, _, _ = ops.batchnorm_relu_train(x, w_, b_)
    g = torch.randn_like(y)
    y.backward(g)
    torch.cuda.synchronize()
    ops.EVENTS = None
    with torch.no_grad():
        report(f"bn fwd (stats + apply) dense {c}ch canvas", timeit(lambda: ops.batchnorm_relu_train(x.detach(), w_, b_)), 3 * nb)
    def fb():
        xx = x.detach().requires_grad_(True)
        yy, _, _ = ops.batchnorm_relu_train(xx, w_, b_)
        yy.backward(g)
    t_all = timeit(fb)
    t_f = timeit(lambda: ops.batchnorm_relu_train(x.detach().requires_grad_(True), w_, b_))
    report(f"bn bwd (reduce + apply) dense {c}ch canvas", t_all - t_f, 5 * nb)

# crop variant: 32-channel canvas, gradient on the box only
x = vol(32, (D, H, W)).requires_grad_(True)
w_, b_ = torch.ones(32, device=DEV, requires_grad=True), torch.zeros(32, device=DEV, requires_grad=True)
crop = ((48, 145), (32, 97), (40, 121))
y, _, _ = ops.batchnorm_relu_train(x, w_, b_, crop=crop)
g = torch.randn_like(y)
def fbc():
    xx = x.detach().requires_grad_(True)
    yy, _, _ = ops.batchnorm_relu_train(xx, w_, b_, crop=crop)
    yy.backward(g)
t_all = timeit(fbc)
t_f = timeit(lambda: ops.batchnorm_relu_train(x.detach().requires_grad_(True), w_, b_, crop=crop))
nb, nbox = x.numel() * 2, y.numel() * 2
report("bn fwd crop 32ch (stats canvas, apply box)", t_f, nb + 2 * nbox)
report("bn bwd crop 32ch (reduce box, apply canvas)", t_all - t_f, 2 * nbox + nb + nbox + nb)

# box ops, 64 channels
S = vol(64, BOX)
with torch.no_grad():
    report("channel_sums 64ch box", timeit(lambda: ops.channel_sums(S)), S.numel() * 2)
    sc, sh = torch.rand(64, device=DEV) + 0.5, torch.randn(64, device=DEV)
    report("affine_relu_geo fwd 64ch box -> box+2", timeit(lambda: ops.affine_relu_geo(S, sc, sh, (48, 32, 40), (46, 30, 38), (101, 69, 85))),
           S.numel() * 2 + B * 64 * 101 * 69 * 85 * 2)
