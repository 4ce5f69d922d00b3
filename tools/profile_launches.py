#!/usr/bin/env python
"""Every GPU kernel of one train step: launches and device time per kernel name (torch.profiler), own kernels marked."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deep-multiview-depth-estimation_b200"))
import torch
from torch.profiler import profile, ProfilerActivity
from mvs_b200.harness import MVSNet, loss_fcn, synthetic_cameras
dev = "cuda:0"
torch.backends.cudnn.benchmark = True
B, V, H, W, D = 4, 3, 512, 640, 192
torch.manual_seed(0)
model = MVSNet(D, 480.0 / D, precision="bf16").to(dev).train()
opt = torch.optim.Adam(model.parameters(), lr=0.005, fused=True)
K, R, T = synthetic_cameras(B, V, H // 4, W // 4)
d_min, d_int = torch.full((B, 1, 1, 1), 425.0), torch.ones(B, 1, 1, 1)
img = torch.randn(B * V, 3, H, W, device=dev)
gt = 425 + 480 * torch.rand(B, 1, H // 4, W // 4, device=dev)
def step():
    opt.zero_grad(set_to_none=True)
    i, r = model(img, K, R, T, d_min, d_int, B, V)
    loss_fcn(gt, i, r)[0].backward()
    opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
acc = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        a = acc[e.name[:90]]; a[0] += 1; a[1] += e.device_time
n = sum(v[0] for v in acc.values()); t = sum(v[1] for v in acc.values())
print(f"kernels {n}  device time {t/1000:.2f} ms")
small = sum(v[0] for v in acc.values() if v[1] / v[0] < 10); tsmall = sum(v[1] for v in acc.values() if v[1] / v[0] < 10)
print(f"kernels under 10 us on average: {small} launches, {tsmall/1000:.2f} ms")
for k, v in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:9.1f} us {v[0]:5d} x  {k}")
