#!/usr/bin/env python
"""Which Python lines issue the big aten::copy_ / fill_ / add kernels of a train step (torch.profiler with stacks)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deep-multiview-depth-estimation_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
from torch.profiler import profile, ProfilerActivity
from mvs_b200.harness import MVSNet, loss_fcn
import plane_sweep as ps
dev = "cuda:0"
B, V, H, W, D = 4, 3, 512, 640, 192
torch.manual_seed(0)
model = MVSNet(D, 480.0 / D, precision="bf16").to(dev).train()
K, R, T = ps.synthetic_cameras(B, V, H // 4, W // 4)
d_min, d_int = torch.full((B, 1, 1, 1), 425.0), torch.ones(B, 1, 1, 1)
img = torch.randn(B * V, 3, H, W, device=dev)
gt = 425 + 480 * torch.rand(B, 1, H // 4, W // 4, device=dev)
def step():
    for p in model.parameters(): p.grad = None
    i, r = model(img, K, R, T, d_min, d_int, B, V)
    loss_fcn(gt, i, r)[0].backward()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], with_stack=True, record_shapes=True) as prof:
    step(); torch.cuda.synchronize()
rows = []
for e in prof.events():
    if e.name in ("aten::copy_", "aten::fill_", "aten::add_", "aten::add", "aten::cat") and e.device_time_total > 60:
        stack = [s for s in (e.stack or []) if "mvs_b200" in s or "harness" in s][:3]
        rows.append((e.device_time_total, e.name, str(e.input_shapes)[:70], " <- ".join(s.split("deep-multiview-depth-estimation_b200/")[-1] for s in stack)))
rows.sort(reverse=True)
for r in rows[:30]:
    print(f"{r[0]:8.0f} us  {r[1]:12s} {r[2]:70s} {r[3]}")
