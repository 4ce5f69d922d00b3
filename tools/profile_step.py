#!/usr/bin/env python
"""Where does a step go?  torch.profiler table of one MVSNet step (own kernels + library kernels)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deep-multiview-depth-estimation_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402
from torch.profiler import profile, ProfilerActivity  # noqa: E402
from mvs_b200.harness import MVSNet, loss_fcn  # noqa: E402
import plane_sweep as ps  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=4)
ap.add_argument("--D", type=int, default=192)
ap.add_argument("--fwd-only", action="store_true")
ap.add_argument("--precision", default="bf16")
ap.add_argument("--rows", type=int, default=45)
ap.add_argument("--shapes", action="store_true")
a = ap.parse_args()
dev = "cuda:0"
torch.backends.cudnn.benchmark = True
B, V, H, W, D = a.B, 3, 512, 640, a.D
torch.manual_seed(0)
model = MVSNet(D, 480.0 / D, precision=a.precision).to(dev).train()
opt = torch.optim.Adam(model.parameters(), lr=0.005)
K, R, T = ps.synthetic_cameras(B, V, H // 4, W // 4)
d_min, d_int = torch.full((B, 1, 1, 1), 425.0), torch.ones(B, 1, 1, 1)
img = torch.randn(B * V, 3, H, W, device=dev)
gt = 425 + 480 * torch.rand(B, 1, H // 4, W // 4, device=dev)


def step():
    if a.fwd_only:
        with torch.no_grad():
            return model(img, K, R, T, d_min, d_int, B, V)
    opt.zero_grad(set_to_none=True)
    i, r = model(img, K, R, T, d_min, d_int, B, V)
    loss_fcn(gt, i, r)[0].backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=a.shapes) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages(group_by_input_shape=a.shapes).table(sort_by="cuda_time_total", row_limit=a.rows, max_name_column_width=60, max_shapes_column_width=110))
print("peak memory GB:", torch.cuda.max_memory_allocated() / 1e9)
