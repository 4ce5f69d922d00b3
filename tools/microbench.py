#!/usr/bin/env python
"""Isolated kernel timings (CUDA events on the launching stream, L2 flushed between launches) for the hot-path
kernels: BASELINE.json configs[4] sweep (views 3-7, D 64-512) plus the cfg1/cfg2 shapes.  One JSON line per case.

    python tools/microbench.py [--cases sweep|cfg|one] [--reps 10] [--kernels fwd,bwd,k4]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deep-multiview-depth-estimation_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import torch  # noqa: E402
import mvs_b200  # noqa: E402
from mvs_b200 import _lib, ops  # noqa: E402
import plane_sweep as ps  # noqa: E402  (camera fixtures only)

DEV = "cuda:0"
PEAK = 6533.5
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timeit(fn, reps, flush):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.add_(1.0)                                   # 512 MB write: evicts the 126 MB L2
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def case(B, V, D, h, w, reps, kernels, flush, smooth=False):
    gen = torch.Generator().manual_seed(0)
    K, R, T = ps.synthetic_cameras(B, V, h, w, seed=0)
    d_min, d_int = torch.full((B, 1, 1, 1), 425.0), torch.ones(B, 1, 1, 1)
    feat = torch.randn(B * V, 32, h, w, generator=gen).to(DEV).contiguous(memory_format=torch.channels_last)
    sweep = ops.PlaneSweep(K, R, T, d_min, d_int, B, V, D, 480.0 / D, h, w, DEV)
    nhwc = feat.permute(0, 2, 3, 1)
    st = torch.cuda.current_stream().cuda_stream
    vox = B * D * h * w
    out = []
    for dt, nbytes, name in ((torch.float32, 4, "f32"), (torch.bfloat16, 2, "bf16")):
        cost = torch.empty((B, 32, D, h, w), dtype=dt, device=DEV, memory_format=torch.channels_last_3d)
        if "fwd" in kernels:
            f = lambda: _lib.call("mvsb200_warp_variance_fwd", nhwc.data_ptr(), sweep.view_params.data_ptr(),
                                  sweep.tinv.data_ptr(), cost.data_ptr(), ops._DT[dt], B, V, 32, D, h, w, st)
            med, best = timeit(f, reps, flush)
            alg = 4 * B * V * 32 * h * w + nbytes * vox * 32
            out.append(dict(kernel="warp_variance_fwd", out=name, B=B, V=V, D=D, h=h, w=w, ms=med, ms_best=best,
                            GBps=alg / med / 1e6, frac_hbm=alg / med / 1e6 / PEAK, Gvox_s=vox / med / 1e6))
        if "bwd" in kernels:
            g = torch.randn((B, 32, D, h, w), device=DEV).to(dt).contiguous(memory_format=torch.channels_last_3d)
            gf = torch.empty_like(nhwc.contiguous())
            f = lambda: _lib.call("mvsb200_warp_variance_bwd", nhwc.data_ptr(), sweep.view_params.data_ptr(),
                                  sweep.tinv.data_ptr(), g.data_ptr(), ops._DT[dt], gf.data_ptr(), B, V, 32, D, h, w, st)
            med, best = timeit(f, reps, flush)
            alg = nbytes * vox * 32 + 2 * 4 * B * V * 32 * h * w
            out.append(dict(kernel="warp_variance_bwd", gcost=name, B=B, V=V, D=D, h=h, w=w, ms=med, ms_best=best,
                            GBps=alg / med / 1e6, frac_hbm=alg / med / 1e6 / PEAK))
            del g
        del cost
    if "k4" in kernels:
        logits = torch.randn(B, 1, D, h, w, device=DEV)
        f = lambda: mvs_b200.softmax_depth(logits, sweep.d_batch_dev)
        med, best = timeit(f, reps, flush)
        alg = 2 * 4 * vox + 4 * B * h * w
        out.append(dict(kernel="softmax_depth_fwd", B=B, D=D, h=h, w=w, ms=med, ms_best=best, GBps=alg / med / 1e6,
                        frac_hbm=alg / med / 1e6 / PEAK))
    for o in out:
        print(json.dumps(o), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="cfg")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--kernels", default="fwd,bwd,k4")
    a = ap.parse_args()
    kernels = a.kernels.split(",")
    flush = torch.zeros(128 * 1024 * 1024, device=DEV)
    if a.cases == "one":
        case(1, 3, 192, 128, 160, a.reps, kernels, flush)
    elif a.cases == "cfg":
        case(1, 3, 192, 128, 160, a.reps, kernels, flush)
        case(4, 3, 192, 128, 160, a.reps, kernels, flush)
    else:
        for V in (3, 4, 5, 6, 7):
            for D in (64, 128, 192, 256, 512):
                case(1, V, D, 128, 160, a.reps, kernels, flush)


if __name__ == "__main__":
    main()
