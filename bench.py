#!/usr/bin/env python
"""bench.py -- headline measurement of the plane-sweep hot path (contract: see the task brief / DESIGN.md §5).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload cfg2|cfg1|cfg4] [--no-graph]

Own arm: one "step" = one MVSNet train step (forward + loss + backward + Adam) on one synthetic DTU-shaped batch
(BASELINE.json configs[1]: bf16, batch 4, 3 views, 640x512 input => 160x128x32 features, D = 192), replayed as one CUDA graph
(harness.GraphedTrainStep; --no-graph issues it eagerly).  `value` is depth maps/s with the batch already resident in HBM;
`e2e` is the same step fed from pinned HOST buffers (H2D of the images, ground truth and sweep geometry, D2H read of the loss
inside the timed region).  `roofline` is the dominant own kernel (the stride-1 tcgen05 convolution, tensor bound),
`roofline_k1` / `roofline_k1_fp32` the fused warp+variance kernel (HBM bound), all timed live with CUDA events on the
launching stream.  Under torchrun (--gpus N): cfg2 shards by scene/batch with a flat-bucket NCCL gradient all-reduce (weak
scaling); cfg4 splits ONE sample into depth slabs across the ranks (strong scaling, mvs_b200.depth_slab).

Reference arm (--impl reference): the oracle port of the reference's CPU algorithm (oracle/cpu_path.py) on the
host cores, each step a bounded sample of the same workload (stated in `cpu_baseline.sample`).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "deep-multiview-depth-estimation_b200"))

TC_KERNELS = ("conv3d_s1_tc", "deconv3d_s2_tc", "conv3d_s2_tc", "conv3d_s1_wgrad_tc", "conv3d_s2_wgrad_tc")     # tcgen05 convolution kernels

WORKLOADS = {
    "cfg2": dict(B=4, V=3, H=512, W=640, D=192, train=True,
                 desc="MVSNet train step bf16, batch 4/GPU, 3 views, 640x512, D=192 (BASELINE.json configs[1])"),
    "cfg1": dict(B=1, V=3, H=512, W=640, D=192, train=False,
                 desc="MVSNet forward, batch 1, 3 views, 640x512, D=192 (BASELINE.json configs[0])"),
    "cfg4": dict(B=1, V=5, H=1184, W=1600, D=256, train=False,
                 desc="MVSNet inference, 5 views, 1600x1184, D=256 on ONE GPU (BASELINE.json configs[3] without the depth-slab split)"),
}


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def _tensor_peak():
    """bf16 tensor peak for a kernel timed inside a long step: the sustained figure."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json, sustained cuBLAS bf16)"
    except Exception:
        return 1500.0, "fallback (B200_PROFILING.md)"


def _ncu_traffic(name):
    """DRAM bytes (read + write) of the committed `ncu --set full` capture of a kernel, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", name)) as f:
            m = json.load(f)["metrics"]
        return (float(m["dram__bytes_read.sum"]) + float(m["dram__bytes_write.sum"])) * 1e6
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=lambda: self.rows.extend(self.proc.stdout.readlines()), daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons)}


def run_reference(args, rank, world):
    """CPU arm: oracle port, bounded sample, all host threads.  Only rank 0 works."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import cpu_path
    wl = WORKLOADS[args.workload]
    d_sample = 48
    torch.set_num_threads(os.cpu_count() or 1)
    s = cpu_path.make_sample(V=wl["V"], D=d_sample, h=wl["H"] // 4, w=wl["W"] // 4, d_total=wl["D"])
    for _ in range(args.warmup):
        cpu_path.hot_path_step(s, backward=wl["train"])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_path.hot_path_step(s, backward=wl["train"])
    dt = (time.perf_counter() - t0) / args.steps
    scale = wl["D"] / d_sample
    value = 1.0 / (dt * scale)
    sample = (f"hot path only (warp+variance+regulariser+depth, {'fwd+bwd' if wl['train'] else 'fwd'}), 1 batch item, "
              f"first {d_sample} of {wl['D']} planes at full 160x128x32; time scaled x{scale:g} linearly in D "
              f"(favours the CPU: the reference's torch.cat growth is super-linear in D)")
    line = {"impl": "reference", "metric": "depth maps/sec", "value": value, "unit": "depth maps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"]},
            "cpu_baseline": {"value": value, "unit": "depth maps/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": "depth maps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the mvs_b200 path has no CPU fallback")
    import mvs_b200
    from mvs_b200 import ops
    from mvs_b200.harness import MVSNet, loss_fcn, FlatGradAllReduce
    sys.path.insert(0, os.path.join(ROOT, "oracle"))      # fixtures (DTU cameras) + the cpu_baseline leg only
    import plane_sweep as ps

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if os.environ.get("MVSB200_CUDNN_BENCHMARK", "1") == "1":
        torch.backends.cudnn.benchmark = True              # library convolutions (2D nets, strided 3D backward): autotuned plans
    wl = WORKLOADS[args.workload]
    B, V, H, W, D, train = wl["B"], wl["V"], wl["H"], wl["W"], wl["D"], wl["train"]
    h, w, C = H // 4, W // 4, 32
    d_scale = 480.0 / D
    torch.manual_seed(0)
    model = MVSNet(D, d_scale, precision="bf16").to(dev)
    model.train()                                          # train-mode BN also at test time (test.py:61)
    params = [p for p in model.parameters()]
    opt = torch.optim.Adam(params, lr=0.005, fused=True)   # train.py:160 (fused: one multi-tensor kernel per step)
    reducer = FlatGradAllReduce(params) if world > 1 else None
    if reducer:
        reducer.broadcast_parameters(list(model.buffers()))

    gen = torch.Generator().manual_seed(1000 + rank)
    K, R, T = ps.synthetic_cameras(B, V, h, w, seed=rank)
    d_min, d_int = torch.full((B, 1, 1, 1), 425.0), torch.ones(B, 1, 1, 1)
    img_host = torch.randn(B * V, 3, H, W, generator=gen).pin_memory()
    gt_host = (425.0 + 480.0 * torch.rand(B, 1, h, w, generator=gen))
    gt_host = (gt_host * (torch.rand(B, 1, h, w, generator=gen) > 0.3)).pin_memory()     # 30 % invalid (loss.py:8)
    img_dev, gt_dev = img_host.to(dev), gt_host.to(dev)
    loss_host = torch.zeros(1).pin_memory()

    slab = None
    if args.workload == "cfg4" and world > 1:              # ONE sample across the ranks: depth-slab split
        from mvs_b200.harness import DepthSlabMVSNet
        K, R, T = ps.synthetic_cameras(B, V, h, w, seed=0)  # every rank works on the same scene
        g0 = torch.Generator().manual_seed(1000)
        img_host = torch.randn(B * V, 3, H, W, generator=g0).pin_memory()
        gt_host = (425.0 + 480.0 * torch.rand(B, 1, h, w, generator=g0)).pin_memory()
        img_dev, gt_dev = img_host.to(dev), gt_host.to(dev)
        slab = DepthSlabMVSNet(model, graph=not args.no_graph)

    gstep = ginf = None
    if train and not args.no_graph:
        from mvs_b200.harness import GraphedTrainStep
        gstep = GraphedTrainStep(model, B, V, H, W, dev)
    elif not train and slab is None and not args.no_graph:
        from mvs_b200.harness import GraphedInference
        ginf = GraphedInference(model, B, V, H, W, dev)

    def step(img, gt):
        if gstep is not None:                              # forward + loss + backward as one CUDA graph
            loss = gstep.run(img, gt, K, R, T, d_min, d_int)
            if reducer:
                reducer.reduce()
            opt.step()
            return loss
        if slab is not None:
            initial, refined = slab.forward(img, K, R, T, d_min, d_int, V)
            return loss_fcn(gt, initial, refined)[0]
        if train:
            opt.zero_grad(set_to_none=True)
            initial, refined = model(img, K, R, T, d_min, d_int, B, V)
            loss, _, _ = loss_fcn(gt, initial, refined)
            loss.backward()
            if reducer:
                reducer.reduce()
            opt.step()
            return loss.detach()
        with torch.no_grad():
            if ginf is not None:                           # inference forward replayed as one CUDA graph
                initial, refined = ginf.run(img, K, R, T, d_min, d_int)
            else:
                initial, refined = model(img, K, R, T, d_min, d_int, B, V)
            return loss_fcn(gt, initial, refined)[0]

    def step_resident():
        return step(img_dev, gt_dev)

    # host-fed step: every step's inputs cross PCIe from pinned host memory inside the timed region, as a prefetching loader
    # would deliver them -- double-buffered device staging, the copy of step i+1 on a side stream while step i computes
    copy_stream = torch.cuda.Stream()
    stage = [(torch.empty_like(img_dev), torch.empty_like(gt_dev)) for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    feed = {"slot": 0, "primed": False}

    def prefetch(slot_):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot_])        # the step that read this staging pair has finished with it
            stage[slot_][0].copy_(img_host, non_blocking=True)
            stage[slot_][1].copy_(gt_host, non_blocking=True)
            ready[slot_].record(copy_stream)

    def step_e2e():
        cur = feed["slot"]
        if not feed["primed"]:
            consumed[0].record(); consumed[1].record()
            prefetch(cur)
            feed["primed"] = True
        torch.cuda.current_stream().wait_event(ready[cur])
        prefetch(cur ^ 1)                                  # next step's H2D overlaps this step's kernels
        loss = step(stage[cur][0], stage[cur][1])
        consumed[cur].record()
        loss_host.copy_(loss.reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()          # the user reads the loss every step
        feed["slot"] = cur ^ 1
        return loss_host

    def timed(fn, n):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    graph_note = None
    if gstep is not None:
        try:
            step_resident()
            torch.cuda.synchronize()
        except Exception as e:                             # capture refused: say so and measure the eager step
            graph_note = f"capture failed, eager step measured: {type(e).__name__}: {str(e)[:200]}"
            sys.stderr.write("bench.py: " + graph_note + "\n")
            gstep = None
            opt.zero_grad(set_to_none=True)
    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if gstep is None and ginf is None:
        ops.EVENTS = {}
    n0 = mvs_b200.launch_count()
    ms = timed(step_resident, args.steps)
    launches = ((gstep.launches * args.steps) if gstep is not None else
                (ginf.launches * args.steps) if ginf is not None else mvs_b200.launch_count() - n0)
    events, ops.EVENTS = ops.EVENTS, None
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    if gstep is not None or ginf is not None:
        # per-kernel durations cannot be bracketed inside a graph replay: the same step, issued eagerly right after the
        # timed region with every libmvs_b200.so launch between two CUDA events on its launching stream
        graphed, gstep, graphed_inf, ginf = gstep, None, ginf, None
        for p_ in params:
            p_.grad = None
        step_resident()
        ops.EVENTS = {}
        for _ in range(args.steps):
            step_resident()
        torch.cuda.synchronize()
        events, ops.EVENTS = ops.EVENTS, None
        gstep, ginf = graphed, graphed_inf                 # (reporting only from here on)

    maps = B * (1 if slab is not None else world) * args.steps
    value, e2e_value = maps / (ms * 1e-3), maps / (ms_e2e * 1e-3)

    # per-kernel live timings -> roofline of the fused warp+variance kernel
    peak, peak_src = _peaks()
    tpeak, tpeak_src = _tensor_peak()
    vox = B * D * h * w
    if slab is not None:                                   # K1 of this rank: its own planes + halo
        k0, k1_ = slab.reg.plan(D).cost_planes(rank)
        vox = B * (k1_ - k0) * h * w
    alg = {"warp_variance_fwd": 4 * B * V * C * h * w + 2 * vox * C,            # fp32 features in, bf16 volume out
           "warp_variance_bwd": 2 * vox * C + 2 * 4 * B * V * C * h * w,        # bf16 gcost + features in, gfeat out
           "softmax_ranks_fwd": 2 * 4 * vox + 4 * 5 * B * h * w}
    kern = {}
    for name, evs in events.items():
        tot = sum(a.elapsed_time(b) for a, b, _ in evs)
        kern[name] = {"ms": tot / len(evs), "launches": len(evs), "ms_per_step": tot / args.steps}
        if name in alg:
            t = tot / len(evs)
            kern[name].update({"alg_bytes": alg[name], "GBps": alg[name] / (t * 1e-3) / 1e9,
                               "frac_hbm": alg[name] / (t * 1e-3) / 1e9 / peak})
        work = [w for _, _, w in evs if w is not None]
        if work and name in TC_KERNELS:
            kern[name].update({"alg_flops_per_step": sum(work) / args.steps, "TFLOPs": sum(work) / (tot * 1e-3) / 1e12})
    k1 = kern.get("warp_variance_fwd", {})
    roofline_k1 = {"kernel": "warp_variance_fwd2_kernel<V=%d, bf16 volume> (as used by the bf16 regulariser)" % V, "bound": "hbm", "achieved": k1.get("GBps"),
                   "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": k1.get("frac_hbm"),
                   "traffic": None, "alg_bytes_per_launch": alg["warp_variance_fwd"], "ms_per_launch": k1.get("ms"),
                   "voxels_per_s": vox / (k1["ms"] * 1e-3) if k1 else None,
                   "note": "ncu (profiles/r01_k1_*): DRAM traffic = algorithmic bytes (fp32-volume capture 7.7 MB read + 446 MB "
                           "written of 511 MB); fp32-volume variant reaches 44 % of peak (tools/microbench.py)"}
    # the same kernel with the fp32 volume the reference produces (SURVEY §8d quotes the HBM roofline on these bytes:
    # 4*V*C*h*w in + 4*C*D*h*w out per sample), timed live on this batch's geometry: L2 flushed between launches, CUDA events
    roofline_k1_fp32 = None
    if slab is None:
        with torch.no_grad():
            sweep = ops.PlaneSweep(K, R, T, d_min, d_int, B, V, D, d_scale, h, w, dev)
            feat = torch.randn(B * V, h, w, C, device=dev).permute(0, 3, 1, 2)
            flush = torch.empty(160 * 1024 * 1024, dtype=torch.float32, device=dev)       # 640 MB > L2
            ts = []
            for i in range(3 + 5):
                flush.zero_()
                a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                cv32 = ops.warp_variance(feat, sweep, torch.float32)
                b_.record()
                torch.cuda.synchronize()
                if i >= 3:
                    ts.append(a.elapsed_time(b_))
                del cv32
            del flush
        t32 = sorted(ts)[len(ts) // 2]
        bytes32 = 4 * B * V * C * h * w + 4 * B * D * h * w * C
        roofline_k1_fp32 = {"kernel": "warp_variance_fwd2_kernel<V=%d, fp32 volume>" % V, "bound": "hbm",
                            "achieved": bytes32 / (t32 * 1e-3) / 1e9, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                            "frac": bytes32 / (t32 * 1e-3) / 1e9 / peak, "alg_bytes_per_launch": bytes32, "ms_per_launch": t32,
                            "voxels_per_s": B * D * h * w / (t32 * 1e-3),
                            "traffic": _ncu_traffic("r01_k1_fwd2_ncu.json"),
                            "note": "isolated launches on the step's shapes (B=%d), median of 5, L2 flushed" % B}
    # dominant own kernel by time in the step: the tcgen05 convolution (all its launches of the timed steps together)
    k3 = kern.get("conv3d_s1_tc", {})
    tc_ms = sum(kern[n]["ms_per_step"] for n in TC_KERNELS if n in kern and "alg_flops_per_step" in kern[n])
    tc_fl = sum(kern[n]["alg_flops_per_step"] for n in TC_KERNELS if n in kern and "alg_flops_per_step" in kern[n])
    roofline = {"kernel": "conv3d_s1_tc_kernel<CIN,NOUT> (tcgen05/TMEM/TMA implicit-GEMM conv3d; %d calls/step: forward and data "
                          "gradient of conv_0_0 and conv_{1,2,3}_1)" % (k3.get("launches", 0) // max(args.steps, 1)),
                "all_tcgen05_convs": {"kernels": [n for n in TC_KERNELS if n in kern], "ms_per_step": tc_ms,
                                      "alg_flops_per_step": tc_fl, "TFLOPs": tc_fl / (tc_ms * 1e-3) / 1e12 if tc_ms else None,
                                      "frac": tc_fl / (tc_ms * 1e-3) / 1e12 / tpeak if tc_ms else None},
                "bound": "tensor", "achieved": k3.get("TFLOPs"), "peak": tpeak, "peak_source": tpeak_src, "unit": "TFLOP/s",
                "frac": (k3["TFLOPs"] / tpeak) if k3.get("TFLOPs") else None,
                "traffic": _ncu_traffic("r01_k3_kdn_ncu.json"),
                "traffic_case": "ncu --set full capture of the 32->32 dense canvas launch (252 MB in + 252 MB out algorithmic)",
                "alg_flops_per_step": k3.get("alg_flops_per_step"), "ms_per_step": k3.get("ms_per_step"),
                "note": "conv3d_s1_kdn_kernel (depth tap folded into the MMA N extent); ncu of the 32->32 dense launch: MAC array busy 67 % of "
                        "cycles, 1154 TFLOP/s in isolation = 70 % of the burst bf16 peak: profiles/r01_k3_notes.md"}

    line = None
    if rank == 0:
        line = {"metric": "depth maps/sec", "value": value, "unit": "depth maps/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong" if slab is not None else "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": wl["desc"], "per_gpu_batch": B, "views": V, "D": D, "features": [C, h, w],
                           "l2": "inputs_exceed_l2 (cost volume %.0f MB per step > 126 MB L2)" % (2 * vox * C / 1e6),
                           "parallelism": (f"depth-slab x{world} (one sample; K1 on own planes + halo, halo/box exchanges and BatchNorm-sum "
                                           f"all-reduces over NCCL, logits re-sharded to rows for K4; hot path "
                                           f"{'replayed as one CUDA graph per rank' if not args.no_graph else 'issued eagerly'})"
                                           if slab is not None else
                                           f"dp{world} (scene/batch sharding, flat-bucket NCCL grad all-reduce)" if world > 1 else "single GPU"),
                           "regulariser_convs": model.cost_volume_reg.conv_backend,
                           "cudnn_benchmark": bool(torch.backends.cudnn.benchmark),
                           "cuda_graph": (f"forward+loss+backward replayed as one CUDA graph ({gstep.launches} libmvs_b200.so "
                                          f"launches per replay); per-kernel timings from {args.steps} eager steps run right "
                                          f"after the timed region") if gstep is not None else
                                         (f"inference forward replayed as one CUDA graph ({ginf.launches} libmvs_b200.so launches per "
                                          f"replay)" if ginf is not None else (graph_note or "off"))},
                "e2e": {"value": e2e_value, "unit": "depth maps/s", "ms_per_step": ms_e2e / args.steps,
                        "h2d_bytes_per_step": img_host.numel() * 4 + gt_host.numel() * 4, "d2h_bytes_per_step": 4,
                        "input_pipeline": "pinned host buffers, double-buffered device staging: the H2D copy of step i+1 runs on a "
                                          "side stream while step i computes; the loss is read back (synchronising) every step"},
                "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "roofline_k1": roofline_k1, "roofline_k1_fp32": roofline_k1_fp32, "kernels": kern,
                "cost_volume_voxels_per_s": roofline_k1["voxels_per_s"]}
    return line


def cpu_baseline(workload):
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_path
    wl = WORKLOADS[workload]
    # cfg1/cfg2: the whole D = 192 sweep of one batch item (~13 s on 16 cores, nothing extrapolated); cfg4: 24 of 256 planes
    # (the reference materialises 19 GB of warped volumes at cfg4)
    d_sample = wl["D"] if wl["H"] * wl["W"] <= 512 * 640 else 24
    torch.set_num_threads(os.cpu_count() or 1)
    s = cpu_path.make_sample(V=wl["V"], D=d_sample, h=wl["H"] // 4, w=wl["W"] // 4, d_total=wl["D"])
    dt, _ = cpu_path.hot_path_step(s, backward=wl["train"])
    scale = wl["D"] / d_sample
    return {"value": 1.0 / (dt * scale), "unit": "depth maps/s", "cores": torch.get_num_threads(), "kind": "port",
            "seconds_sample": dt,
            "sample": (f"oracle port (oracle/cpu_path.py), hot path {'fwd+bwd' if wl['train'] else 'fwd'}, 1 batch item, first "
                       f"{d_sample} of {wl['D']} planes at 160x128x32, one un-warmed pass; time scaled x{scale:g} linearly in D")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="issue the train step eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    line = run_b200(args, rank, world, local_rank)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.workload)
        print(json.dumps(line), flush=True)
    if world > 1:
        if args.workload == "cfg4" and not args.no_graph:
            # a replayed graph that holds NCCL work: drain and leave without the collective teardown (it does not return)
            torch.cuda.synchronize()
            sys.stdout.flush(); sys.stderr.flush()
            os._exit(0)
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
