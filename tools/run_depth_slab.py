#!/usr/bin/env python
"""NCCL run of the depth-slab path (one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/run_depth_slab.py [--D 256 --H 1184 --W 1600 --V 5] [--reps 5]

Every rank runs its slab; rank 0 additionally runs the single-GPU module on the whole volume and reports the differences
(logits, probability volume, depth map) and the timings as one JSON line."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deep-multiview-depth-estimation_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))          # synthetic DTU cameras only


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--D", type=int, default=256)
    ap.add_argument("--H", type=int, default=1184)
    ap.add_argument("--W", type=int, default=1600)
    ap.add_argument("--V", type=int, default=5)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--profile", action="store_true", help="torch.profiler table of one eager pass on rank 0 (to stderr)")
    ap.add_argument("--graph", action="store_true", help="replay the rank's whole forward (kernels + NCCL exchanges) as one CUDA graph")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import mvs_b200
    from mvs_b200 import ops
    from mvs_b200.depth_slab import DepthSlabCostVolumeReg, slab_cost_fn, SlabPlan
    import plane_sweep as ps

    D, V, h, w = args.D, args.V, args.H // 4, args.W // 4
    torch.manual_seed(0)
    K, R, T = ps.synthetic_cameras(1, V, h, w)
    d_min, d_int = torch.full((1, 1, 1, 1), 425.0), torch.ones(1, 1, 1, 1)
    feat = torch.randn(V, 32, h, w, device=dev)
    reg = mvs_b200.CostVolumeReg(device=dev, precision="bf16").train()
    for t in list(reg.parameters()) + list(reg.buffers()):
        dist.broadcast(t.data, 0)
    dist.broadcast(feat, 0)
    sweep = mvs_b200.PlaneSweep(K, R, T, d_min, d_int, 1, V, D, 480.0 / D, h, w, dev)
    sharded = DepthSlabCostVolumeReg(reg)
    cost_fn = slab_cost_fn(feat, sweep, torch.bfloat16)

    def run():
        return sharded.forward(cost_fn, sweep.d_batch_dev, 1, D, h, w)

    for _ in range(2):
        depth, prob_rows, rows = run()
    dist.barrier(); torch.cuda.synchronize()
    if args.profile:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            run(); torch.cuda.synchronize()
        if rank == 0:
            sys.stderr.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70) + "\n")
        dist.barrier()
    if args.graph:
        eager = run
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(); dist.barrier()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            captured = eager()

        def run():
            g.replay()
            return captured
        run(); torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        depth, prob_rows, rows = run()
    e1.record(); torch.cuda.synchronize(); dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / args.reps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    logits_all = [torch.empty((1, 1, b - a, h, w), device=dev) for a, b in SlabPlan(D, world).canvas]
    dist.all_gather(logits_all, sharded.last_logits.contiguous())      # equal slabs when D % (2 world) == 0
    out = {"world": world, "D": D, "V": V, "h": h, "w": w, "ms_sharded": float(ms.item()), "cuda_graph": bool(args.graph)}
    if rank == 0:
        import copy
        reg1 = copy.deepcopy(reg)
        def single():
            cost = ops.warp_variance(feat, sweep, torch.bfloat16)
            lg = reg1.logits(cost, mvs_b200.conv3d.get(reg1.conv_backend))
            return (lg,) + tuple(ops.softmax_depth(lg, sweep.d_batch_dev, 5))

        with torch.no_grad():
            for _ in range(2):
                ref_logits, ref_prob, ref_depth = single()
            torch.cuda.synchronize()
            run1 = single
            if args.graph:                              # the unsharded pass replayed as a graph too: like for like
                g1 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g1):
                    held = single()
                ref_logits, ref_prob, ref_depth = held
                run1 = g1.replay
                run1(); torch.cuda.synchronize()
            e0.record()
            for _ in range(args.reps):
                run1()
            e1.record(); torch.cuda.synchronize()
        logits = torch.cat(logits_all, 2)
        step = 480.0 / D
        out.update({"ms_single_gpu": e0.elapsed_time(e1) / args.reps,
                    "logits_rel_err": float((logits - ref_logits).abs().max() / ref_logits.abs().max()),
                    "depth_within_5pct_step": float(((depth - ref_depth).abs() < 0.05 * step).float().mean()),
                    "depth_median_abs_err_steps": float((depth - ref_depth).abs().median() / step)})
        print(json.dumps(out), flush=True)
    if args.graph:
        # a live graph that holds NCCL work keeps the communicator busy in its teardown: drop it, drain the device, and leave
        # without the collective destroy (observed: destroy_process_group() after a captured forward does not return)
        del run, captured, g
        torch.cuda.synchronize()
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
