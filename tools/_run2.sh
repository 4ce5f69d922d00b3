# gpurun payload, two B200:  gpurun --gpus 2 --timeout 700 -- 'bash tools/_run2.sh'   (depth-slab split over NCCL, parity vs 1 GPU)
export NCCL_DEBUG=WARN
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/run_depth_slab.py --reps 10 --graph > gpurun_out/slab2_graph.json 2> gpurun_out/slab2_graph.err; echo "rc=$?" >> gpurun_out/slab2_graph.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --workload cfg4 --steps 5 --warmup 3 > gpurun_out/bench_cfg4_slab2g.json 2> gpurun_out/bench_cfg4_slab2g.err; echo "rc=$?" >> gpurun_out/bench_cfg4_slab2g.err
