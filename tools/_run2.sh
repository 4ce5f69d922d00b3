export NCCL_DEBUG=WARN
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/run_depth_slab.py --reps 5 --graph > gpurun_out/slab2_graph.json 2> gpurun_out/slab2_graph.err; echo "rc=$?" >> gpurun_out/slab2_graph.err
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 tools/run_depth_slab.py --reps 5 > gpurun_out/slab2_eager.json 2> gpurun_out/slab2_eager.err; echo "rc=$?" >> gpurun_out/slab2_eager.err
