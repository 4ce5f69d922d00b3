export NCCL_DEBUG=WARN
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/run_depth_slab.py --reps 5 --graph > gpurun_out/slab2_graph.json 2> gpurun_out/slab2_graph.err; echo "rc=$?" >> gpurun_out/slab2_graph.err
