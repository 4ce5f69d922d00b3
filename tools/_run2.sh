export NCCL_DEBUG=WARN
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus 2 --workload cfg4 --steps 5 --warmup 3 > gpurun_out/bench_cfg4_slab2g.json 2> gpurun_out/bench_cfg4_slab2g.err; echo "rc=$?" >> gpurun_out/bench_cfg4_slab2g.err
