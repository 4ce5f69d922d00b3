export NCCL_DEBUG=WARN
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 8 --workload cfg4 --steps 10 --warmup 3 > gpurun_out/bench_cfg4_slab8g.json 2> gpurun_out/bench_cfg4_slab8g.err; echo "rc=$?" >> gpurun_out/bench_cfg4_slab8g.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_cfg2_dp8.json 2> gpurun_out/bench_cfg2_dp8.err; echo "rc=$?" >> gpurun_out/bench_cfg2_dp8.err
