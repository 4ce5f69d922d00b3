export NCCL_DEBUG=WARN
for n in 8 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n tools/run_depth_slab.py --reps 5 > gpurun_out/slab${n}_cfg4.json 2> gpurun_out/slab${n}_cfg4.err; echo "rc=$?" >> gpurun_out/slab${n}_cfg4.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --workload cfg4 --steps 5 --warmup 3 > gpurun_out/bench_cfg4_slab8.json 2> gpurun_out/bench_cfg4_slab8.err; echo "rc=$?" >> gpurun_out/bench_cfg4_slab8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_cfg2_dp8.json 2> gpurun_out/bench_cfg2_dp8.err; echo "rc=$?" >> gpurun_out/bench_cfg2_dp8.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/bench_cfg2_dp4.json 2> gpurun_out/bench_cfg2_dp4.err; echo "rc=$?" >> gpurun_out/bench_cfg2_dp4.err
