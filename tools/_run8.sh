# gpurun payload, eight B200:  gpurun --gpus 8 --timeout 900 -- 'bash tools/_run8.sh'   (scaling runs of profiles/r01_scaling.md)
export NCCL_DEBUG=WARN
for n in 8 4; do
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2957$n tools/run_depth_slab.py --reps 10 --graph > gpurun_out/slab${n}_graph.json 2> gpurun_out/slab${n}_graph.err; echo "rc=$?" >> gpurun_out/slab${n}_graph.err
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 8 --workload cfg4 --steps 10 --warmup 3 > gpurun_out/bench_cfg4_slab8g.json 2> gpurun_out/bench_cfg4_slab8g.err; echo "rc=$?" >> gpurun_out/bench_cfg4_slab8g.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_cfg2_dp8.json 2> gpurun_out/bench_cfg2_dp8.err; echo "rc=$?" >> gpurun_out/bench_cfg2_dp8.err
