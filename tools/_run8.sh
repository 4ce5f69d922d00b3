# gpurun payload, eight B200:  gpurun --gpus 8 --timeout 900 -- 'bash tools/_run8.sh'
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_dp8.json 2> gpurun_out/bench_dp8.err; echo "rc=$?" >> gpurun_out/bench_dp8.err
tail -4 gpurun_out/bench_dp8.err; head -c 400 gpurun_out/bench_dp8.json
