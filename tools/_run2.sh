# gpurun payload, two B200:  gpurun --gpus 2 --timeout 900 -- 'bash tools/_run2.sh'
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_dp2.json 2> gpurun_out/bench_dp2.err; echo "rc=$?" >> gpurun_out/bench_dp2.err
tail -5 gpurun_out/bench_dp2.err; head -c 300 gpurun_out/bench_dp2.json
