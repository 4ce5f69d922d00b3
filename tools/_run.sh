python tools/microbench.py --cases sweep --kernels fwd,bwd --reps 5 > gpurun_out/sweep_cfg5.jsonl 2> gpurun_out/sweep.err
python bench.py --workload cfg1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err
python tools/microbench.py --cases cfg --kernels fwd,bwd,k4 > gpurun_out/micro_cfg7.jsonl 2>> gpurun_out/sweep.err
