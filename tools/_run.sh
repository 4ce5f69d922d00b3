timeout 300 python bench.py --workload cfg1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg1_g.json 2> gpurun_out/bench_cfg1_g.err; echo "rc=$?" >> gpurun_out/bench_cfg1_g.err
timeout 300 python bench.py --workload cfg1 --steps 10 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_cfg1_e.json 2> gpurun_out/bench_cfg1_e.err
timeout 300 python bench.py --workload cfg4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4_g.json 2> gpurun_out/bench_cfg4_g.err; echo "rc=$?" >> gpurun_out/bench_cfg4_g.err
