timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench24.json 2> gpurun_out/bench24.err; echo "rc=$?" >> gpurun_out/bench24.err
