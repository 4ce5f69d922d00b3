python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_for_ncu2.json 2> gpurun_out/bench_for_ncu2.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r01c.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches2.log 2>&1
echo "rc=$?" >> gpurun_out/ncu_launches2.log
