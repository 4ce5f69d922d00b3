python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_plain_for_ncu.json 2> gpurun_out/bench_plain_for_ncu.err && \
ncu --set full --clock-control none --import-source on -k regex:deconv3d_s2_tc -s 9 -c 1 -o gpurun_out/k3_deconv_r1 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu_deconv.log 2>&1
echo "rc=$?" >> gpurun_out/ncu_deconv.log
