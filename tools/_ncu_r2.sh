# round-2 evidence: ncu --set full of the two new tcgen05 kernels (tools/bench_s2_wgrad.py, tools/bench_s2_dgrad.py), then the
# launch list of the bench step.  The list is taken on the eager form of the step (--no-graph): a replayed CUDA graph is one
# opaque launch to ncu's per-kernel list, and with the graph form ncu 2025.2 stopped with "An error was reported by the driver"
# at the first launch it met inside the capture.  the whole run is listed and tools/launch_summary.py
# keeps its last two steady-state steps (a step starts at image_to_rows8_kernel, the encoder's first launch).
if [ "$1" != "launches" ]; then
python tools/bench_s2_wgrad.py > gpurun_out/plain_s2w.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv3d_s2_wgrad_lines -s 2 -c 1 -f -o gpurun_out/r02_s2_wgrad_lines python tools/bench_s2_wgrad.py > gpurun_out/ncu_s2w.log 2>&1
echo "rc=$?" >> gpurun_out/ncu_s2w.log; tail -n 2 gpurun_out/ncu_s2w.log
python tools/bench_s2_dgrad.py > gpurun_out/plain_kc.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:deconv3d_s2_kc -s 2 -c 1 -f -o gpurun_out/r02_deconv_kc python tools/bench_s2_dgrad.py > gpurun_out/ncu_kc.log 2>&1
echo "rc=$?" >> gpurun_out/ncu_kc.log; tail -n 2 gpurun_out/ncu_kc.log
fi
python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-extras > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/r02_bench_launches_all.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-extras > gpurun_out/ncu_bench.log 2>&1
echo "rc=$?" >> gpurun_out/ncu_bench.log; tail -n 2 gpurun_out/ncu_bench.log; wc -l gpurun_out/r02_bench_launches_all.csv
python tools/launch_summary.py gpurun_out/r02_bench_launches_all.csv gpurun_out/r02_bench_launches.csv 2 > gpurun_out/r02_bench_launches_summary.md; head -12 gpurun_out/r02_bench_launches_summary.md
