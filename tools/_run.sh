timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python tools/microbench.py --cases sweep --kernels fwd --reps 5 > gpurun_out/micro_cfg5_fwd2.jsonl 2>&1
python tools/microbench.py --cases one --reps 2 --kernels fwd > gpurun_out/k1v2_plain.jsonl 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:warp_variance_fwd2 -s 2 -c 1 -o gpurun_out/k1_fwd2_r1 -f python tools/microbench.py --cases one --reps 2 --kernels fwd > gpurun_out/ncu_k1_fwd2.log 2>&1
echo "rc=$?" >> gpurun_out/ncu_k1_fwd2.log
