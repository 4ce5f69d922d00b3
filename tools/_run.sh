python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python tools/microbench.py --cases cfg --kernels fwd > gpurun_out/micro_cfg5.jsonl 2> gpurun_out/micro.err
ncu --set full --clock-control none --import-source on -k regex:warp_variance_fwd -s 3 -c 1 -o gpurun_out/k1_r1e python tools/microbench.py --cases one --reps 2 --kernels fwd > gpurun_out/ncu_k1e.log 2>&1
