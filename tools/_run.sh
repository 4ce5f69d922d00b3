python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python tools/microbench.py --cases cfg --kernels fwd,bwd > gpurun_out/micro_cfg6.jsonl 2> gpurun_out/micro.err
