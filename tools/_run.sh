timeout 600 python -m pytest tests/test_gpu_batchnorm.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_bn.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_bn.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench23.json 2> gpurun_out/bench23.err; echo "rc=$?" >> gpurun_out/bench23.err
