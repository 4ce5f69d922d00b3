timeout 600 python -m pytest tests/test_gpu_conv_tc.py -m gpu -q -x -k "conv_out" > gpurun_out/pytest_co.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_co.log
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench14.json 2> gpurun_out/bench14.err; echo "rc=$?" >> gpurun_out/bench14.err
