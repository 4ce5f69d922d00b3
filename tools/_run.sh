timeout 600 python -m pytest tests/test_gpu_graph.py -m gpu -q -x > gpurun_out/pytest_graph.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_graph.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench10.json 2> gpurun_out/bench10.err; echo "rc=$?" >> gpurun_out/bench10.err
