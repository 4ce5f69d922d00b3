timeout 600 python -m pytest tests/test_gpu_graph.py -m gpu -q -x > gpurun_out/pytest_graph.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_graph.log
