timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "hanging" > gpurun_out/pytest_edges.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_edges.log
