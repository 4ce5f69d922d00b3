for i in 1 2 3 4 5 6 7 8; do timeout 600 python -m pytest tests/test_gpu_depth_slab.py -m gpu -q 2>&1 | grep -E "^FAILED|passed|failed" >> gpurun_out/pytest_flaky3.log; done
