timeout 600 python -m pytest tests/test_gpu_depth_slab.py -m gpu -q -x > gpurun_out/pytest_slab.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_slab.log
