timeout 600 python -m pytest tests/test_dropin_boundary.py -m gpu -q -x > gpurun_out/pytest_dropin.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_dropin.log
