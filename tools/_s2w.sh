timeout 600 python -m pytest tests/test_gpu_conv_tc.py -q -x -k "stride2_conv or transposed_conv" > gpurun_out/pytest_s2w.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_s2w.log
tail -25 gpurun_out/pytest_s2w.log
