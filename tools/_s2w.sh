timeout 600 python -m pytest tests/test_gpu_conv_tc.py -q -x -k "stride2_conv or transposed_conv" > gpurun_out/pytest_s2w.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_s2w.log
tail -3 gpurun_out/pytest_s2w.log
timeout 300 python tools/bench_s2_wgrad.py 2>&1 | cut -c1-150
echo nkd1; MVSB200_S2WG_NKD=1 timeout 300 python tools/bench_s2_wgrad.py 2>&1 | cut -c1-150
