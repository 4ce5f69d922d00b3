timeout 600 python -m pytest tests/test_gpu_conv_tc.py -q -x -k "stride2_conv or transposed_conv" > gpurun_out/pytest_s2w.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_s2w.log
tail -4 gpurun_out/pytest_s2w.log
for d in 4 5 6; do echo "dbg=$d"; MVSB200_S2WG_DBG=$d timeout 300 python tools/bench_s2_wgrad.py 2>&1 | sort -u | cut -c1-150; done
