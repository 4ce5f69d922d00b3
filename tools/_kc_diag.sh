timeout 600 python -m pytest tests/test_gpu_conv_tc.py -q -x -k "stride2_conv" 2>&1 | grep -B5 -A12 "^E " | head -60
