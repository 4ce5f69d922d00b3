timeout 600 python tools/bench_conv_tc.py 4 > gpurun_out/conv_tc_bench_b4.jsonl 2> gpurun_out/conv_tc_bench_b4.err; echo "rc=$?" >> gpurun_out/conv_tc_bench_b4.err
timeout 600 python tools/bench_conv_tc.py 1 > gpurun_out/conv_tc_bench_b1.jsonl 2> gpurun_out/conv_tc_bench_b1.err
