C=deep-multiview-depth-estimation_b200/csrc
echo "8-lane minb3" > gpurun_out/micro_k1b.jsonl; MVSB200_K1_LANES=8 python tools/microbench.py --cases one --kernels fwd >> gpurun_out/micro_k1b.jsonl 2>> gpurun_out/micro.err
echo "8-lane minb2 (110 regs)" >> gpurun_out/micro_k1b.jsonl; MVSB200_LIB=$PWD/$C/libmvs_b200_minb2.so MVSB200_K1_LANES=8 python tools/microbench.py --cases one --kernels fwd >> gpurun_out/micro_k1b.jsonl 2>> gpurun_out/micro.err
