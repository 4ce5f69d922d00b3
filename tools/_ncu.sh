MVSB200_K2=3 python tools/microbench.py --cases one --reps 2 --kernels bwd > gpurun_out/plain_k2.log 2>&1 && \
MVSB200_K2=3 ncu --set full --clock-control none --import-source on -k regex:warp_variance_bwd3 -s 3 -c 1 -f -o gpurun_out/k2_bwd3b_r2 python tools/microbench.py --cases one --reps 2 --kernels bwd > gpurun_out/ncu_k2_bwd3.log 2>&1
echo "rc=$?" >> gpurun_out/ncu_k2_bwd3.log
tail -n 3 gpurun_out/ncu_k2_bwd3.log
