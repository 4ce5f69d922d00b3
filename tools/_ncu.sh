# gpurun payload: ncu --set full of the forward kernels (both forms) on the cfg1 case
python tools/microbench.py --cases one --reps 2 --kernels fwd > gpurun_out/plain_k1.log 2>&1 && \
MVSB200_K1=3 ncu --set full --clock-control none --import-source on -k regex:warp_variance_fwd3 -s 3 -c 1 -f -o gpurun_out/k1_fwd3_r2 python tools/microbench.py --cases one --reps 2 --kernels fwd > gpurun_out/ncu_k1_fwd3.log 2>&1
echo "rc=$?" >> gpurun_out/ncu_k1_fwd3.log
MVSB200_K1=2 python tools/microbench.py --cases one --reps 2 --kernels fwd > gpurun_out/plain_k1b.log 2>&1 && \
MVSB200_K1=2 ncu --set full --clock-control none --import-source on -k regex:warp_variance_fwd2 -s 3 -c 1 -f -o gpurun_out/k1_fwd2_r2 python tools/microbench.py --cases one --reps 2 --kernels fwd > gpurun_out/ncu_k1_fwd2.log 2>&1
echo "rc=$?" >> gpurun_out/ncu_k1_fwd2.log
tail -3 gpurun_out/ncu_k1_fwd3.log gpurun_out/ncu_k1_fwd2.log
