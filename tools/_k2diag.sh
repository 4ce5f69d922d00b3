for d in 0 1 2 3; do echo "diag=$d"; MVSB200_K2_DIAG=$d python tools/microbench.py --cases cfg --kernels bwd --reps 5 2>&1 | grep -E '"B": 4' ; done > gpurun_out/k2diag.log 2>&1
cat gpurun_out/k2diag.log
