for m in 0 1 2 3 4; do echo "K1=2 MIX=$m"; MVSB200_K1=2 MVSB200_K1_MIX=$m python tools/microbench.py --cases cfg --kernels fwd --reps 7 2>&1 | cut -c1-200; done > gpurun_out/k1lab.log 2>&1
MVSB200_K1=2 MVSB200_K1_MIX=4 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cost_volume" 2>&1 | tail -2 >> gpurun_out/k1lab.log
cat gpurun_out/k1lab.log
