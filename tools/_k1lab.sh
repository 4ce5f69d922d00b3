for m in 0 1 2; do echo "PIPE=$m"; MVSB200_K1_PIPE=$m python tools/microbench.py --cases cfg --kernels fwd --reps 7 2>&1 | cut -c1-200; done > gpurun_out/k1lab.log 2>&1
for V in 5 7; do for m in 0 1; do echo "PIPE=$m V=$V"; MVSB200_K1_PIPE=$m python - <<PY 2>&1 | cut -c1-200
import sys; sys.argv=['x']; sys.path.insert(0,'tools')
import microbench as m, torch
flush = torch.zeros(128*1024*1024, device='cuda:0')
m.case(1, $V, 256, 128, 160, 5, ['fwd'], flush)
PY
done; done >> gpurun_out/k1lab.log 2>&1
MVSB200_K1_PIPE=1 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -2 >> gpurun_out/k1lab.log
cat gpurun_out/k1lab.log
