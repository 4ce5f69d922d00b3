timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/pytest_k2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_k2.log
tail -5 gpurun_out/pytest_k2.log
for f in 2 3; do echo "K2=$f"; MVSB200_K2=$f python tools/microbench.py --cases cfg --kernels bwd --reps 7 2>&1 | cut -c1-200; done > gpurun_out/k2lab.log 2>&1
for V in 5 7; do for f in 2 3; do echo "K2=$f V=$V"; MVSB200_K2=$f python - <<PY 2>&1 | cut -c1-200
import sys; sys.argv=['x']; sys.path.insert(0,'tools')
import microbench as m, torch
flush = torch.zeros(128*1024*1024, device='cuda:0')
m.case(1, $V, 256, 128, 160, 5, ['bwd'], flush)
PY
done; done >> gpurun_out/k2lab.log 2>&1
cat gpurun_out/k2lab.log
