timeout 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_parity.py tests/test_gpu_graph.py tests/test_gpu_depth_slab.py -m gpu -q -x > gpurun_out/pytest_k3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_k3.log
tail -5 gpurun_out/pytest_k3.log
timeout 500 python bench.py --no-cpu-baseline --no-extras > gpurun_out/bench_k3.json 2> gpurun_out/bench_k3.err; echo "rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_k3.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'])
r=d['roofline']; print('family',r['achieved'],r['frac'],r['ms_per_step'])
for n,v in r['per_kernel'].items(): print(' ',n,round(v['ms_per_step'],3),round(v['TFLOPs'],1),round(v['frac'],3))
PY
