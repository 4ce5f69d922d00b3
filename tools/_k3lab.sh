timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_k3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_k3.log
tail -5 gpurun_out/pytest_k3.log

timeout 500 python bench.py --no-cpu-baseline --no-extras > gpurun_out/bench_k3.json 2> gpurun_out/bench_k3.err; echo "rc=$?"; tail -3 gpurun_out/bench_k3.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_k3.json'))
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'])
k=d['kernels']
for n,v in sorted(k.items(), key=lambda kv:-kv[1]['ms_per_step']): print(n, round(v['ms_per_step'],3), v['launches']//d['steps'])
print(sum(v['ms_per_step'] for v in k.values()))
PY
