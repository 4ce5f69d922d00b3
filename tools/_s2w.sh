MVSB200_S2_WGRAD=tcgen05 timeout 500 python bench.py --no-cpu-baseline --no-extras > gpurun_out/bench_s2w.json 2> gpurun_out/bench_s2w.err; echo "rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_s2w.json'))
print('value',d['value'],'ms',d['ms_per_step'])
k=d['kernels']
for n,v in sorted(k.items(), key=lambda kv:-kv[1]['ms_per_step'])[:8]: print(n, round(v['ms_per_step'],3), v['launches']//d['steps'])
print(sum(v['ms_per_step'] for v in k.values()))
PY
