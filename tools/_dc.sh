for m in fused kc; do echo "DECONV=$m"; MVSB200_DECONV=$m timeout 300 python -m pytest tests/test_gpu_conv_tc.py -q -x -k "transposed_conv and fused" 2>&1 | tail -1
MVSB200_DECONV=$m timeout 500 python bench.py --no-cpu-baseline --no-extras > gpurun_out/bench_dc_$m.json 2> gpurun_out/bench_dc_$m.err; python - <<PY
import json
d=json.load(open('gpurun_out/bench_dc_$m.json'))
print('value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'deconv',round(d['kernels']['deconv3d_s2_tc']['ms_per_step'],3), d['kernels']['deconv3d_s2_tc']['launches'])
PY
done
