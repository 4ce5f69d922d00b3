python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python tools/profile_step.py --B 1 --rows 40 > gpurun_out/prof_step_b1_v4.txt 2>&1
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench4.json 2> gpurun_out/bench4.err
