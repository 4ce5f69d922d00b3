for i in 1 2 3; do timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -1 >> gpurun_out/pytest_gpu_rep.log; done
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench18.json 2> gpurun_out/bench18.err; echo "rc=$?" >> gpurun_out/bench18.err
