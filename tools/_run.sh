( time python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2> gpurun_out/bench_default.time; echo "rc=$?" >> gpurun_out/bench_default.err
( time python bench.py --impl reference > gpurun_out/bench_default_ref.json 2> gpurun_out/bench_default_ref.err ) 2> gpurun_out/bench_default_ref.time; echo "rc=$?" >> gpurun_out/bench_default_ref.err
( time python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1 ) 2> gpurun_out/smoke.time
nproc > gpurun_out/nproc.txt
