python tools/profile_copies.py > gpurun_out/copies_r2.log 2>&1; tail -35 gpurun_out/copies_r2.log
python tools/profile_step.py --rows 40 > gpurun_out/step_r2.log 2>&1; head -60 gpurun_out/step_r2.log | cut -c1-200
