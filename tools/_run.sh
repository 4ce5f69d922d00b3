timeout 300 python tools/profile_step.py --rows 70 > gpurun_out/prof_step_b4_r4.txt 2>&1
