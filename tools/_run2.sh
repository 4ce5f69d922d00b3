export NCCL_DEBUG=WARN
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 tools/run_depth_slab.py --reps 3 --profile > gpurun_out/slab2_prof.json 2> gpurun_out/slab2_prof.err; echo "rc=$?" >> gpurun_out/slab2_prof.err
