for d in 0 1 2; do echo "dbg=$d"; MVSB200_S2WG_DBG=$d timeout 300 python tools/bench_s2_wgrad.py 2>&1 | cut -c1-190; done
