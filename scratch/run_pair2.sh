timeout 120 ./scratch/mma2_bench
timeout 300 python scratch/run_pair_dbg.py 2>&1 | tail -80
