#!/bin/bash
# round 2, GPU job 10: radix scatter through shared memory + coalesced final_write: parity, timings, ncu source pages
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_topo.py tests/test_gpu_fuzz.py tests/test_gpu_reference.py tests/test_gpu_subbasin.py -x -q ) > gpurun_out/j10_pytest_topo.log 2>&1
( timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -k "path_a" ) > gpurun_out/j10_pytest_full.log 2>&1
timeout 300 python tools/profile_path_a.py --size 8192 --runs 3 > gpurun_out/j10_patha_8192.log 2>&1
timeout 300 python tools/profile_path_a.py --size 20000 --runs 3 > gpurun_out/j10_patha_20000.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/j10_launches_patha.csv python tools/profile_path_a.py --size 8192 --runs 2 > gpurun_out/j10_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_pf_local2|k_rs_scatter|k_pf_emit|k_pf_final_write|k_pf_local1|k_rs_hist' --launch-skip 0 --launch-count 8 \
   -o gpurun_out/j10_pf_r2 -f python tools/profile_path_a.py --size 8192 --runs 1 > gpurun_out/j10_ncu_pf.log 2>&1
ls -la gpurun_out > gpurun_out/j10_ls.txt
