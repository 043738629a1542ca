#!/bin/bash
# round 2, GPU job 35: occupancy variants of the tile-path kernels: per-kernel times at 8192^2 and C3
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_topo.py -x -q -p no:cacheprovider ) > gpurun_out/j35_pytest_topo.log 2>&1
timeout 300 python tools/profile_path_a.py --size 8192 --runs 3 > gpurun_out/j35_patha_8192.log 2>&1
timeout 300 python tools/profile_path_a.py --size 20000 --runs 3 > gpurun_out/j35_patha_20000.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/j35_launches_patha.csv python tools/profile_path_a.py --size 8192 --runs 2 > gpurun_out/j35_ncu_launch.log 2>&1
