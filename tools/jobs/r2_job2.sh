#!/bin/bash
# round 2, GPU job 2: first run of the tile path of basin_find (csrc/pathfind.cu): stage-by-stage debug, topology tests,
# timings at 8192^2 and C3, member-minor kernel recapture.
mkdir -p gpurun_out
{
for args in "--kind partial --nc 150 --nr 130 --seed 0" "--kind partial --nc 70 --nr 60 --seed 3" "--kind G2 --nc 200 --nr 140 --seed 1" "--kind G1 --nc 130 --nr 129 --seed 2" "--kind partial --nc 260 --nr 200 --seed 5"; do
  timeout 300 python tools/debug_pathfind.py $args
done
} > gpurun_out/j2_debug.log 2>&1
( timeout 900 python -m pytest tests/test_gpu_topo.py tests/test_gpu_fuzz.py tests/test_gpu_reference.py tests/test_gpu_subbasin.py -x -q -k "not shia" ) > gpurun_out/j2_pytest_topo.log 2>&1
( timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -k "path_a" ) > gpurun_out/j2_pytest_full.log 2>&1
timeout 300 python tools/profile_path_a.py --size 8192 --runs 3 > gpurun_out/j2_patha_8192.log 2>&1
timeout 300 python tools/profile_path_a.py --size 20000 --runs 3 > gpurun_out/j2_patha_20000.log 2>&1
WMF_FIND_ALGO=bfs WMF_ACUM_ALGO=climb timeout 300 python tools/profile_path_a.py --size 8192 --runs 2 > gpurun_out/j2_patha_8192_bfs.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/j2_launches_patha.csv python tools/profile_path_a.py --size 8192 --runs 2 > gpurun_out/j2_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shia_tb --launch-skip 280 --launch-count 1 \
   -o gpurun_out/j2_ens_r2 -f python tools/profile_shia.py --workload c4 --members 256 --runs 1 > gpurun_out/j2_ncu_ens.log 2>&1
ls -la gpurun_out > gpurun_out/j2_ls.txt
