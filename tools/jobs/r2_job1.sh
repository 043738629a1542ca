#!/bin/bash
# round 2, GPU job 1: the whole -m gpu suite, the new bench line, baseline ncu counters of the Path A kernels and of the
# member-minor SHIA kernel.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/j1_smi.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/j1_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/j1_pytest.log
( time timeout 900 python bench.py --steps 3 --warmup 2 ) > gpurun_out/j1_bench.json 2> gpurun_out/j1_bench.err
echo "bench exit $?" >> gpurun_out/j1_bench.err
timeout 300 python tools/profile_path_a.py --size 8192 --runs 2 --basics > gpurun_out/j1_patha_8192.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_basin_find|k_acum_climb|k_basics_cell|k_basics_reduce|k_geo_hand|k_basin_cut' \
   -o gpurun_out/j1_patha_r2 -f python tools/profile_path_a.py --size 8192 --runs 1 --basics > gpurun_out/j1_ncu_patha.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shia_tb --launch-skip 500 --launch-count 1 \
   -o gpurun_out/j1_ens_r2 -f python tools/profile_shia.py --workload c4 --members 256 --runs 1 --steps 200 > gpurun_out/j1_ncu_ens.log 2>&1
ls -la gpurun_out > gpurun_out/j1_ls.txt
