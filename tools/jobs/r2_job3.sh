#!/bin/bash
# round 2, GPU job 3: tile path after the per-tile emit + outlet-cycle check, SHIA plan cache; ncu of the new kernels
mkdir -p gpurun_out
{
for args in "--kind partial --nc 150 --nr 130 --seed 1" "--kind G2 --nc 300 --nr 240 --seed 4"; do
  timeout 300 python tools/debug_pathfind.py $args
done
} > gpurun_out/j3_debug.log 2>&1
( timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/j3_pytest.log 2>&1
timeout 300 python tools/profile_path_a.py --size 8192 --runs 3 > gpurun_out/j3_patha_8192.log 2>&1
timeout 300 python tools/profile_path_a.py --size 20000 --runs 3 > gpurun_out/j3_patha_20000.log 2>&1
( timeout 600 python bench.py --workload c2 --no-extras --steps 5 --warmup 3 ) > gpurun_out/j3_bench_c2.json 2> gpurun_out/j3_bench_c2.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_pf_local2|k_rs_scatter|k_pf_emit|k_pf_final_write|k_pf_local1|k_rs_hist|k_pf_bd_jump' --launch-skip 0 --launch-count 12 \
   -o gpurun_out/j3_pf_r2 -f python tools/profile_path_a.py --size 8192 --runs 1 > gpurun_out/j3_ncu_pf.log 2>&1
ls -la gpurun_out > gpurun_out/j3_ls.txt
