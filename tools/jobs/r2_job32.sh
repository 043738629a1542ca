#!/bin/bash
# round 2, GPU job 32: exit cells renumbered densely for the boundary jump rounds: parity, stage debug tool, timings
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_topo.py tests/test_gpu_fuzz.py tests/test_gpu_reference.py tests/test_gpu_subbasin.py tests/test_golden.py tests/test_gpu_simubasin_replay.py -x -q ) > gpurun_out/j32_pytest_topo.log 2>&1
( timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -k "path_a" ) > gpurun_out/j32_pytest_full.log 2>&1
timeout 300 python tools/debug_pathfind.py --kind partial --nc 150 --nr 130 --seed 1 > gpurun_out/j32_debug.log 2>&1
timeout 300 python tools/profile_path_a.py --size 512 --runs 5 > gpurun_out/j32_patha_512.log 2>&1
timeout 300 python tools/profile_path_a.py --size 8192 --runs 3 > gpurun_out/j32_patha_8192.log 2>&1
timeout 300 python tools/profile_path_a.py --size 20000 --runs 3 > gpurun_out/j32_patha_20000.log 2>&1
