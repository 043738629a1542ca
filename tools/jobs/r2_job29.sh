#!/bin/bash
# round 2, GPU job 29: boundary-graph jump rounds of small rasters in one CTA: topology parity, C1 timing
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_topo.py tests/test_gpu_fuzz.py tests/test_gpu_reference.py tests/test_gpu_subbasin.py tests/test_golden.py tests/test_gpu_simubasin_replay.py -x -q ) > gpurun_out/j29_pytest_topo.log 2>&1
timeout 300 python tools/profile_path_a.py --size 512 --runs 6 > gpurun_out/j29_patha_512.log 2>&1
timeout 300 python tools/profile_path_a.py --size 1024 --runs 4 > gpurun_out/j29_patha_1024.log 2>&1
timeout 300 python tools/profile_path_a.py --size 2048 --runs 4 > gpurun_out/j29_patha_2048.log 2>&1
