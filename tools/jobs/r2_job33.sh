#!/bin/bash
# round 2, GPU job 33: the alternative code paths still agree: level-by-level basin_find (WMF_FIND_ALGO=bfs) + atomic climb,
# sediments on one stream, one-step kernel (WMF_SHIA_K=0)
mkdir -p gpurun_out
( WMF_FIND_ALGO=bfs WMF_ACUM_ALGO=climb timeout 900 python -m pytest tests/test_gpu_topo.py tests/test_gpu_fuzz.py -x -q -p no:cacheprovider -k "not shia" ) > gpurun_out/j33_pytest_bfs.log 2>&1
( WMF_SED_STREAMS=1 timeout 900 python -m pytest tests/test_gpu_shia.py -x -q -p no:cacheprovider -k "sed or slide or scenario or combination" ) > gpurun_out/j33_pytest_onestream.log 2>&1
( WMF_SHIA_K=0 timeout 900 python -m pytest tests/test_gpu_shia.py tests/test_gpu_fuzz.py -x -q -p no:cacheprovider -k "baseline or sediments or slides or fuzz_shia or kinematic" ) > gpurun_out/j33_pytest_wave.log 2>&1
