#!/bin/bash
# round 2, GPU job 26: compute-sanitizer (memcheck, then initcheck) over the kernels written this round: the tile path of
# basin_find, k_sed_tb, the landslide ring, the two-stream schedule
mkdir -p gpurun_out
export WMF_SHIA_K=4
( timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_shia.py -x -q -p no:cacheprovider -k "test_sediments or test_slides or scenario_ensemble or combinations" ) > gpurun_out/j26_memcheck_shia.log 2>&1
echo "exit $?" >> gpurun_out/j26_memcheck_shia.log
( timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_topo.py -x -q -p no:cacheprovider ) > gpurun_out/j26_memcheck_topo.log 2>&1
echo "exit $?" >> gpurun_out/j26_memcheck_topo.log
( timeout 900 compute-sanitizer --tool initcheck --error-exitcode 9 python -m pytest tests/test_gpu_shia.py -x -q -p no:cacheprovider -k "test_sediments or scenario_ensemble" ) > gpurun_out/j26_initcheck_shia.log 2>&1
echo "exit $?" >> gpurun_out/j26_initcheck_shia.log
( timeout 900 compute-sanitizer --tool initcheck --error-exitcode 9 python -m pytest tests/test_gpu_topo.py -x -q -p no:cacheprovider -k "kat or c1 or partial" ) > gpurun_out/j26_initcheck_topo.log 2>&1
echo "exit $?" >> gpurun_out/j26_initcheck_topo.log
