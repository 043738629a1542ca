#!/bin/bash
# round 2, GPU job 18: k_sed_tb with the next step's inputs prefetched (3 and 2 CTAs per SM): parity of the sediment tests, C5 throughput
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_shia.py -x -q -k "sed or scenario or combination" ) > gpurun_out/j18_pytest.log 2>&1
for lib in libwmf_b200.so libwmf_b200_sed2.so; do
  WMF_B200_LIB=$PWD/wmf_b200/$lib timeout 600 python tools/bench_scenarios.py --member-axis 1,8 > gpurun_out/j18_scen_${lib%.so}.jsonl 2> gpurun_out/j18_scen_${lib%.so}.err
done
