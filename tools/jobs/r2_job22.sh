#!/bin/bash
# round 2, GPU job 22: K = 8 time blocks for the sediment / landslide variants (members >= 4): parity, C5 throughput at K = 4 and 8
mkdir -p gpurun_out
( WMF_SHIA_K=8 timeout 900 python -m pytest tests/test_gpu_shia.py -x -q -k "slide or sed or scenario" -p no:cacheprovider ) > gpurun_out/j22_pytest_k8.log 2>&1
for K in 4 8; do WMF_SHIA_K=$K timeout 600 python tools/bench_scenarios.py --member-axis 1,8 2> gpurun_out/j22_scen_k$K.err | cut -c150-300; done > gpurun_out/j22_scen.log
