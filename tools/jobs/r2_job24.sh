#!/bin/bash
# round 2, GPU job 24: sediment kernel of wave w on a second stream next to the hydrology of wave w + 1: parity with the
# overlap forced on for every run, C5 throughput with one / two streams
mkdir -p gpurun_out
( WMF_SED_STREAMS=2 timeout 1200 python -m pytest tests/test_gpu_shia.py tests/test_gpu_fuzz.py tests/test_golden.py tests/test_gpu_reference.py -x -q -k "slide or sed or scenario or combination or fuzz_shia or golden or reference" -p no:cacheprovider ) > gpurun_out/j24_pytest_two.log 2>&1
( WMF_SED_STREAMS=2 timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q -k "c5" -p no:cacheprovider ) > gpurun_out/j24_pytest_c5_two.log 2>&1
for S in 1 2; do echo "streams $S"; WMF_SED_STREAMS=$S timeout 600 python tools/bench_scenarios.py --member-axis 1,4,8 2> gpurun_out/j24_scen_s$S.err | cut -c150-300; done > gpurun_out/j24_scen.log
