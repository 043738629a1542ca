#!/bin/bash
# round 2, GPU job 16: sediments in their own kernel (k_sed_tb behind every wave of k_shia_tb): parity, C5 throughput
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_shia.py tests/test_gpu_fuzz.py tests/test_golden.py tests/test_gpu_reference.py -x -q ) > gpurun_out/j16_pytest.log 2>&1
( timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -k "c5 or c4 or schedules" ) > gpurun_out/j16_pytest_full.log 2>&1
for lib in libwmf_b200.so libwmf_b200_sed2.so; do
  ( WMF_B200_LIB=$PWD/wmf_b200/$lib timeout 600 python bench.py --workload c5 --no-extras --steps 3 --warmup 2 ) > gpurun_out/j16_bench_c5_${lib%.so}.json 2> gpurun_out/j16_bench_c5_${lib%.so}.err
  WMF_B200_LIB=$PWD/wmf_b200/$lib timeout 600 python tools/bench_scenarios.py --member-axis 1,4,8 > gpurun_out/j16_scen_${lib%.so}.jsonl 2> gpurun_out/j16_scen_${lib%.so}.err
done
