#!/bin/bash
# round 2, GPU job 11: tile path tuning (4096-item radix tiles, 128-bit histogram loads, batched loads in emit / final_write,
# settled boundary nodes skipped, division-free tile setup): parity, timings; SED variant with 3 blocks per SM
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_topo.py tests/test_gpu_fuzz.py tests/test_gpu_reference.py tests/test_gpu_subbasin.py -x -q ) > gpurun_out/j11_pytest_topo.log 2>&1
( timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -k "path_a" ) > gpurun_out/j11_pytest_full.log 2>&1
timeout 300 python tools/profile_path_a.py --size 8192 --runs 3 > gpurun_out/j11_patha_8192.log 2>&1
timeout 300 python tools/profile_path_a.py --size 20000 --runs 3 > gpurun_out/j11_patha_20000.log 2>&1
timeout 300 python tools/profile_path_a.py --size 512 --runs 5 > gpurun_out/j11_patha_512.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/j11_launches_patha.csv python tools/profile_path_a.py --size 8192 --runs 2 > gpurun_out/j11_ncu_launch.log 2>&1
for lib in libwmf_b200.so libwmf_b200_sed3.so; do
  ( WMF_B200_LIB=$PWD/wmf_b200/$lib timeout 600 python bench.py --workload c5 --no-extras --steps 3 --warmup 2 ) > gpurun_out/j11_bench_c5_${lib%.so}.json 2> gpurun_out/j11_bench_c5_${lib%.so}.err
  WMF_B200_LIB=$PWD/wmf_b200/$lib timeout 600 python tools/bench_scenarios.py --member-axis 8 > gpurun_out/j11_scen8_${lib%.so}.jsonl 2> gpurun_out/j11_scen8_${lib%.so}.err
done
ls -la gpurun_out > gpurun_out/j11_ls.txt
