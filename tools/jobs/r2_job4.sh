#!/bin/bash
# round 2, GPU job 4: cp.async staging of the time block's inputs in k_shia_tb (A/B against a build without it),
# the SimuBasin replay test, C5 through bench.py
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_shia.py tests/test_gpu_simubasin_replay.py tests/test_gpu_fuzz.py tests/test_golden.py -x -q ) > gpurun_out/j4_pytest.log 2>&1
( timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q -k "schedules or c2_basin or c5" ) > gpurun_out/j4_pytest_full.log 2>&1
for lib in libwmf_b200.so libwmf_b200_nostage.so; do
  for wl in c2 c5; do
    ( WMF_B200_LIB=$PWD/wmf_b200/$lib timeout 600 python bench.py --workload $wl --no-extras --steps 5 --warmup 3 ) > gpurun_out/j4_bench_${wl}_${lib%.so}.json 2> gpurun_out/j4_bench_${wl}_${lib%.so}.err
  done
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shia_tb --launch-skip 700 --launch-count 1 \
   -o gpurun_out/j4_c2_stage -f python tools/profile_shia.py --workload c2 --runs 1 > gpurun_out/j4_ncu_c2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shia_tb --launch-skip 300 --launch-count 1 \
   -o gpurun_out/j4_c5_sed -f python tools/profile_shia.py --workload c5 --runs 1 > gpurun_out/j4_ncu_c5.log 2>&1
ls -la gpurun_out > gpurun_out/j4_ls.txt
