#!/bin/bash
# round 2, GPU job 13: switch combinations behind one shia_v1 call (new test + fuzz without switched-off means),
# cp.async staging A/B on C2 (two builds of the library), ensemble bench line with the ncu-derived traffic
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_shia.py tests/test_gpu_fuzz.py -x -q ) > gpurun_out/j13_pytest.log 2>&1
for lib in libwmf_b200.so libwmf_b200_stage.so; do
  ( WMF_B200_LIB=$PWD/wmf_b200/$lib timeout 600 python bench.py --workload c2 --no-extras --steps 10 --warmup 3 ) > gpurun_out/j13_bench_c2_${lib%.so}.json 2> gpurun_out/j13_bench_c2_${lib%.so}.err
done
for lib in libwmf_b200_stage.so libwmf_b200.so; do
  ( WMF_B200_LIB=$PWD/wmf_b200/$lib timeout 600 python bench.py --workload c2 --no-extras --steps 10 --warmup 3 ) > gpurun_out/j13_bench_c2_b_${lib%.so}.json 2> gpurun_out/j13_bench_c2_b_${lib%.so}.err
done
ls -la gpurun_out > gpurun_out/j13_ls.txt
