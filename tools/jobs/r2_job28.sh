#!/bin/bash
# round 2, GPU job 28: rain statistics in one pass (k_rain_stats): parity of the SHIA tests, setup time on C2, floods combination test
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_shia.py tests/test_gpu_fuzz.py tests/test_golden.py tests/test_gpu_reference.py tests/test_gpu_simubasin_replay.py tests/test_gpu_rain.py -x -q ) > gpurun_out/j28_pytest.log 2>&1
( timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -k "c2 or schedules" ) > gpurun_out/j28_pytest_full.log 2>&1
( timeout 600 python bench.py --workload c2 --no-extras --steps 10 --warmup 3 ) > gpurun_out/j28_bench_c2.json 2> gpurun_out/j28_bench_c2.err
( timeout 600 python bench.py --workload c5 --no-extras --steps 5 --warmup 3 ) > gpurun_out/j28_bench_c5.json 2> gpurun_out/j28_bench_c5.err
