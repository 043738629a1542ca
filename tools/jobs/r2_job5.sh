#!/bin/bash
# round 2, GPU job 5: scenario/member axis with sediments + slides (parity test, C5 throughput), hoisted So**1.664
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_shia.py tests/test_gpu_fuzz.py tests/test_golden.py tests/test_gpu_reference.py -x -q ) > gpurun_out/j5_pytest.log 2>&1
( timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q -k "c5 or c4" ) > gpurun_out/j5_pytest_full.log 2>&1
( timeout 600 python bench.py --workload c5 --no-extras --steps 5 --warmup 3 ) > gpurun_out/j5_bench_c5.json 2> gpurun_out/j5_bench_c5.err
timeout 900 python tools/bench_scenarios.py --member-axis 1,2,3,4,6,8 > gpurun_out/j5_scen_member.jsonl 2> gpurun_out/j5_scen_member.err
timeout 600 python tools/bench_scenarios.py --concurrency 1,3 > gpurun_out/j5_scen_ctx.jsonl 2> gpurun_out/j5_scen_ctx.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shia_tb --launch-skip 300 --launch-count 1 \
   -o gpurun_out/j5_c5_scen4 -f python tools/bench_scenarios.py --member-axis 4 > gpurun_out/j5_ncu_scen.log 2>&1
ls -la gpurun_out > gpurun_out/j5_ls.txt
