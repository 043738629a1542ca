#!/bin/bash
# round 2, GPU job 19: landslide constants hoisted per block: slide parity tests, C5 throughput
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_shia.py tests/test_gpu_fuzz.py tests/test_golden.py -x -q -k "slide or sed or scenario or combination or fuzz_shia or golden" ) > gpurun_out/j19_pytest.log 2>&1
timeout 600 python tools/bench_scenarios.py --member-axis 1,8 > gpurun_out/j19_scen.jsonl 2> gpurun_out/j19_scen.err
timeout 300 python tools/profile_shia.py --workload c5 --members 8 --no-sed --runs 2 2>&1 | grep -E "^run" > gpurun_out/j19_kin_slides.log
