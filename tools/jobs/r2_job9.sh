#!/bin/bash
# round 2, GPU job 9: per-kernel launch list of Path A at C3, C5 through bench.py, the scenario member axis, C2 alone
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/j9_launches_patha_c3.csv python tools/profile_path_a.py --size 20000 --runs 2 > gpurun_out/j9_ncu_launch.log 2>&1
( timeout 600 python bench.py --workload c5 --no-extras --steps 5 --warmup 3 ) > gpurun_out/j9_bench_c5.json 2> gpurun_out/j9_bench_c5.err
( timeout 600 python bench.py --workload c2 --no-extras --steps 5 --warmup 3 ) > gpurun_out/j9_bench_c2.json 2> gpurun_out/j9_bench_c2.err
timeout 900 python tools/bench_scenarios.py --member-axis 1,2,4,8 > gpurun_out/j9_scen_member.jsonl 2> gpurun_out/j9_scen_member.err
timeout 300 python tools/profile_path_a.py --size 8192 --runs 3 --basics > gpurun_out/j9_patha_8192_basics.log 2>&1
ls -la gpurun_out > gpurun_out/j9_ls.txt
