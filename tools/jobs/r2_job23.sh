#!/bin/bash
# round 2, GPU job 23: verification of the current tree: whole -m gpu suite, smoke, the default bench line, C5 / C2 lines,
# the ncu launch list of the bench command
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > gpurun_out/j23_pytest.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/j23_smoke.log 2>&1
( time timeout 900 python bench.py ) > gpurun_out/j23_bench.json 2> gpurun_out/j23_bench.err
( timeout 600 python bench.py --workload c5 --no-extras --steps 5 --warmup 3 ) > gpurun_out/j23_bench_c5.json 2> gpurun_out/j23_bench_c5.err
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/j23_bench_ref.json 2> gpurun_out/j23_bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/j23_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-extras --cpu-members 1 > gpurun_out/j23_ncu_bench.log 2>&1
ls -la gpurun_out > gpurun_out/j23_ls.txt
