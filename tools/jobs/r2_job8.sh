#!/bin/bash
# round 2, GPU job 8: re-measure after the session restart (gpurun_out of jobs 1-7 was lost): whole -m gpu suite,
# scores), Path A timings, the default bench line
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > gpurun_out/j8_pytest.log 2>&1
timeout 300 python tools/profile_path_a.py --size 8192 --runs 3 > gpurun_out/j8_patha_8192.log 2>&1
timeout 300 python tools/profile_path_a.py --size 20000 --runs 3 > gpurun_out/j8_patha_20000.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/j8_launches_patha.csv python tools/profile_path_a.py --size 8192 --runs 2 > gpurun_out/j8_ncu_launch.log 2>&1
( time timeout 900 python bench.py ) > gpurun_out/j8_bench.json 2> gpurun_out/j8_bench.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/j8_smoke.log 2>&1
ls -la gpurun_out > gpurun_out/j8_ls.txt
