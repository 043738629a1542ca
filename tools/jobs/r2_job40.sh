#!/bin/bash
# round 2, GPU job 40: last check of the tree: whole -m gpu suite, smoke, C5 line
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > gpurun_out/j40_pytest.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/j40_smoke.log 2>&1
( timeout 600 python bench.py --workload c5 --no-extras --steps 5 --warmup 3 ) > gpurun_out/j40_bench_c5.json 2> gpurun_out/j40_bench_c5.err
