#!/bin/bash
# round 2, GPU job 48: final check of the tree: whole -m gpu suite, smoke, the default bench line (with extra.c2 / c5 / path_a + HAND)
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > gpurun_out/j48_pytest.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/j48_smoke.log 2>&1
( time timeout 900 python bench.py ) > gpurun_out/j48_bench.json 2> gpurun_out/j48_bench.err
