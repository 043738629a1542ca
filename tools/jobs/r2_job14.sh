#!/bin/bash
# round 2, GPU job 14: where the C5 time goes -- the C5 basin with linear speeds, kinematic speeds only, + sediments, + slides, + both
mkdir -p gpurun_out
for M in 1 8; do
for v in "--no-sed --no-slides --linear" "--no-sed --no-slides" "--no-slides" "--no-sed" ""; do
  echo "== members $M $v"
  timeout 300 python tools/profile_shia.py --workload c5 --members $M $v --runs 2 2>&1 | grep -E "^run|qmax|Error|error"
done; done > gpurun_out/j14_c5_variants.log 2>&1
