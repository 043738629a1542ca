#!/bin/bash
# round 2, GPU job 21: is the landslide cost the downstream marking walk?  kinematic + slides with and without sl_gullienogullie
mkdir -p gpurun_out
for v in "--no-sed" "--no-sed --no-gullie"; do echo "== $v"; timeout 300 python tools/profile_shia.py --workload c5 --members 8 $v --runs 2 2>&1 | grep -E "^run|Error"; done > gpurun_out/j21_gullie.log 2>&1
