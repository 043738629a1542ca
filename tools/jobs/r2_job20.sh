#!/bin/bash
# round 2, GPU job 20: how much of the C5 slowdown is K = 4: the kinematic run of the C5 basin (8 members) at K = 4 / 8 / 16
mkdir -p gpurun_out
for K in 4 8 16; do echo "K=$K"; WMF_SHIA_K=$K timeout 300 python tools/profile_shia.py --workload c5 --members 8 --no-sed --no-slides --runs 2 2>&1 | grep -E "^run"; done > gpurun_out/j20_kin_K.log 2>&1
