#!/bin/bash
# round 2, GPU job 15: source-level ncu capture of the SED + slides kernel with 8 scenarios on the member axis (C5)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_shia_tb --launch-skip 800 --launch-count 1 \
   -o gpurun_out/j15_c5_m8 -f python tools/profile_shia.py --workload c5 --members 8 --runs 1 > gpurun_out/j15_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_shia_tb --launch-skip 800 --launch-count 1 \
   -o gpurun_out/j15_c5_m8_kin -f python tools/profile_shia.py --workload c5 --members 8 --no-sed --no-slides --runs 1 > gpurun_out/j15_ncu_kin.log 2>&1
