#!/bin/bash
# round 2, GPU job 17: ncu of the two kernels of a C5 wave (8 scenarios): k_shia_tb<4,KIN,SED,SLIDES> (staging) and k_sed_tb<4>
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_shia_tb|k_sed_tb' --launch-skip 1600 --launch-count 2 \
   -o gpurun_out/j17_c5_split -f python tools/profile_shia.py --workload c5 --members 8 --runs 1 > gpurun_out/j17_ncu.log 2>&1
