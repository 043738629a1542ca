#!/bin/bash
# round 2, GPU job 31: per-kernel launch list of Path A at C3 after the tuning of this round
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/j31_launches_patha_c3.csv python tools/profile_path_a.py --size 20000 --runs 2 > gpurun_out/j31_ncu_launch.log 2>&1
