#!/bin/bash
# round 2, GPU job 42: DRAM bytes of every Path A kernel at 8192^2 (one find + cut + acum), for roofline.traffic of extra.path_a
mkdir -p gpurun_out
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/j42_dram_patha.csv python tools/profile_path_a.py --size 8192 --runs 1 > gpurun_out/j42_ncu.log 2>&1
