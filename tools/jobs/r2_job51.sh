#!/bin/bash
# round 2, GPU job 51: ncu --set full of the tile-path kernels as they stand at the end of the round (8192^2) and of the
# geo_hand_global kernels
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_pf_local2|k_rs_scatter|k_pf_emit|k_pf_final_write|k_pf_local1|k_rs_hist|k_pf_bd_jump$' --launch-skip 0 --launch-count 10 \
   -o gpurun_out/j51_pf_final -f python tools/profile_path_a.py --size 8192 --runs 1 > gpurun_out/j51_ncu_pf.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_hg_' --launch-skip 0 --launch-count 4 \
   -o gpurun_out/j51_hg -f python - > gpurun_out/j51_ncu_hg.log 2>&1 <<'PY'
import sys
sys.path.insert(0, ".")
import bench
from wmf_b200 import Context, CuModule
cu = CuModule(Context(0))
bench.bench_path_a(cu, 8192, 8192, reps=1, cpu=False, hand=True)
PY
