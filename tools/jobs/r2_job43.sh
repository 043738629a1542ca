#!/bin/bash
# round 2, GPU job 43: DRAM bytes of every Path A kernel at 8192^2 (for roofline.traffic of extra.path_a); raster-wide HAND at C3
mkdir -p gpurun_out
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/j42_dram_patha.csv python tools/profile_path_a.py --size 8192 --runs 1 > gpurun_out/j42_ncu.log 2>&1
timeout 600 python - > gpurun_out/j43_hand_c3.json 2> gpurun_out/j43_hand_c3.err <<'PY'
import json, sys
sys.path.insert(0, ".")
import bench
from wmf_b200 import Context, CuModule
cu = CuModule(Context(0))
for size in (4096, 20000):
    r = bench.bench_path_a(cu, size, size, reps=2, cpu=False, hand=True)
    print(json.dumps(r), flush=True)
PY
