#!/bin/bash
# round 2, GPU job 45: geo_hand_global at C3, warm pools, walk vs jump
mkdir -p gpurun_out
timeout 900 python - > gpurun_out/j45_hand.json 2> gpurun_out/j45_hand.err <<'PY'
import json, os, sys
sys.path.insert(0, ".")
import bench
from wmf_b200 import Context, CuModule
cu = CuModule(Context(0))
for size in (8192, 20000):
    out = {}
    for algo in ("walk", "jump"):
        os.environ["WMF_HAND_ALGO"] = algo
        r = bench.bench_path_a(cu, size, size, reps=1, cpu=False, hand=True)
        out[algo] = {k: v for k, v in r["hand"].items() if k != "what"}
    print(json.dumps({"size": size, **out}), flush=True)
PY
