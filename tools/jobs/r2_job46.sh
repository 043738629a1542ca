#!/bin/bash
# round 2, GPU job 46: geo_hand_global with the tile pass: parity, walk / jump_global / jump at 8192^2 and C3, element-wise equality
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_topo.py tests/test_gpu_reference.py tests/test_golden.py tests/test_gpu_fuzz.py -x -q -p no:cacheprovider -k "not shia" ) > gpurun_out/j46_pytest.log 2>&1
timeout 900 python - > gpurun_out/j46_hand.json 2> gpurun_out/j46_hand.err <<'PY'
import json, os, sys
sys.path.insert(0, ".")
import torch
import bench
from wmf_b200 import Context, CuModule, synth
cu = CuModule(Context(0))
for size in (8192, 20000):
    out = {}
    for algo in ("walk", "jump_global", "jump"):
        os.environ["WMF_HAND_ALGO"] = algo
        r = bench.bench_path_a(cu, size, size, reps=1, cpu=False, hand=True)
        out[algo] = r["hand"]["ms_geo_hand_global"]
    print(json.dumps({"size": size, **out}), flush=True)
for nc, nr, seed in ((3000, 3000, 3), (1500, 2100, 5)):
    d_dir = synth.make_dir_torch("G2", nc, nr, seed, device="cuda:0")
    g = torch.Generator(device="cuda:0"); g.manual_seed(seed)
    holes = torch.rand((nr, nc), device="cuda:0", generator=g) < 0.002
    d_dir = torch.where(holes, torch.full_like(d_dir, int(synth.NODATA)), d_dir)
    bad = torch.rand((nr, nc), device="cuda:0", generator=g) < 0.0005
    d_dir = torch.where(bad, torch.full_like(d_dir, 5), d_dir)                 # invalid codes too
    dem = torch.rand((nr, nc), device="cuda:0", generator=g) * 100
    red = (torch.rand((nr, nc), device="cuda:0", generator=g) < 0.01).int()
    red = torch.where(torch.rand((nr, nc), device="cuda:0", generator=g) < 0.001, torch.full_like(red, int(synth.NODATA)), red)
    cu.ncols, cu.nrows, cu.xll, cu.yll, cu.dx, cu.dy, cu.dxp, cu.nodata = nc, nr, 0.0, 0.0, 30.0, 30.0, 30.0, synth.NODATA
    res = {}
    for algo in ("walk", "jump_global", "jump"):
        os.environ["WMF_HAND_ALGO"] = algo
        res[algo] = cu.geo_hand_global(dem, d_dir.contiguous(), red.contiguous())
    print(json.dumps({"raster": [nc, nr], "walk_equals_jump": bool(torch.equal(res["walk"], res["jump"])),
                      "walk_equals_jump_global": bool(torch.equal(res["walk"], res["jump_global"])),
                      "with_value": int((res["jump"] != synth.NODATA).sum().item())}), flush=True)
PY
