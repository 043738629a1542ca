#!/usr/bin/env python
"""e2e probe: time models.shia_v1 with host (pinned) buffers on the bench workload, for tuning the rain streaming."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from wmf_b200 import Context, CuModule, ModelsModule, synth
ctx = Context(0)
cu, m = CuModule(ctx), ModelsModule(ctx)
wl = bench.build_workload("c2", cu)
inp = wl["inp"]
synth.apply_to_models(m, inp)
pin = torch.from_numpy(inp["rain"]).pin_memory()
m.set_rain(pin.numpy(), inp["pos_evento"])
n_cont = int(np.count_nonzero(inp["control"])) + 1
for i in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ret = m.shia_v1(None, None, inp["calib"], n_cont, 1, wl["T"], None, None, wl["n"])
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"e2e {dt*1e3:.1f} ms  (wave loop {ctx.last_kernel_ms():.1f} ms)")
