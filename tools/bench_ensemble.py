#!/usr/bin/env python
"""Calibration-ensemble throughput (BASELINE.json configs[3] shape): M SHIA parameter sets on one ~250k-cell basin in ONE
wavefront launch sequence per batch (member-minor state layout), everything resident in HBM, scores computed on the
device.  Prints one JSON line per batch size."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--members", default="1,8,64,256")
    ap.add_argument("--reps", type=int, default=2)
    a = ap.parse_args()
    import torch
    from wmf_b200 import Context, CuModule, ModelsModule, synth
    ctx = Context(0)
    cu, m = CuModule(ctx), ModelsModule(ctx)
    nc = nr = a.size
    DIR, DEM = synth.make_raster("G2", nc, nr, seed=0)
    cu.ncols, cu.nrows, cu.xll, cu.yll, cu.dx, cu.dy, cu.dxp, cu.nodata = nc, nr, 0.0, 0.0, 30.0, 30.0, 30.0, synth.NODATA
    x, y = synth.outlet_xy(nc, nr, 0.0, 0.0, 30.0)
    n = cu.basin_find(x, y, DIR, nc, nr)
    b = cu.basin_cut(n)
    dv = cu.basin_map2basin(b, DIR.astype(np.float32), 0.0, 0.0, 30.0, 30.0, synth.NODATA).astype(np.int32)
    de = cu.basin_map2basin(b, DEM, 0.0, 0.0, 30.0, 30.0, synth.NODATA)
    acum, lng, pend, _ = cu.basin_basics(b, de, dv)
    T = a.steps
    inp = synth.make_shia_inputs(b, acum, lng, pend, dxp=30.0, dt=3600.0, n_reg=T, seed=0, n_control=16)
    synth.apply_to_models(m, inp)
    m.make_resident()
    d_rain = torch.from_numpy(inp["rain"]).cuda()
    n_cont = int(np.count_nonzero(inp["control"])) + 1
    peak = 6453.1
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    qobs = None
    for M in [int(v) for v in a.members.split(",")]:
        calibs = synth.make_ensemble_calibs(max(M, 2), seed=0)[:M]
        best = None
        for _ in range(a.reps + 1):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = m._run(calibs, n_cont, 1, T, d_rain, inp["pos_evento"], device_out=True, want={"q"})
            if qobs is None:
                qobs = out["q"].reshape(-1, T, n_cont)[0, :, 0].contiguous()
            sc = torch.empty((M, 2), dtype=torch.float32, device="cuda")
            ctx.check(ctx.L.wmf_eval_scores(ctx.h, out["q"].data_ptr(), M, n_cont, T, 0, qobs.data_ptr(), sc.data_ptr()))
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        wave_ms = ctx.last_kernel_ms()
        cs = float(n) * T * M
        print(json.dumps({"cells": int(n), "steps": T, "members": M, "ms": best * 1e3, "wave_ms": wave_ms,
                          "cell_steps_per_s": cs / best, "algorithmic_gbs_80B": 80.0 * cs / (wave_ms / 1e3) / 1e9,
                          "frac_of_measured_hbm_80B": 80.0 * cs / (wave_ms / 1e3) / 1e9 / peak,
                          "nash_member0": float(sc[0, 0])}), flush=True)


if __name__ == "__main__":
    main()
