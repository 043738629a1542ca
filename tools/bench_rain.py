#!/usr/bin/env python
"""Rain interpolation at scale: rain_idw / rain_mit on the C2 basin (1024x1024 G2 raster, 1000 steps), record table kept in
HBM, then the "gauges -> hydrograph" chain rain_idw(resident) + shia_v1.  Next to it the reference's own Fortran
(oracle/_ref) on a bounded sample of the same steps.  One JSON line per measurement."""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--stations", type=int, default=46)
    ap.add_argument("--cpu-steps", type=int, default=12)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    import torch
    from scipy.spatial import Delaunay
    from wmf_b200 import Context, CuModule, ModelsModule, synth
    ctx = Context(0)
    cu = CuModule(ctx)
    nc = nr = a.size
    DX = 30.0
    DIR, DEM = synth.make_raster("G2", nc, nr, seed=0)
    cu.ncols, cu.nrows, cu.xll, cu.yll, cu.dx, cu.dy, cu.dxp, cu.nodata = nc, nr, 0.0, 0.0, DX, DX, DX, synth.NODATA
    x, y = synth.outlet_xy(nc, nr, 0.0, 0.0, DX)
    n = cu.basin_find(x, y, DIR, nc, nr)
    b = cu.basin_cut(n)
    dv = cu.basin_map2basin(b, DIR.astype(np.float32), 0.0, 0.0, DX, DX, synth.NODATA).astype(np.int32)
    de = cu.basin_map2basin(b, DEM, 0.0, 0.0, DX, DX, synth.NODATA)
    acum, lng, pend, _ = cu.basin_basics(b, de, dv)
    cx, cy = cu.basin_coordxy(b)
    xy = np.asfortranarray(np.vstack([cx, cy]), dtype=np.float32)
    coord, rain = synth.make_stations(nc, nr, DX, a.stations, a.steps, 0)
    ns, T = coord.shape[1], a.steps
    inp = synth.make_shia_inputs(b, acum, lng, pend, dxp=DX, n_reg=T, seed=0, n_control=16)
    m = ModelsModule(ctx)
    synth.apply_to_models(m, inp)
    m.make_resident()
    ones = np.ones(n)
    n_cont = int(np.count_nonzero(inp["control"])) + 1

    def timed(fn):
        best = None
        for _ in range(a.reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = fn()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        return best, r

    t_idw, (mean, pos) = timed(lambda: m.rain_idw(xy, coord, rain, 2.0, 0, None, 0.0, ones, resident=True))
    wet = int((pos != 1).sum())
    base = {"cells": int(n), "steps": T, "stations": ns, "wet_steps": wet}
    print(json.dumps({**base, "what": "rain_idw, record table produced in HBM (one pass over the steps with a positive reading)", "ms": 1e3 * t_idw,
                      "g_cell_steps_per_s": n * T / t_idw / 1e9, "g_station_cell_steps_per_s": 2.0 * n * T * ns / t_idw / 1e9,
                      "table_gb": 4.0 * n * (wet + 1) / 1e9}), flush=True)
    t_run, ret = timed(lambda: m.shia_v1(None, None, inp["calib"], n_cont, 1, T, None, None, n))
    t_chain, _ = timed(lambda: (m.rain_idw(xy, coord, rain, 2.0, 0, None, 0.0, ones, resident=True),
                                m.shia_v1(None, None, inp["calib"], n_cont, 1, T, None, None, n)))
    print(json.dumps({**base, "what": "gauges -> hydrograph: rain_idw(resident) + shia_v1 on the device table", "ms": 1e3 * t_chain,
                      "shia_only_ms": 1e3 * t_run, "g_cell_steps_per_s": n * T / t_chain / 1e9, "q_peak": float(ret[0].max())}), flush=True)
    tin = np.asfortranarray((Delaunay(coord.T.astype(np.float64)).simplices.T + 1).astype(np.int32))
    from oracle import ref
    perte = None
    if ref.available():
        perte = ref.rain_pre_mit(xy, tin, coord)
    if perte is not None and perte.min() >= 1:
        t_mit, (mean2, pos2) = timed(lambda: m.rain_mit(xy, coord, rain, tin, perte, 0, None, 0.0, ones, resident=True))
        print(json.dumps({**base, "what": "rain_mit, table resident in HBM", "ms": 1e3 * t_mit, "triangles": int(tin.shape[1]),
                          "g_cell_steps_per_s": n * T / t_mit / 1e9}), flush=True)
    if ref.available():
        k = a.cpu_steps
        with tempfile.TemporaryDirectory() as d:
            t0 = time.perf_counter()
            r = ref.rain_idw(xy, coord, rain[:, :k], 2.0, 0.0, workdir=d)
            t_ref = time.perf_counter() - t0
        np.testing.assert_array_equal(r["pos_ids"], pos[:k])
        print(json.dumps({**base, "what": f"reference Fortran rain_idw (oracle/_ref), first {k} steps, 1 core", "ms": 1e3 * t_ref,
                          "g_cell_steps_per_s": n * k / t_ref / 1e9,
                          "gpu_speedup_per_step": (t_ref / k) / (t_idw / T)}), flush=True)


if __name__ == "__main__":
    main()
