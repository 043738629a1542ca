#!/usr/bin/env python
"""Profiling driver for Path A (used under ncu; see profiles/README.md): basin_find + basin_cut + basin_acum,
basin_basics and geo_hand on a G2 raster resident in HBM."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=4096)
    ap.add_argument("--runs", type=int, default=2)
    ap.add_argument("--basics", action="store_true", help="also basin_basics and geo_hand")
    a = ap.parse_args()
    import torch
    from wmf_b200 import Context, CuModule, synth
    ctx = Context(0)
    cu = CuModule(ctx)
    nc = nr = a.size
    d_dir = synth.make_dir_torch("G2", nc, nr, 0, device="cuda:0")
    cu.ncols, cu.nrows, cu.xll, cu.yll, cu.dx, cu.dy, cu.dxp, cu.nodata = nc, nr, 0.0, 0.0, 30.0, 30.0, 30.0, synth.NODATA
    x, y = synth.outlet_xy(nc, nr, 0.0, 0.0, 30.0)
    for _ in range(a.runs):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = cu.basin_find(x, y, d_dir)
        t1 = time.perf_counter()
        b = cu.basin_cut(n, resident=True)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        acum = cu.basin_acum(b)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        print(f"find {1e3 * (t1 - t0):.2f} ms  cut {1e3 * (t2 - t1):.2f} ms  acum {1e3 * (t3 - t2):.2f} ms  "
              f"{n / (t3 - t0) / 1e6:.1f} Mcell/s  ({n} cells, {cu.basin_find_depth()} levels)", flush=True)
    if a.basics:
        lin = (b[:, 2].long() - 1) * nc + (b[:, 1].long() - 1)
        dirv = d_dir.reshape(-1)[lin].contiguous()
        lev = torch.zeros(n, dtype=torch.float32, device="cuda:0")
        demv = (10.0 + 1e-3 * (n - torch.arange(n, device="cuda:0", dtype=torch.float32))).contiguous()   # decreasing towards the outlet
        for _ in range(a.runs):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ac, lng, pend, elev = cu.basin_basics(b, demv, dirv)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            cauce = (ac >= 500).to(torch.int32)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            hand, hdnd, aq = cu.geo_hand(b, elev, lng, cauce)
            torch.cuda.synchronize()
            t3 = time.perf_counter()
            print(f"basics {1e3 * (t1 - t0):.2f} ms  geo_hand {1e3 * (t3 - t2):.2f} ms", flush=True)


if __name__ == "__main__":
    main()
