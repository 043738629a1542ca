#!/usr/bin/env python
"""C4 shape across GPUs: a population of P SHIA parameter sets on one ~250k-cell basin, sharded over the ranks of a
torchrun launch (wmf_b200.ensemble.evaluate_population: no data-path collective, one NCCL all-gather of the scores).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_ensemble_dist.py
"""
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
    ap.add_argument("--population", type=int, default=512)
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--reps", type=int, default=2)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    from wmf_b200 import Context, CuModule, ModelsModule, synth
    from wmf_b200.ensemble import evaluate_population
    ctx = Context(local)
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
    calibs = synth.make_ensemble_calibs(a.population, seed=0)
    # "observations" = member 0's hydrograph at the outlet
    q0 = m._run(calibs[:1], n_cont, 1, T, d_rain, inp["pos_evento"], want={"q"})["q"][0]
    best = None
    for _ in range(a.reps + 1):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sc = evaluate_population(m, calibs, q0, n_cont, T, d_rain, inp["pos_evento"], row=0, members_per_launch=a.batch)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        best = float(dt) if best is None else min(best, float(dt))
    if rank == 0:
        sc = sc.cpu().numpy()
        print(json.dumps({"n_gpus": world, "cells": int(n), "steps": T, "population": a.population, "batch": a.batch,
                          "seconds": best, "member_cell_steps_per_s": float(n) * T * a.population / best,
                          "nash_member0": float(sc[0, 0]), "nash_min": float(np.nanmin(sc[:, 0])), "scores_shape": list(sc.shape)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
