#!/usr/bin/env python
"""C5 shape (BASELINE.json configs[4]): storm scenarios of the SED + landslide model on the ~2.1 M-cell, 12 m basin.  Scenarios
differ in rainfall and are independent runs; K of them share one GPU through K contexts (one CUDA stream each, driven by K
host threads -- ctypes releases the GIL), so that one scenario's half-empty wavefronts leave room for the others'.
Prints one JSON line per K: scenarios per second and cell-steps/s."""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c5")
    ap.add_argument("--concurrency", default="1,2,3,4")
    ap.add_argument("--per-thread", type=int, default=3, help="scenarios each thread runs in the timed region")
    ap.add_argument("--plain", action="store_true", help="plain hydrology instead of SED + slides")
    ap.add_argument("--member-axis", default="", help="comma list of scenario counts run as ONE launch sequence (scenario = member "
                                                      "axis, each with its own sequence of the shared rain records)")
    a = ap.parse_args()
    import torch
    import bench
    from wmf_b200 import Context, CuModule, ModelsModule, synth
    ctx0 = Context(0)
    wl = bench.build_workload(a.workload, CuModule(ctx0), verbose=True)
    inp, n, T = wl["inp"], wl["n"], wl["T"]
    if not a.plain:
        rng = np.random.default_rng(0)
        inp.update({"sim_sediments": 1, "sed_factor": 1.0, "wi": [0.036, 2.2e-4, 8.6e-7], "diametro": [0.35, 0.016, 0.001],
                    "krus": rng.uniform(0.1, 0.4, n).astype(np.float32)[None, :], "crus": rng.uniform(0.001, 0.3, n).astype(np.float32)[None, :],
                    "prus": np.ones((1, n), np.float32), "parliac": np.tile(np.array([[40.0], [40.0], [20.0]], np.float32), (1, n)),
                    "speed_type": [2, 2, 1], "sim_slides": 1, "sl_gullienogullie": 1, "sl_fs": 1.0, "sl_gammaw": 9.8,
                    "sl_zs": rng.uniform(0.2, 1.0, n).astype(np.float32)[None, :], "sl_gammas": np.full((1, n), 18.0, np.float32),
                    "sl_cohesion": rng.uniform(2, 10, n).astype(np.float32)[None, :],
                    "sl_frictionangle": np.deg2rad(rng.uniform(25, 35, n)).astype(np.float32)[None, :],
                    "sl_radslope": np.deg2rad(rng.uniform(20, 50, n)).astype(np.float32)[None, :]})
    n_cont = int(np.count_nonzero(inp["control"])) + 1
    calib = np.asarray(inp["calib"], np.float32)[None, :]
    want = {"q", "stoout"} | (set() if a.plain else {"qsed", "sl_slideocurrence"})
    if a.member_axis:
        from wmf_b200.ensemble import run_scenario_ensemble
        m = ModelsModule(ctx0)
        synth.apply_to_models(m, inp)
        m.make_resident()
        d_rain = torch.from_numpy(inp["rain"]).cuda()
        pos = np.asarray(inp["pos_evento"], np.int32)
        for S in [int(v) for v in a.member_axis.split(",")]:
            scen = np.stack([np.roll(pos, 11 * k) for k in range(S)]).astype(np.int32)    # the storm arriving 11 steps later each
            best = None
            for _ in range(3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                _, outs = run_scenario_ensemble(m, inp["calib"], scen, n_cont, T, d_rain, scenarios_per_launch=S, device_out=True, want=want)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            print(json.dumps({"workload": a.workload, "cells": int(n), "steps": T, "mode": "member axis, one launch sequence",
                              "sub_models": "plain" if a.plain else "SED + slides, speed_type=[2,2,1]", "scenarios": S, "seconds": best,
                              "ms_per_scenario": 1e3 * best / S, "g_cell_steps_per_s": n * T * S / best / 1e9,
                              "wave_ms": ctx0.last_kernel_ms(), "q_peak": float(outs[0]["q"].max())}), flush=True)
        return
    kmax = max(int(v) for v in a.concurrency.split(","))
    workers = []
    for k in range(kmax):       # one context, one resident copy of the maps and one scenario's rain per worker
        ctx = Context(0)
        m = ModelsModule(ctx)
        synth.apply_to_models(m, inp)
        m.make_resident()
        rain = np.roll(inp["rain"], 17 * k, axis=1) if k else inp["rain"]       # another storm: the field shifted over the basin
        workers.append((ctx, m, torch.from_numpy(np.ascontiguousarray(rain)).cuda()))
    def run(w, reps, out):
        ctx, m, d_rain = w
        for _ in range(reps):
            r = m._run(calib, n_cont, 1, T, d_rain, inp["pos_evento"], device_out=True, want=want)
        ctx.synchronize()
        out.append(float(r["q"].max()))
    for w in workers:
        run(w, 1, [])            # warm-up: pools, plans
    for K in [int(v) for v in a.concurrency.split(",")]:
        outs = []
        th = [threading.Thread(target=run, args=(workers[k], a.per_thread, outs)) for k in range(K)]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        S = K * a.per_thread
        print(json.dumps({"workload": a.workload, "cells": int(n), "steps": T, "sub_models": "plain" if a.plain else "SED + slides, speed_type=[2,2,1]",
                          "concurrent_scenarios": K, "scenarios": S, "seconds": dt, "ms_per_scenario": 1e3 * dt / S,
                          "scenarios_per_s": S / dt, "g_cell_steps_per_s": n * T * S / dt / 1e9, "q_peaks": outs[:K]}), flush=True)


if __name__ == "__main__":
    main()
