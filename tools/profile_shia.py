#!/usr/bin/env python
"""Profiling driver: one warm-up + one resident SHIA run of a bench workload (used under ncu; see profiles/README.md)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--runs", type=int, default=2)
    ap.add_argument("--members", type=int, default=1)
    ap.add_argument("--sed", action="store_true", help="enable the SED sediment sub-model (kinematic hillslope speeds)")
    ap.add_argument("--slides", action="store_true", help="enable the landslide sub-model")
    ap.add_argument("--kin", action="store_true", help="kinematic hillslope speeds speed_type=[2,2,1] without sub-models")
    ap.add_argument("--steps", type=int, default=0, help="override the number of steps")
    ap.add_argument("--no-sed", action="store_true", help="switch the workload's sediments off (c5)")
    ap.add_argument("--no-slides", action="store_true", help="switch the workload's landslides off (c5)")
    ap.add_argument("--linear", action="store_true", help="linear hillslope speeds speed_type=[1,1,1] (needs --no-sed)")
    ap.add_argument("--no-gullie", action="store_true", help="landslides without the downstream marking (sl_gullienogullie = 0)")
    a = ap.parse_args()
    import torch
    import bench
    from wmf_b200 import Context, CuModule, ModelsModule, synth
    ctx = Context(0)
    cu, m = CuModule(ctx), ModelsModule(ctx)
    wl = bench.build_workload(a.workload, cu, verbose=True)
    inp = wl["inp"]
    if a.steps:
        wl["T"] = a.steps
    if a.sed:
        rng = np.random.default_rng(0)
        n = wl["n"]
        inp.update({"sim_sediments": 1, "sed_factor": 1.0, "wi": [0.036, 2.2e-4, 8.6e-7], "diametro": [0.35, 0.016, 0.001],
                    "krus": rng.uniform(0.1, 0.4, n).astype(np.float32)[None, :], "crus": rng.uniform(0.001, 0.3, n).astype(np.float32)[None, :],
                    "prus": np.ones((1, n), np.float32), "parliac": np.tile(np.array([[40.0], [40.0], [20.0]], np.float32), (1, n)),
                    "speed_type": [2, 2, 1]})
    if a.kin:
        inp["speed_type"] = [2, 2, 1]
    if a.slides:
        rng = np.random.default_rng(1)
        n = wl["n"]
        inp.update({"sim_slides": 1, "sl_gullienogullie": 1, "sl_fs": 1.0, "sl_gammaw": 9.8,
                    "sl_zs": rng.uniform(0.2, 1.0, n).astype(np.float32)[None, :], "sl_gammas": np.full((1, n), 18.0, np.float32),
                    "sl_cohesion": rng.uniform(2, 10, n).astype(np.float32)[None, :],
                    "sl_frictionangle": np.deg2rad(rng.uniform(25, 35, n)).astype(np.float32)[None, :],
                    "sl_radslope": np.deg2rad(rng.uniform(20, 50, n)).astype(np.float32)[None, :]})
    if a.no_sed:
        inp["sim_sediments"] = 0
    if a.no_slides:
        inp["sim_slides"] = 0
    if a.linear:
        inp["speed_type"] = [1, 1, 1]
    if a.no_gullie:
        inp["sl_gullienogullie"] = 0
    a.sed = a.sed or int(inp.get("sim_sediments", 0)) == 1
    a.slides = a.slides or int(inp.get("sim_slides", 0)) == 1
    synth.apply_to_models(m, inp)
    m.make_resident()
    d_rain = torch.from_numpy(inp["rain"]).cuda()
    calibs = synth.make_ensemble_calibs(max(a.members, 2), 0)[:a.members]
    n_cont = int(np.count_nonzero(inp["control"])) + 1
    for _ in range(a.runs):
        out = m._run(calibs, n_cont, 1, wl["T"], d_rain, inp["pos_evento"], device_out=True, want={"q", "stoout", "balance"} | ({"qsed", "vs"} if a.sed else set()) | ({"sl_slideocurrence"} if a.slides else set()))
        print(f"run: {ctx.last_kernel_ms():.2f} ms wave loop, levels/waves {ctx.last_shia_stats()}, "
              f"{wl['n'] * wl['T'] * a.members / ctx.last_kernel_ms() / 1e6:.2f} G cell-steps/s", flush=True)
    print("qmax", float(out["q"].max()))


if __name__ == "__main__":
    main()
