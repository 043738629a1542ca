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
    a = ap.parse_args()
    import torch
    import bench
    from wmf_b200 import Context, CuModule, ModelsModule, synth
    ctx = Context(0)
    cu, m = CuModule(ctx), ModelsModule(ctx)
    wl = bench.build_workload(a.workload, cu, verbose=True)
    inp = wl["inp"]
    synth.apply_to_models(m, inp)
    m.make_resident()
    d_rain = torch.from_numpy(inp["rain"]).cuda()
    calibs = synth.make_ensemble_calibs(max(a.members, 2), 0)[:a.members]
    n_cont = int(np.count_nonzero(inp["control"])) + 1
    for _ in range(a.runs):
        out = m._run(calibs, n_cont, 1, wl["T"], d_rain, inp["pos_evento"], device_out=True, want={"q", "stoout", "balance"})
        print(f"run: {ctx.last_kernel_ms():.2f} ms wave loop, levels/waves {ctx.last_shia_stats()}, "
              f"{wl['n'] * wl['T'] * a.members / ctx.last_kernel_ms() / 1e6:.2f} G cell-steps/s", flush=True)
    print("qmax", float(out["q"].max()))


if __name__ == "__main__":
    main()
