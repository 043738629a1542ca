#!/usr/bin/env python
"""Parity at BASELINE.json's FULL sizes against the reference's own compiled Fortran (oracle/_ref).

  --what c2        C2: SHIA on the 1 044 484-cell basin x --steps steps: the GPU run (host buffers through models.shia_v1)
                   vs the Fortran on one host core, same inputs (1000 steps take the Fortran about a minute)
  --what c4        C4 shape: an M-member calibration ensemble on a ~260k-cell basin in one GPU launch sequence; --check
                   randomly chosen members are re-run one by one by the Fortran and their hydrographs compared
Prints one JSON line with the worst relative deviations (north-star tolerance: 1e-5 relative, fp32)."""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def rel(got, want):
    got, want = np.asarray(got, np.float64).reshape(-1), np.asarray(want, np.float64).reshape(-1)
    return float(np.abs(got - want).max() / max(np.abs(want).max(), 1e-30))


def excess(got, want, rtol=1e-5, afloor=1e-7):
    """The tests' criterion |got-want| <= rtol*|want| + afloor*max|want| as one number: the worst error in units of its
    tolerance (<= 1 passes)."""
    got, want = np.asarray(got, np.float64).reshape(-1), np.asarray(want, np.float64).reshape(-1)
    return float((np.abs(got - want) / (rtol * np.abs(want) + afloor * np.abs(want).max() + 1e-300)).max())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="c2", choices=["c2", "c4", "c5"])
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--members", type=int, default=256)
    ap.add_argument("--check", type=int, default=6)
    a = ap.parse_args()
    import torch
    import bench
    from oracle import ref
    from wmf_b200 import Context, CuModule, ModelsModule, synth
    assert ref.available(), "oracle/_ref/libwmf_ref.so is missing"
    ctx = Context(0)
    cu, m = CuModule(ctx), ModelsModule(ctx)
    if a.what == "c2":
        wl = bench.build_workload("c2", cu)
        inp, n, T = wl["inp"], wl["n"], min(a.steps, wl["T"])
        inp = dict(inp, n_reg=T, pos_evento=np.ascontiguousarray(inp["pos_evento"][:T]), evpserie=inp["evpserie"][:T])
        synth.apply_to_models(m, inp)
        m.set_rain(inp["rain"], inp["pos_evento"])
        n_cont = int(np.count_nonzero(inp["control"])) + 1
        t0 = time.perf_counter()
        ret = m.shia_v1(None, None, inp["calib"], n_cont, 1, T, None, None, n)
        t_gpu = time.perf_counter() - t0
        with tempfile.TemporaryDirectory() as d:
            files = (os.path.join(d, "rain.bin"), os.path.join(d, "rain.hdr"))
            ref.write_rain(files[0], files[1], inp["rain"], inp["pos_evento"])
            t0 = time.perf_counter()
            R = ref.shia_v1(inp, rain_files=files)
            t_ref = time.perf_counter() - t0
        line = {"what": f"C2 full size: {n} cells x {T} steps, GPU (models.shia_v1, host buffers) vs the reference's compiled Fortran",
                "cells": int(n), "steps": T, "gpu_s": t_gpu, "reference_fortran_s_1core": t_ref, "speedup": t_ref / t_gpu,
                "q_outlet_max_rel": rel(ret[0][0], R["q"][0]), "q_all_points_max_rel": rel(ret[0], R["q"]),
                "stoout_max_rel": rel(ret[9], R["stoout"]),
                "q_worst_error_in_units_of_tolerance": excess(ret[0], R["q"]),
                "stoout_worst_error_in_units_of_tolerance": excess(ret[9], R["stoout"]),
                "mean_rain_max_rel": rel(m.mean_rain, R["mean_rain"]),
                "acum_rain_bit_exact": bool(np.array_equal(np.asarray(m.acum_rain).reshape(-1), R["acum_rain"])),
                "q_peak_m3s": float(R["q"][0].max())}
        print(json.dumps(line), flush=True)
        return
    if a.what == "c5":
        # C5: SED sediments + landslides on the ~2.1 M-cell, 12 m basin, 288 five-minute steps, kinematic hillslope speeds
        wl = bench.build_workload("c5", cu)
        inp, n, T = wl["inp"], wl["n"], min(a.steps, wl["T"])
        rng = np.random.default_rng(0)
        inp = dict(inp, n_reg=T, pos_evento=np.ascontiguousarray(inp["pos_evento"][:T]), evpserie=inp["evpserie"][:T])
        inp.update({"sim_sediments": 1, "sed_factor": 1.0, "wi": [0.036, 2.2e-4, 8.6e-7], "diametro": [0.35, 0.016, 0.001],
                    "krus": rng.uniform(0.1, 0.4, n).astype(np.float32)[None, :], "crus": rng.uniform(0.001, 0.3, n).astype(np.float32)[None, :],
                    "prus": np.ones((1, n), np.float32), "parliac": np.asfortranarray(np.tile(np.array([[40.0], [40.0], [20.0]], np.float32), (1, n))),
                    "speed_type": [2, 2, 1], "sim_slides": 1, "sl_gullienogullie": 1, "sl_fs": 1.0, "sl_gammaw": 9.8,
                    "sl_zs": rng.uniform(0.2, 1.0, n).astype(np.float32)[None, :], "sl_gammas": np.full((1, n), 18.0, np.float32),
                    "sl_cohesion": rng.uniform(2, 10, n).astype(np.float32)[None, :],
                    "sl_frictionangle": np.deg2rad(rng.uniform(25, 35, n)).astype(np.float32)[None, :],
                    "sl_radslope": np.deg2rad(rng.uniform(20, 50, n)).astype(np.float32)[None, :]})
        synth.apply_to_models(m, inp)
        m.set_rain(inp["rain"], inp["pos_evento"])
        n_cont = int(np.count_nonzero(inp["control"])) + 1
        m.shia_v1(None, None, inp["calib"], n_cont, 1, T, None, None, n)          # warm-up: pools, plans
        m.vd = None        # Vd carries over from one run to the next (sed_allocate, modelosv2.f90:1301-1304): start fresh like the Fortran below
        t0 = time.perf_counter()
        ret = m.shia_v1(None, None, inp["calib"], n_cont, 1, T, None, None, n)
        t_gpu = time.perf_counter() - t0
        with tempfile.TemporaryDirectory() as d:
            files = (os.path.join(d, "rain.bin"), os.path.join(d, "rain.hdr"))
            ref.write_rain(files[0], files[1], inp["rain"], inp["pos_evento"])
            t0 = time.perf_counter()
            R = ref.shia_v1(inp, rain_files=files)
            t_ref = time.perf_counter() - t0
        occ_g, occ_r = np.asarray(m.sl_slideocurrence).reshape(-1), R["sl_slideocurrence"]
        line = {"what": f"C5 full size: SED + slides, {n} cells x {T} steps, GPU (models.shia_v1, host buffers) vs the reference's compiled Fortran",
                "cells": int(n), "steps": T, "gpu_s": t_gpu, "reference_fortran_s_1core": t_ref, "speedup": t_ref / t_gpu,
                "q_worst_error_in_units_of_tolerance_1e-5": excess(ret[0], R["q"]),
                "stoout_worst_error_in_units_of_tolerance_1e-5": excess(ret[9], R["stoout"]),
                "qsed_worst_error_in_units_of_tolerance_5e-5": excess(ret[1], R["qsed"], rtol=5e-5),
                "vd_worst_error_in_units_of_tolerance_5e-5": excess(m.vd, R["vd"], rtol=5e-5),
                "voldepo_worst_error_in_units_of_tolerance_5e-5": excess(m.voldepo, R["voldepo"], rtol=5e-5),
                "slide_cells_reference": int(np.count_nonzero(occ_r)), "slide_cells_differing": int(np.count_nonzero(occ_g != occ_r)),
                "riskvector_bit_exact": bool(np.array_equal(np.asarray(m.sl_riskvector).reshape(-1), R["sl_riskvector"]))}
        print(json.dumps(line), flush=True)
        return
    # C4 shape
    nc = nr = 512
    DIR, DEM = synth.make_raster("G2", nc, nr, seed=0)
    cu.ncols, cu.nrows, cu.xll, cu.yll, cu.dx, cu.dy, cu.dxp, cu.nodata = nc, nr, 0.0, 0.0, 30.0, 30.0, 30.0, synth.NODATA
    x, y = synth.outlet_xy(nc, nr, 0.0, 0.0, 30.0)
    n = cu.basin_find(x, y, DIR, nc, nr)
    b = cu.basin_cut(n)
    dv = cu.basin_map2basin(b, DIR.astype(np.float32), 0.0, 0.0, 30.0, 30.0, synth.NODATA).astype(np.int32)
    de = cu.basin_map2basin(b, DEM, 0.0, 0.0, 30.0, 30.0, synth.NODATA)
    acum, lng, pend, _ = cu.basin_basics(b, de, dv)
    T, M = min(a.steps, 200), a.members
    inp = synth.make_shia_inputs(b, acum, lng, pend, dxp=30.0, dt=3600.0, n_reg=T, seed=0, n_control=16)
    synth.apply_to_models(m, inp)
    m.make_resident()
    d_rain = torch.from_numpy(inp["rain"]).cuda()
    n_cont = int(np.count_nonzero(inp["control"])) + 1
    calibs = synth.make_ensemble_calibs(M, seed=0)
    out = m._run(calibs, n_cont, 1, T, d_rain, inp["pos_evento"], device_out=True, want={"q"})
    q = out["q"].cpu().numpy().reshape(M, T, n_cont)              # (member, step, control row)
    rng = np.random.default_rng(0)
    picks = sorted(rng.choice(M, size=min(a.check, M), replace=False).tolist())
    worst, worst_pt = 0.0, 0.0
    with tempfile.TemporaryDirectory() as d:
        files = (os.path.join(d, "rain.bin"), os.path.join(d, "rain.hdr"))
        ref.write_rain(files[0], files[1], inp["rain"], inp["pos_evento"])
        for k in picks:
            R = ref.shia_v1(dict(inp, calib=calibs[k]), rain_files=files)
            worst = max(worst, rel(q[k].T, R["q"]))
            worst_pt = max(worst_pt, excess(q[k].T, R["q"]))
    print(json.dumps({"what": f"C4 shape: {M}-member ensemble on a {n}-cell basin x {T} steps in one launch sequence; members {picks} "
                              "re-run by the reference's compiled Fortran", "cells": int(n), "steps": T, "members": M,
                      "members_checked": picks, "q_max_rel_worst_member": worst, "q_worst_error_in_units_of_tolerance": worst_pt}), flush=True)


if __name__ == "__main__":
    main()
