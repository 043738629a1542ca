#!/usr/bin/env python
"""bench.py -- the reference's headline metric on B200: "Mcell/s D8 flow-acc; SHIA cell-steps/s (ensemble) at 1/2/4/8 B200
vs host CPU" (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W             # our CUDA path, one rank per GPU (torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU algorithm on the host cores

Headline workload (every N): **C4**, BASELINE.json configs[3] -- an NSGA-II calibration population on the ~250k-cell basin
(512 x 512 G2 raster, 260 100 cells), 1000 hourly steps, 512 parameter sets PER GPU (weak scaling: 8 GPUs = the 4096-member
target), evaluated through the public multi-GPU call `wmf_b200.ensemble.evaluate_population` (the replacement of
wmf.py:872-878 `__ejec_parallel__` + `__eval_nash__`/`__eval_q_pico__`): members shard over the ranks with no data-path
collective, the per-member scores are all-gathered over NCCL.  A "step" is one evaluation of the whole population.

Numbers on the JSON line
  value      member-cell-steps/s = ranks x members x cells x steps / max-over-ranks time; rain and maps resident in HBM;
             timed with CUDA events on the launch stream
  e2e        the same call with HOST inputs (pinned rain table, numpy maps): every rank uploads 1/N of the rain records
             and the ranks all-gather the rest over NVLink; scores come back to the host.  Copies inside the timed region.
  roofline   dominant kernel k_shia_tb<K,...> (time-blocked wavefront, member-minor layout): algorithmic bytes
             (80 B per member-cell-step, SURVEY 8d / DESIGN.md) / CUDA-event time of the wave launches, against the
             measured HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline  the reference's own compiled Fortran (oracle/_ref, kind "reference"; the oracle's C port when that library
             is absent) running whole members of the same population on ONE core; the hydrographs it produces are compared
             with the GPU's inside the bench (`extra.parity_vs_cpu`)
  extra      c2: the single-basin run of BASELINE configs[1] (1 044 484 cells x 1000 steps, 144 B/cell-step roofline, its
             own e2e through models.shia_v1 with host buffers) -- N=1 only;
             c5: BASELINE configs[4], SED sediments + landslides on the 2.09 M-cell basin: one storm, and 8 storm scenarios as
             one launch sequence, with the reference's Fortran on a few steps -- N=1 only;
             path_a: "Mcell/s D8 flow-acc" on C1 (512 x 512) and C3 (20000 x 20000): basin_find + basin_cut + basin_acum,
             per-call ms, levels, roofline at 28 B/cell, cpu_baseline of the reference on the same raster, bit-exactness
`--workload c2 | c5` make the single-basin run (C5: SED sediments + landslides) the headline instead.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic bytes
BYTES_PER_CELL_STEP = 144.0            # one cell-step of a single run
BYTES_PER_MEMBER_CELL_STEP = 80.0      # one member-cell-step of an ensemble (statics amortised over the members of a cell)
BYTES_SED, BYTES_SLIDES = 120.0, 28.0  # what the sediment / landslide sub-models add per cell-step
BYTES_PER_CELL_PATH_A = 28.0           # basin_find + basin_cut 16 B, basin_acum 12 B per basin cell
# what the tile path really moves: sum of dram__bytes_read + dram__bytes_write over every kernel of one basin_find +
# basin_cut + basin_acum at 8192 x 8192 (67.1 M cells), ncu, profiles/dram_path_a_8192_r2.txt
TRAFFIC_PER_CELL_PATH_A = 160.6
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the ncu --set full captures under
# profiles/ (cold-cache replay of one mid-run launch); None until a capture exists
TRAFFIC = {
    "single": (60.0e6, "ncu --set full, one mid-run launch of k_shia_tb<4,...> on C2 = 1.33e6 cell-steps (cold-cache replay): "
                       "56.5 MB read + 3.5 MB written = 45 B per cell-step against 144 algorithmic; profiles/README.md"),
    # per member-cell-step; scaled to the average launch of the timed run in run_ours()
    "ensemble": (31.8, "ncu --set full, one mid-run launch of k_shia_tb<16,...> (C4 shape, 256 members, 47636 CTAs x 256 threads x 16 "
                       "steps = 1.95e8 member-cell-steps, cold-cache replay): 3.11 GB read + 3.09 GB written = 31.8 B per "
                       "member-cell-step against 80 algorithmic (the outflow ring is produced and consumed through L2); this "
                       "figure x the member-cell-steps of an average launch; profiles/ncu_k_shia_tb_ensemble_r2_raw.txt"),
    "c5": (None, "ncu captures of k_shia_tb<4,KIN,SED,SLIDES> are under profiles/ (ncu_k_shia_tb_c5_*_r2_raw.txt): a mid-run "
                 "launch of 415 CTAs moved 58 MB; not scaled to the average launch"),
}
WORKLOADS = {
    # BASELINE.json configs[3]: calibration ensemble on the ~250k-cell basin
    "c4": dict(nc=512, nr=512, T=1000, kind="ensemble", members=512, batch=256),
    "c4-small": dict(nc=130, nr=130, T=120, kind="ensemble", members=16, batch=8),
    # BASELINE.json configs[1]: one basin, one run
    "c2": dict(nc=1024, nr=1024, T=1000, kind="single"),
    "c2-small": dict(nc=258, nr=258, T=200, kind="single"),
    # BASELINE.json configs[4] basin: ~2.1 M cells, 12 m, 5-minute steps, kinematic hillslopes, sediments + landslides
    "c5": dict(nc=1448, nr=1448, T=288, kind="single", dx=12.0, dt=300.0, thresholds=(200, 2000), speed_type=(2, 2, 1),
               sediments=True, slides=True),
}
DX = 30.0


def wl_cfg(name):
    c = dict(gen="G2", dx=DX, dt=3600.0, thresholds=(30, 500), speed_type=(1, 1, 1), sediments=False, slides=False)
    c.update(WORKLOADS[name])
    return c


def peaks():
    """(HBM GB/s, how) -- MEASURED_PEAKS.json when the driver wrote it, else the profiling guide's fallback."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (torch copy, read+write bytes)"
        except Exception:
            pass
    return 6300.0, "fallback: B200_PROFILING.md measured copy bandwidth"


class ClockSampler:
    """nvidia-smi SM clock + throttle reasons sampled during the timed region (the profiling recipe's clocks line)."""

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.th = index, [], None, None

    def start(self):
        if os.environ.get("BENCH_CLOCKS_MS") == "0":          # (A/B of the sampler's own cost; the bench line then says so)
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm,clocks_throttle_reasons.hw_slowdown,"
                 "clocks_throttle_reasons.hw_thermal_slowdown,clocks_throttle_reasons.sw_thermal_slowdown,"
                 "clocks_throttle_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", os.environ.get("BENCH_CLOCKS_MS", "200")],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.th = threading.Thread(target=self._read, daemon=True)
        self.th.start()

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([x.strip() for x in ln.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        mhz, mx = [], []
        for r in self.rows:
            try:
                mhz.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                pass
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 6 for k in range(4) if r[2 + k].lower().startswith("active")})
        return {"sm_mhz": float(np.median(mhz)) if mhz else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(mhz)}


def trace_basin(cu, DIR, DEM, nc, nr, dx):
    """basin_find -> basin_cut -> map2basin -> basin_basics with the module object ``cu`` (ours, or the CPU reference's)."""
    from wmf_b200 import synth
    cu.ncols, cu.nrows, cu.xll, cu.yll, cu.dx, cu.dy, cu.dxp, cu.nodata = nc, nr, 0.0, 0.0, dx, dx, dx, synth.NODATA
    x, y = synth.outlet_xy(nc, nr, 0.0, 0.0, dx)
    n = cu.basin_find(x, y, DIR, nc, nr) if hasattr(cu, "ctx") else cu.basin_find(x, y, DIR)
    b = cu.basin_cut(n)
    dv = cu.basin_map2basin(b, DIR.astype(np.float32), 0.0, 0.0, dx, dx, synth.NODATA).astype(np.int32)
    de = cu.basin_map2basin(b, DEM, 0.0, 0.0, dx, dx, synth.NODATA)
    acum, lng, pend, _ = cu.basin_basics(b, de, dv)
    return x, y, n, b, acum, lng, pend


def make_inputs(cfg, b, acum, lng, pend, T=None):
    from wmf_b200 import synth
    return synth.make_shia_inputs(b, acum, lng, pend, dxp=cfg["dx"], dt=cfg["dt"], n_reg=T or cfg["T"], seed=0,
                                  thresholds=cfg["thresholds"], speed_type=cfg["speed_type"], n_control=16,
                                  sediments=cfg["sediments"], slides=cfg["slides"])


def build_workload(name: str, cu, rank: int = 0, verbose: bool = False):
    """Synthetic basin + SHIA inputs of workload ``name``; the topology is traced with OUR Path-A kernels."""
    from wmf_b200 import synth
    cfg = wl_cfg(name)
    nc, nr, T = cfg["nc"], cfg["nr"], cfg["T"]
    t0 = time.time()
    DIR, DEM = synth.make_raster(cfg["gen"], nc, nr, seed=0)
    x, y, n, b, acum, lng, pend = trace_basin(cu, DIR, DEM, nc, nr, cfg["dx"])
    inp = make_inputs(cfg, b, acum, lng, pend)
    if verbose:
        print(f"[bench] workload {name}: {n} cells, {T} steps, {inp['rain'].shape[0]} rain records, "
              f"built in {time.time() - t0:.1f}s", file=sys.stderr)
    return dict(name=name, cfg=cfg, DIR=DIR, DEM=DEM, x=x, y=y, n=n, basin=b, inp=inp, T=T, nc=nc, nr=nr)


# ---------------------------------------------------------------------------------------------------------------------
# Path A: "Mcell/s D8 flow-acc"
# ---------------------------------------------------------------------------------------------------------------------
def bench_path_a(cu, nc, nr, reps=3, cpu=True, gen="G2", seed=0, hand=False):
    """basin_find + basin_cut + basin_acum on a ``nc x nr`` raster resident in HBM: Mcell/s, per-call ms, roofline at
    28 B/cell, and -- ``cpu`` -- the reference's CPU path on the same raster with a bit-exactness check of both tables."""
    import torch
    from wmf_b200 import synth
    dev = f"cuda:{cu.ctx.device}"
    d_dir = synth.make_dir_torch(gen, nc, nr, seed, device=dev)
    cu.ncols, cu.nrows, cu.xll, cu.yll, cu.dx, cu.dy, cu.dxp, cu.nodata = nc, nr, 0.0, 0.0, DX, DX, DX, synth.NODATA
    x, y = synth.outlet_xy(nc, nr, 0.0, 0.0, DX)
    best = None
    for _ in range(reps + 1):                      # first pass warms the memory pools
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
        ts = (t1 - t0, t2 - t1, t3 - t2)
        if best is None or sum(ts) < sum(best):
            best = ts
    assert int(acum[-1].item()) == n
    tot = sum(best)
    peak, peak_kind = peaks()
    ach = BYTES_PER_CELL_PATH_A * n / tot / 1e9
    res = {"raster": f"{nc}x{nr} {gen}", "cells": int(n), "levels": int(cu.basin_find_depth()),
           "mcell_per_s": n / tot / 1e6, "ms": tot * 1e3,
           "ms_find": best[0] * 1e3, "ms_cut": best[1] * 1e3, "ms_acum": best[2] * 1e3,
           "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                        "bytes_per_cell": BYTES_PER_CELL_PATH_A, "peak_kind": peak_kind, "traffic": TRAFFIC_PER_CELL_PATH_A * n,
                        "traffic_note": "160.6 B of DRAM traffic per basin cell (ncu, sum over the 13 passes of the tile path at "
                                        "8192 x 8192, profiles/dram_path_a_8192_r2.txt) x the cells of this raster: the BFS order "
                                        "without the level chain costs two radix passes, the preorder scatter and the tile tables"},
           "timing": "host clock around the three synchronous C-ABI calls, DIR resident in HBM, best of %d" % reps}
    if hand:
        # BASELINE configs[2] also names HAND: raster-wide height above the nearest drainage (geo_hand_global,
        # cuencas.f90:2175-2219) with the network = accumulation > 1000 cells; DEM = 10 + 0.5 x Chebyshev distance to the
        # outlet, which falls along every G2 flow path.  Informational (not part of the 28 B/cell metric); never fatal.
        try:
            oc, orow = synth.outlet_colrow(nc, nr)
            rr = (torch.arange(nr, device=dev, dtype=torch.int32) - orow).abs()[:, None]
            cc = (torch.arange(nc, device=dev, dtype=torch.int32) - oc).abs()[None, :]
            d_dem = (10.0 + 0.5 * torch.maximum(rr, cc).float()).contiguous()
            del rr, cc

            def timed(fn):                             # best of two: the first call sizes the context's memory pool
                best_t, r = None, None
                for _ in range(2):
                    del r
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    r = fn()
                    torch.cuda.synchronize()
                    dt = (time.perf_counter() - t0) * 1e3
                    best_t = dt if best_t is None else min(best_t, dt)
                return best_t, r
            t_v2m, acum_map = timed(lambda: cu.basin_float_var2map(b, acum.float(), nc, nr))
            t_net, red = timed(lambda: cu.geo_acum_to_cauce(acum_map.int(), 1000))
            del acum_map
            t_hg, hand_g = timed(lambda: cu.geo_hand_global(d_dem, d_dir, red))
            res["hand"] = {"ms_var2map": t_v2m, "ms_acum_to_cauce": t_net, "ms_geo_hand_global": t_hg,
                           "network_cells": int(red.sum().item()), "hand_max_m": float(hand_g.max().item()),
                           "gbs_16B_per_raster_cell": 16.0 * nc * nr / (t_hg / 1e3) / 1e9,
                           "what": "acum back on the raster, network mask acum > 1000 cells, every other cell walks down DIR to the network"}
            del d_dem, red, hand_g
        except Exception as e:                         # keep the bench line
            res["hand"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    if cpu:
        from oracle import ref as _ref
        import oracle
        kind = cpu_kind()
        ocu = _ref.RefCu() if kind == "reference" else oracle.Cu(fast=True)
        ocu.ncols, ocu.nrows, ocu.xll, ocu.yll, ocu.dx, ocu.dy, ocu.dxp, ocu.nodata = nc, nr, 0.0, 0.0, DX, DX, DX, synth.NODATA
        h_dir = d_dir.cpu().numpy().T                  # (ncols, nrows) F-ordered view, as wmf.py passes it
        t0 = time.perf_counter()
        n_c = ocu.basin_find(x, y, h_dir)
        b_c = ocu.basin_cut(n_c)
        a_c = ocu.basin_acum(b_c)
        t_cpu = time.perf_counter() - t0
        same = (n_c == n)
        if same:                                       # compared on the device, in slabs
            step = 1 << 26
            bc, ac = np.ascontiguousarray(np.asarray(b_c).T), np.asarray(a_c).reshape(-1)
            for s in range(0, n, step):
                e = min(n, s + step)
                same = same and bool(torch.equal(b[s:e], torch.from_numpy(bc[s:e]).to(dev)))
                same = same and bool(torch.equal(acum[s:e], torch.from_numpy(np.ascontiguousarray(ac[s:e])).to(dev)))
        res["cpu_baseline"] = {"value": n_c / t_cpu / 1e6, "unit": "Mcell/s", "cores": 1, "kind": kind,
                               "sample": f"the whole {nc}x{nr} raster, find+cut+acum in {t_cpu:.2f} s, single thread as the reference runs"}
        res["bit_exact_vs_cpu"] = bool(same)
        res["speedup_vs_cpu"] = (n / tot) / (n_c / t_cpu)
    del d_dir, b, acum
    torch.cuda.empty_cache()
    return res


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm
# ---------------------------------------------------------------------------------------------------------------------
def cpu_kind():
    """"reference" when the reference's own compiled Fortran is loadable (oracle/_ref/libwmf_ref.so, built from the
    objects in the reference tree and shipped to the GPU box), else "port" (the C restatement in oracle/)."""
    from oracle import ref
    return "reference" if ref.available() else "port"


def cpu_prepare(wl, sample_steps, workdir, calib=None):
    """Inputs of a CPU sample: the first ``sample_steps`` steps of the workload; for the reference arm the rain table is
    written to the .bin / .hdr pair the Fortran reads record by record (modelosv2.f90:444-449), outside the timed call."""
    inp = dict(wl["inp"])
    inp["n_reg"] = sample_steps
    pos = np.ascontiguousarray(inp["pos_evento"][:sample_steps])
    inp["evpserie"] = inp["evpserie"][:sample_steps]
    if calib is not None:
        inp["calib"] = np.asarray(calib, np.float32)
    # keep only the rain records the sample touches (record 1 = zeros stays first)
    keep = np.unique(np.concatenate([[1], pos]))
    remap = np.zeros(int(inp["rain"].shape[0]) + 1, np.int32)
    remap[keep] = np.arange(1, keep.size + 1)
    inp["rain"] = np.ascontiguousarray(inp["rain"][keep - 1])
    inp["pos_evento"] = remap[pos]
    files = None
    if cpu_kind() == "reference":
        from oracle import ref
        files = (os.path.join(workdir, "rain.bin"), os.path.join(workdir, "rain.hdr"))
        if not os.path.exists(files[0]):
            ref.write_rain(files[0], files[1], inp["rain"], inp["pos_evento"])
    return inp, files


def cpu_run(wl, sample_steps, workdir, calib=None):
    """(cell-steps/s, seconds, result dict) of the reference's CPU path on the first ``sample_steps`` steps, 1 thread."""
    inp, files = cpu_prepare(wl, sample_steps, workdir, calib)
    t0 = time.perf_counter()
    if files is not None:
        from oracle import ref
        out = ref.shia_v1(inp, rain_files=files)
    else:
        import oracle
        out = oracle.shia_v1(inp, fast=True)
    dt = time.perf_counter() - t0
    return wl["n"] * sample_steps / dt, dt, out


def cpu_sample(wl, sample_steps, workdir):
    v, dt, _ = cpu_run(wl, sample_steps, workdir)
    return v, dt


def cpu_label():
    return ("the reference's own compiled Fortran (gfortran 5.3.0 -O3 -funroll-loops objects from the reference tree, "
            "oracle/_ref), rain read from its .bin file every step" if cpu_kind() == "reference"
            else "oracle port -O3 -funroll-loops, rain in memory")


def _ref_worker(args):
    sample_steps, workdir, k = args
    os.environ["CUDA_VISIBLE_DEVICES"] = ""
    import oracle  # noqa: F401
    wl = _REF_WL[0]
    cal = _REF_WL[1]
    v, dt, _ = cpu_run(wl, sample_steps, workdir, None if cal is None else cal[k % cal.shape[0]])
    return v, dt


_REF_WL = [None, None]


def describe(name, cfg, n, T):
    if cfg["kind"] == "ensemble":
        return (f"{name}: NSGA-II calibration ensemble (BASELINE configs[3]), SHIA/TETIS 5-tank model, {n}-cell synthetic basin "
                f"({cfg['nc']}x{cfg['nr']} G2 raster), {T} hourly steps, {cfg['members']} parameter sets per GPU")
    sub = " + SED sediments + landslides, kinematic hillslope speeds" if cfg["sediments"] else ""
    return (f"{name}: SHIA/TETIS 5-tank run{sub}, {n}-cell synthetic basin ({cfg['nc']}x{cfg['nr']} G2 raster), "
            f"{T} steps of {cfg['dt']:.0f} s")


def run_reference(args):
    """--impl reference: the reference's CPU implementation (its own compiled Fortran through oracle/_ref when that
    library is present, else the oracle port) on the host cores.  The reference's only parallelism is one process per
    ensemble member (wmf.py:872-878 Pool.map), so all cores run one member each -- for the ensemble workload with that
    member's own parameter set -- on a bounded sample of the workload's steps."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    import oracle
    oracle.build()
    name = args.workload
    cfg = wl_cfg(name)
    nc, nr, T = cfg["nc"], cfg["nr"], cfg["T"]
    # the topology must come from somewhere that is not our GPU code: the CPU reference traces it
    from wmf_b200 import synth
    DIR, DEM = synth.make_raster(cfg["gen"], nc, nr, seed=0)
    from oracle import ref as _ref
    ocu = _ref.RefCu() if _ref.available() else oracle.Cu(fast=True)
    x, y, n, b, acum, lng, pend = trace_basin(ocu, DIR, DEM, nc, nr, cfg["dx"])
    sample_steps = max(2, min(T, args.ref_sample_steps))
    inp = make_inputs(cfg, b, acum, lng, pend)
    # keep only the rain records the sample touches (memory: every worker holds a copy)
    pos = inp["pos_evento"][:sample_steps]
    keep = np.unique(np.concatenate([[1], pos]))
    remap = np.zeros(inp["rain"].shape[0] + 1, np.int32)
    remap[keep] = np.arange(1, keep.size + 1)
    inp["rain"] = np.ascontiguousarray(inp["rain"][keep - 1])
    inp["pos_evento"] = np.concatenate([remap[pos], np.ones(T - sample_steps, np.int32)])
    _REF_WL[0] = dict(inp=inp, n=n)
    cores = os.cpu_count() or 1
    _REF_WL[1] = synth.make_ensemble_calibs(max(cores, 2), seed=0) if cfg["kind"] == "ensemble" else None
    ctx = mp.get_context("fork")
    times = []
    tmp = tempfile.TemporaryDirectory(prefix="wmf_ref_")
    cpu_prepare(_REF_WL[0], sample_steps, tmp.name)          # writes the rain files once, before the pool forks
    with ctx.Pool(cores) as pool:
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            pool.map(_ref_worker, [(sample_steps, tmp.name, k) for k in range(cores)])
            dt = time.perf_counter() - t0
            if it >= args.warmup:
                times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = cores * n * sample_steps / (ms / 1e3)
    line = {
        "impl": "reference", "metric": "shia_cell_steps_per_s", "value": value, "unit": "cell-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": describe(name, cfg, n, T), "cells": n, "steps_per_run": T},
        "cpu_baseline": {"value": value, "unit": "cell-steps/s", "cores": cores, "kind": cpu_kind(),
                         "sample": f"first {sample_steps} of {T} steps of the same basin and rainfall, one ensemble member "
                                   f"per core ({cores} processes, as wmf.py:872-878 does); {cpu_label()}"},
        "e2e": {"value": value, "unit": "cell-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------------------------------
# GPU arms
# ---------------------------------------------------------------------------------------------------------------------
_STATIC_KEYS = ("drena", "unit_type", "hill_long", "stream_long", "elem_area", "max_capilar", "max_gravita", "max_aquifer",
                "hill_slope", "stream_slope", "v_coef", "h_coef", "h_exp", "storage", "control", "control_h", "krus", "crus",
                "prus", "parliac", "sl_zs", "sl_gammas", "sl_cohesion", "sl_frictionangle", "sl_radslope")


def static_bytes(m):
    a = object.__getattribute__(m, "__dict__")["_a"]
    return sum(v.nbytes for k, v in a.items() if k in _STATIC_KEYS)


def time_single(args, env, wl, warmup, steps, e2e_steps):
    """One basin, one run per step (C2 / C5): resident arm through the marshalling layer with device buffers, e2e arm
    through the drop-in call models.shia_v1 with pinned host buffers."""
    import torch
    import torch.distributed as dist
    from wmf_b200 import ModelsModule, synth
    ctx, stream, world, rank, local = env["ctx"], env["stream"], env["world"], env["rank"], env["local"]
    m = ModelsModule(ctx)
    inp, n, T = wl["inp"], wl["n"], wl["T"]
    synth.apply_to_models(m, inp)
    calibs = synth.make_ensemble_calibs(max(world, 1), seed=0)
    calib = calibs[rank % calibs.shape[0]][None, :].copy()          # one ensemble member per GPU
    n_cont = int(np.count_nonzero(inp["control"])) + 1
    m.make_resident()
    d_rain = torch.from_numpy(inp["rain"]).to(f"cuda:{local}")
    pos = inp["pos_evento"]
    want = {"q", "stoout", "balance", "mean_rain"}
    if wl["cfg"]["sediments"]:
        want |= {"qsed", "vs", "vd", "vsc", "vdc", "volero", "voldepo", "erot", "dept"}
    if wl["cfg"]["slides"]:
        want |= {"sl_slideocurrence", "sl_slideacumulate", "sl_slidencelltime", "sl_riskvector", "sl_zcrit", "sl_zmin", "sl_zmax", "sl_bo"}
    qobs = [None]

    def step_resident():
        out = m._run(calib, n_cont, 1, T, d_rain, pos, device_out=True, want=want)
        if qobs[0] is None:
            qobs[0] = out["q"][:, 0].contiguous()                    # "observed" series = this member's first run
        scores = torch.empty((1, 2), dtype=torch.float32, device=f"cuda:{local}")
        ctx.check(ctx.L.wmf_eval_scores(ctx.h, out["q"].data_ptr(), 1, n_cont, T, 0, qobs[0].data_ptr(), scores.data_ptr()))
        if world > 1:                                                # the only collective: gather the members' scores
            allsc = [torch.empty_like(scores) for _ in range(world)]
            with torch.cuda.stream(stream):
                dist.all_gather(allsc, scores)
        return out

    for _ in range(warmup):
        step_resident()
    env["barrier"]()
    env["clock_start"]()
    l0 = ctx.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wave_ms = []
    ev0.record(stream)
    for _ in range(steps):
        step_resident()
        wave_ms.append(ctx.last_kernel_ms())
    ev1.record(stream)
    env["barrier"]()
    clocks = env["clock_stop"]()
    launches = ctx.launches - l0
    ms = ev0.elapsed_time(ev1) / steps
    n_levels, n_waves = ctx.last_shia_stats()
    # --- e2e arm: the drop-in call with host buffers (pinned), copies inside the timed region ---
    m.make_resident(False)
    pin_rain = torch.from_numpy(inp["rain"]).pin_memory()
    m.set_rain(pin_rain.numpy(), pos)
    h2d = static_bytes(m) + pin_rain.numel() * 4 + T * 8 + 44
    d2h = 4 * (n_cont * T * 3 + 3 * T + T + 5 * n + 4 * n + T + n)
    if wl["cfg"]["sediments"]:
        d2h += 4 * (n_cont * 3 * T + 14 * n + 6)
    if wl["cfg"]["slides"]:
        d2h += 4 * (7 * n + T)

    def step_e2e():
        return m.shia_v1(None, None, calib[0], n_cont, 1, T, None, None, n)

    for _ in range(max(warmup, 3)):            # also fills torch's pinned-memory cache for the output buffers
        step_e2e()
    env["barrier"]()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(e2e_steps):
        ret = step_e2e()
    e1.record(stream)
    env["barrier"]()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3) / e2e_steps
    assert np.isfinite(ret[0]).all()
    ms, e2e_ms, wave = env["max_over_ranks"]([ms, e2e_ms, float(np.mean(wave_ms))])
    del d_rain
    bpc = BYTES_PER_CELL_STEP + (BYTES_SED if wl["cfg"]["sediments"] else 0.0) + (BYTES_SLIDES if wl["cfg"]["slides"] else 0.0)
    return dict(ms=ms, e2e_ms=e2e_ms, wave_ms=wave, launches=int(launches), h2d=int(h2d), d2h=int(d2h), levels=n_levels,
                waves=n_waves, units=float(n) * T, bytes_per_unit=bpc, clocks=clocks, models=m, n_cont=n_cont,
                kernel="k_shia_tb<4,%s,%s,%s>" % tuple(str(bool(v)).lower() for v in (
                    any(s == 2 for s in wl["cfg"]["speed_type"]), wl["cfg"]["sediments"], wl["cfg"]["slides"])) + (
                    " + k_sed_tb<4> behind every wave" if wl["cfg"]["sediments"] else ""))


def time_ensemble(args, env, wl, warmup, steps, e2e_steps):
    """The calibration population through wmf_b200.ensemble.evaluate_population: cfg['members'] parameter sets per rank."""
    import torch
    from wmf_b200 import ModelsModule, synth
    from wmf_b200.ensemble import evaluate_population
    ctx, stream, world, rank, local = env["ctx"], env["stream"], env["world"], env["rank"], env["local"]
    cfg = wl["cfg"]
    m = ModelsModule(ctx)
    inp, n, T = wl["inp"], wl["n"], wl["T"]
    synth.apply_to_models(m, inp)
    P = cfg["members"] * world
    calibs = synth.make_ensemble_calibs(P, seed=0)
    n_cont = int(np.count_nonzero(inp["control"])) + 1
    pos = inp["pos_evento"]
    m.make_resident()
    d_rain = torch.from_numpy(inp["rain"]).to(f"cuda:{local}")
    # "observations": the hydrograph of parameter set 0 at the outlet (its Nash score must come back as 1)
    q0 = m._run(calibs[:1], n_cont, 1, T, d_rain, pos, want={"q"})["q"][0].copy()
    stats = {}

    def step_resident():
        return evaluate_population(m, calibs, q0, n_cont, T, d_rain, pos, row=0, members_per_launch=cfg["batch"], stats=stats)

    for _ in range(warmup):
        step_resident()
    env["barrier"]()
    env["clock_start"]()
    l0 = ctx.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stats.clear()
    ev0.record(stream)
    for _ in range(steps):
        sc = step_resident()
    ev1.record(stream)
    env["barrier"]()
    clocks = env["clock_stop"]()
    launches = ctx.launches - l0
    ms = ev0.elapsed_time(ev1) / steps
    wave = stats.get("wave_ms", 0.0) / steps
    n_waves = stats.get("waves", 0) // steps
    n_levels = ctx.last_shia_stats()[0]
    sc_h = sc.cpu().numpy()
    assert sc_h.shape == (P, 2) and abs(float(sc_h[0, 0]) - 1.0) < 1e-6, "member 0 must reproduce its own hydrograph (Nash = 1)"
    # --- e2e arm: host inputs; the rain table crosses PCIe once per evaluation, 1/world of it per rank ---
    m.make_resident(False)
    pin_rain = torch.from_numpy(inp["rain"]).pin_memory().numpy()
    h2d = static_bytes(m) + (pin_rain.size * 4 + world - 1) // world + T * 8 + calibs[rank::world].nbytes + T * 4
    d2h = 4 * 2 * P

    def step_e2e():
        return evaluate_population(m, calibs, q0, n_cont, T, pin_rain, pos, row=0, members_per_launch=cfg["batch"]).cpu().numpy()

    for _ in range(max(1, min(warmup, 2))):
        step_e2e()
    env["barrier"]()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(e2e_steps):
        sc2 = step_e2e()
    e1.record(stream)
    env["barrier"]()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3) / e2e_steps
    assert np.array_equal(sc2, sc_h, equal_nan=True), "host-buffer evaluation differs from the resident one"
    ms, e2e_ms, wave = env["max_over_ranks"]([ms, e2e_ms, wave])
    m.make_resident()
    return dict(ms=ms, e2e_ms=e2e_ms, wave_ms=wave, launches=int(launches), h2d=int(h2d), d2h=int(d2h), levels=n_levels,
                waves=n_waves, units=float(n) * T * cfg["members"], bytes_per_unit=BYTES_PER_MEMBER_CELL_STEP, clocks=clocks,
                models=m, n_cont=n_cont, calibs=calibs, d_rain=d_rain, scores=sc_h,
                kernel="k_shia_tb<K,...> member-minor (K = 16, 8 or 4 by the ring-memory rule)")


def ensemble_parity(args, env, wl, res, workdir):
    """cpu_baseline of the ensemble workload = whole members of the population run by the reference's CPU path on one
    core; their hydrographs double as an in-bench parity check of the GPU's (north-star tolerance 1e-5)."""
    import torch
    m, calibs, n_cont, T = res["models"], res["calibs"], res["n_cont"], wl["T"]
    rng = np.random.default_rng(1)
    picks = sorted(int(k) for k in rng.choice(calibs.shape[0], size=max(1, args.cpu_members), replace=False))
    out = m._run(np.ascontiguousarray(calibs[picks]), n_cont, 1, T, res["d_rain"], wl["inp"]["pos_evento"], device_out=True,
                 want={"q"}, ensemble=True)
    q = out["q"].cpu().numpy().reshape(len(picks), T, n_cont)
    secs, worst = 0.0, 0.0
    for j, k in enumerate(picks):
        v, dt, ref = cpu_run(wl, T, workdir, calib=calibs[k])
        secs += dt
        want, got = np.asarray(ref["q"], np.float64), q[j].T.astype(np.float64)
        tol = 1e-5 * np.abs(want) + 1e-7 * np.abs(want).max()
        worst = max(worst, float((np.abs(got - want) / tol).max()))
    value = wl["n"] * T * len(picks) / secs
    return ({"value": value, "unit": "cell-steps/s", "cores": 1, "kind": cpu_kind(),
             "sample": f"{len(picks)} whole members of the population (parameter sets {picks}, {T} steps each, {secs:.1f} s), "
                       f"single thread as the reference runs one member; {cpu_label()}"},
            {"members": picks, "worst_error_in_units_of_tolerance_1e-5": worst, "pass": bool(worst <= 1.0),
             "what": "Q at the outlet and 16 control points, every step, GPU ensemble launch vs the CPU reference"})


def bench_c5(args, env, cu, peak):
    """BASELINE configs[4] as a second field: SED sediments + landslides on the 2.09 M-cell 12 m basin, one storm through the
    resident marshalling layer and 8 storm scenarios as ONE launch sequence (scenario = member axis,
    ensemble.run_scenario_ensemble); CUDA-event time of the launches, the reference's Fortran on a few steps next to it."""
    import torch
    from wmf_b200 import ModelsModule, synth
    from wmf_b200.ensemble import run_scenario_ensemble
    ctx = env["ctx"]
    wl = build_workload("c5", cu, 0, verbose=True)
    inp, n, T = wl["inp"], wl["n"], wl["T"]
    m = ModelsModule(ctx)
    synth.apply_to_models(m, inp)
    m.make_resident()
    d_rain = torch.from_numpy(inp["rain"]).to(f"cuda:{env['local']}")
    n_cont = int(np.count_nonzero(inp["control"])) + 1
    pos = np.asarray(inp["pos_evento"], np.int32)
    want = {"q", "stoout", "qsed", "sl_slideocurrence"}
    calib = np.asarray(inp["calib"], np.float32)[None, :]
    res = {"workload": describe("c5", wl["cfg"], n, T)}
    for label, S in (("one_scenario", 1), ("eight_scenarios", 8)):
        scen = np.stack([np.roll(pos, 11 * k) for k in range(S)]).astype(np.int32)   # the storm arriving 11 steps later each
        best_wall, best_ev = None, None
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if S == 1:
                m._run(calib, n_cont, 1, T, d_rain, pos, device_out=True, want=want)
            else:
                run_scenario_ensemble(m, inp["calib"], scen, n_cont, T, d_rain, scenarios_per_launch=S, device_out=True, want=want)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if best_wall is None or dt < best_wall:
                best_wall, best_ev = dt, ctx.last_kernel_ms()
        bpc = BYTES_PER_CELL_STEP + BYTES_SED + BYTES_SLIDES
        ach = bpc * n * T * S / (best_ev / 1e3) / 1e9
        res[label] = {"scenarios": S, "ms_per_scenario": 1e3 * best_wall / S, "value": n * T * S / best_wall, "unit": "cell-steps/s",
                      "launch_ms": best_ev,
                      "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                                   "bytes_per_unit": bpc, "kernel": "k_shia_tb<K,KIN,SED,SLIDES> + k_sed_tb<K> (K = 4, 8 from 4 members on)"}}
    with tempfile.TemporaryDirectory(prefix="wmf_cpu_") as tmpd:
        steps = max(2, min(T, args.cpu_sample_steps // 12))
        cpu_v, cpu_s = cpu_sample(wl, steps, tmpd)
    res["cpu_baseline"] = {"value": cpu_v, "unit": "cell-steps/s", "cores": 1, "kind": cpu_kind(),
                           "sample": f"first {steps} of {T} steps of one scenario ({cpu_s:.1f} s), single thread; {cpu_label()}"}
    del d_rain, m
    torch.cuda.empty_cache()
    return res


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- wmf_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    from wmf_b200 import Context, CuModule
    ctx = Context(local)
    cu = CuModule(ctx)
    wl = build_workload(args.workload, cu, rank, verbose=(rank == 0))
    cfg, n, T = wl["cfg"], wl["n"], wl["T"]
    stream = torch.cuda.Stream()
    ctx.use_stream(stream.cuda_stream)
    sampler = ClockSampler(local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(vals):
        tt = torch.tensor(vals, dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return [float(v) for v in tt.cpu()]

    env = dict(ctx=ctx, stream=stream, world=world, rank=rank, local=local, barrier=barrier, max_over_ranks=max_over_ranks,
               clock_start=(sampler.start if rank == 0 else (lambda: None)),
               clock_stop=(sampler.stop if rank == 0 else (lambda: None)))
    timer = time_ensemble if cfg["kind"] == "ensemble" else time_single
    res = timer(args, env, wl, args.warmup, args.steps, max(1, min(args.steps, 5)))
    if rank == 0:
        peak, peak_kind = peaks()
        ms, e2e_ms, wave, units = res["ms"], res["e2e_ms"], res["wave_ms"], res["units"]
        achieved = res["bytes_per_unit"] * units / (wave / 1e3) / 1e9
        traffic, traffic_note = TRAFFIC[cfg["kind"]] if not cfg["sediments"] else TRAFFIC["c5"]
        if cfg["kind"] == "ensemble":
            traffic = traffic * units / max(res["waves"], 1)
        extra = {"levels": res["levels"], "waves_per_run": res["waves"], "wave_kernel_ms_per_step": wave,
                 "setup_and_epilogue_ms_per_step": ms - wave}
        with tempfile.TemporaryDirectory(prefix="wmf_cpu_") as tmpd:
            if cfg["kind"] == "ensemble":
                cpu_b, parity = ensemble_parity(args, env, wl, res, tmpd)
                extra["parity_vs_cpu"] = parity
            else:
                cpu_steps = max(2, min(T, args.cpu_sample_steps))
                cpu_v, cpu_s = cpu_sample(wl, cpu_steps, tmpd)
                cpu_b = {"value": cpu_v, "unit": "cell-steps/s", "cores": 1, "kind": cpu_kind(),
                         "sample": f"first {cpu_steps} of {T} steps of the same basin ({cpu_s:.1f} s), single thread as the "
                                   f"reference runs; {cpu_label()}"}
        line = {
            "metric": "shia_cell_steps_per_s", "value": world * units / (ms / 1e3), "unit": "cell-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": describe(args.workload, cfg, n, T), "cells": n, "steps_per_run": T,
                       "members_per_gpu": cfg.get("members", 1),
                       "l2": "state and outflow ring of a launch (%.1f GB) are larger than L2" % (
                           (cfg.get("batch", 1) * n * 36 * 2) / 1e9) if cfg["kind"] == "ensemble" else
                             "inputs larger than L2 (rain table %.2f GB + %.0f MB state/maps)" % (
                                 wl["inp"]["rain"].nbytes / 1e9, 150.0 * n / 1e6)},
            "e2e": {"value": world * units / (e2e_ms / 1e3), "unit": "cell-steps/s",
                    "h2d_bytes_per_step": res["h2d"], "d2h_bytes_per_step": res["d2h"], "ms_per_step": e2e_ms},
            "gpu_launches": res["launches"],
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_kind": peak_kind, "kernel": res["kernel"],
                         "bytes_per_unit": res["bytes_per_unit"], "launches_per_step": res["waves"],
                         "achieved_per_launch_bytes": res["bytes_per_unit"] * units / max(res["waves"], 1),
                         "traffic_note": traffic_note},
            "cpu_baseline": cpu_b, "clocks": res["clocks"], "extra": extra,
        }
        if world == 1 and not args.no_extras:
            res.pop("d_rain", None)
            res.pop("models", None)
            torch.cuda.empty_cache()
            if cfg["kind"] == "ensemble":          # the single-basin run of BASELINE configs[1] as a second field
                wl2 = build_workload("c2", cu, 0, verbose=True)
                r2 = time_single(args, dict(env, clock_start=lambda: None, clock_stop=lambda: None), wl2, 3, 5, 3)
                u2 = r2["units"]
                ach2 = r2["bytes_per_unit"] * u2 / (r2["wave_ms"] / 1e3) / 1e9
                with tempfile.TemporaryDirectory(prefix="wmf_cpu_") as tmpd:
                    cpu_v, cpu_s = cpu_sample(wl2, min(wl2["T"], args.cpu_sample_steps), tmpd)
                extra["c2"] = {
                    "workload": describe("c2", wl2["cfg"], wl2["n"], wl2["T"]), "value": u2 / (r2["ms"] / 1e3), "unit": "cell-steps/s",
                    "ms_per_step": r2["ms"], "wave_kernel_ms": r2["wave_ms"], "levels": r2["levels"], "waves_per_run": r2["waves"],
                    "e2e": {"value": u2 / (r2["e2e_ms"] / 1e3), "unit": "cell-steps/s", "ms_per_step": r2["e2e_ms"],
                            "h2d_bytes_per_step": r2["h2d"], "d2h_bytes_per_step": r2["d2h"]},
                    "roofline": {"bound": "hbm", "achieved": ach2, "peak": peak, "unit": "GB/s", "frac": ach2 / peak,
                                 "traffic": TRAFFIC["single"][0], "kernel": r2["kernel"], "bytes_per_unit": r2["bytes_per_unit"],
                                 "traffic_note": TRAFFIC["single"][1]},
                    "cpu_baseline": {"value": cpu_v, "unit": "cell-steps/s", "cores": 1, "kind": cpu_kind(),
                                     "sample": f"first {min(wl2['T'], args.cpu_sample_steps)} steps ({cpu_s:.1f} s), single thread"}}
                del r2, wl2
                torch.cuda.empty_cache()
            if not args.skip_c5:
                extra["c5"] = bench_c5(args, env, cu, peak)
            pa = {"c1": bench_path_a(cu, 512, 512, reps=5)}
            if not args.skip_c3:
                pa["c3"] = bench_path_a(cu, 20000, 20000, reps=2, cpu=not args.skip_c3_cpu, hand=True)
            extra["path_a"] = pa
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def _quiet_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on stdout
    when NCCL_DEBUG=VERSION is set on the box), so file descriptor 1 is pointed at stderr for the whole run and the
    JSON line goes to the saved descriptor."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sample-steps", type=int, default=120, help="single-run workloads: steps timed on one CPU core")
    ap.add_argument("--cpu-members", type=int, default=2, help="ensemble workload: whole members re-run on one CPU core")
    ap.add_argument("--ref-sample-steps", type=int, default=100, help="steps per reference-arm iteration (per member)")
    ap.add_argument("--no-extras", action="store_true", help="skip extra.c2 and extra.path_a")
    ap.add_argument("--skip-c3", action="store_true", help="skip the 20000 x 20000 Path A raster")
    ap.add_argument("--skip-c5", action="store_true", help="skip extra.c5 (sediments + landslides, storm scenarios)")
    ap.add_argument("--skip-c3-cpu", action="store_true", help="C3 without the CPU reference run (13 GB of host memory, ~30 s)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
