import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device here; GPU parity tests run with -m gpu on the B200 box")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def make_basin(kind="G2", nc=96, nr=80, seed=0, dx=30.0, xll=0.0, yll=0.0):
    """Synthetic raster + oracle-traced basin; returns a dict shared by CPU and GPU tests."""
    import oracle
    from wmf_b200 import synth
    DIR, DEM = synth.make_raster(kind, nc, nr, seed)
    ocu = oracle.Cu()
    ocu.ncols, ocu.nrows, ocu.xll, ocu.yll = nc, nr, xll, yll
    ocu.dx = ocu.dy = dx
    ocu.dxp = dx if dx >= 1 else 30.0
    ocu.nodata = synth.NODATA
    x, y = synth.outlet_xy(nc, nr, xll, yll, dx)
    n = ocu.basin_find(x, y, DIR)
    b = ocu.basin_cut(n)
    return dict(DIR=DIR, DEM=DEM, cu=ocu, x=x, y=y, n=n, basin=b, nc=nc, nr=nr, dx=dx, xll=xll, yll=yll)


def geo_to(cu, bs):
    o = bs["cu"]
    cu.ncols, cu.nrows, cu.xll, cu.yll, cu.dx, cu.dy, cu.dxp, cu.nodata = (
        o.ncols, o.nrows, o.xll, o.yll, o.dx, o.dy, o.dxp, o.nodata)


def basin_vectors(bs):
    """DEM / DIR vectors, acum, long, pend of the oracle basin."""
    o, b = bs["cu"], bs["basin"]
    dv = o.basin_map2basin(b, bs["DIR"].astype(np.float32), bs["xll"], bs["yll"], bs["dx"], bs["dx"], o.nodata).astype(np.int32)
    de = o.basin_map2basin(b, bs["DEM"], bs["xll"], bs["yll"], bs["dx"], bs["dx"], o.nodata)
    acum, lng, pend, elev = o.basin_basics(b, de, dv)
    return dict(dirv=dv, demv=de, acum=acum, lng=lng, pend=pend, elev=elev)


FUZZ_SED = ("qsed", "vs", "vd", "vsc", "vdc", "volero", "voldepo")
FUZZ_SLIDES = ("sl_zmin", "sl_zcrit", "sl_zmax", "sl_bo", "sl_riskvector", "sl_slideocurrence", "sl_slideacumulate", "sl_slidencelltime")


def fuzz_case(case: int):
    """A random partial basin (irregular tree, up to 8 cells draining into one) with a random combination of the model's
    switches; topology by the oracle.  Returns None when the random outlet catches fewer than 50 cells."""
    import oracle
    from wmf_b200 import synth
    rng = np.random.default_rng(1000 + case)
    nc, nr, dx = int(rng.integers(40, 90)), int(rng.integers(36, 80)), float(rng.choice([12.0, 30.0, 90.0]))
    DIR, (oc_, orow) = synth.make_partial_raster(nc, nr, 100 + case, noise=float(rng.uniform(1.0, 2.2)))
    x, y = dx * (oc_ - 0.5), dx * (nr - orow + 0.5)
    oc = oracle.Cu()
    geo = (nc, nr, 0.0, 0.0, dx, dx, dx, synth.NODATA)
    oc.ncols, oc.nrows, oc.xll, oc.yll, oc.dx, oc.dy, oc.dxp, oc.nodata = geo
    n = oc.basin_find(x, y, DIR)
    if n < 50:
        return None
    b = oc.basin_cut(n)
    par = n - b[0]
    hop = np.zeros(n + 1)
    for i in range(n - 2, -1, -1):
        hop[i] = hop[par[i]] + 1
    demv = (10 + 0.5 * hop[:n] + 0.2 * rng.random(n)).astype(np.float32)
    dirv = oc.basin_map2basin(b, DIR.astype(np.float32), 0, 0, dx, dx, synth.NODATA).astype(np.int32)
    basics = oc.basin_basics(b, demv, dirv)
    acum, lng, pend, _ = basics
    thr = (int(rng.integers(3, 15)), int(rng.integers(20, 120)))
    st = tuple(int(v) for v in rng.choice([1, 2], 3))
    kw = dict(speed_type=st, retorno_gr=int(rng.integers(0, 2)), retorno_aq=int(rng.integers(0, 2)),
              rain_peak_mm=float(rng.uniform(15, 90)), n_control_h=int(rng.integers(0, 4)))
    mode = str(rng.choice(["plain", "sed", "slides", "sed_slides", "sepf", "sepr", "dumps"]))
    if mode in ("sed", "sed_slides"):
        kw["sediments"] = True
    if mode in ("slides", "sed_slides"):
        kw["slides"] = True
    T = int(rng.integers(20, 70))
    inp = synth.make_shia_inputs(b, acum, lng, pend, dxp=dx, dt=float(rng.choice([300.0, 3600.0])), n_reg=T, seed=case,
                                 thresholds=thr, **kw)
    if kw["retorno_gr"]:
        inp["calib"][9] = float(rng.choice([0.02, 1.0]))
    flags = {k: int(rng.integers(0, 2)) for k in ("show_storage", "show_speed", "show_mean_speed")}
    if kw["retorno_gr"]:
        flags["show_mean_retorno"] = int(rng.integers(0, 2))
    extra = ()
    if mode == "sepf":
        flags["separate_fluxes"] = 1
        extra = ("qseparated",)
    if mode == "sepr":
        inp.update(synth.make_rain_types(inp, case))
        extra = ("qsep_byrain", "storage_conv", "storage_stra")
    if mode == "dumps":
        flags.update(save_vfluxes=1, save_rc=1)
        extra = ("mean_vfluxes", "rc_coef")
    if mode in ("sed", "sed_slides"):
        extra = FUZZ_SED
    if mode in ("slides", "sed_slides"):
        extra = tuple(extra) + FUZZ_SLIDES
    if rng.random() < 0.3:
        s0 = np.array(inp["storage"], copy=True)
        s0[0, ::5] = 0.0
        inp["storage"] = s0
    inp.update(flags)
    return dict(inp=inp, mode=mode, extra=extra, flags=flags, n=n, basin=b, DIR=DIR, x=x, y=y, geo=geo, demv=demv, dirv=dirv,
                basics=basics, T=T)
