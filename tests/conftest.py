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
