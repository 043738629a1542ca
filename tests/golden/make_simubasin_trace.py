#!/usr/bin/env python
"""Record what the reference's own `wmf.py` asks of its Fortran modules while it runs a model end to end.

`/root/reference/wmf/wmf.py` is imported UNMODIFIED (plot / GIS libraries that are not installed are stubbed in
`sys.modules`), with the f2py modules `cu` and `models` replaced by recording proxies in front of the CPU oracle
(oracle.Cu / oracle.shia_v1 -- bit-identical to the reference's compiled Fortran, tests/test_reference_pin.py).  The script
then drives the modelling workflow exactly as a user would:

    SimuBasin(lat, lon, DEM, DIR, ...)  ->  set_Geomorphology  ->  set_PhysicVariables x 7  ->  set_Storage x 5
    ->  set_Control  ->  rain_interpolate_idw  ->  run_shia

and writes every attribute assignment and every call (name, positional arguments as wmf.py passed them, results) to
`tests/golden/simubasin_trace.pkl.gz`.  `tests/test_gpu_simubasin_replay.py` replays that trace against
`wmf_b200.cu` / `wmf_b200.models` on the GPU box -- where the reference tree does not exist -- and requires the same
results: the drop-in boundary is exercised with the reference's own call sequence, dims, dtypes and in-place mutations.

Run here (CPU, reference tree present):  python tests/golden/make_simubasin_trace.py
"""
from __future__ import annotations

import gzip
import importlib.util
import os
import pickle
import sys
import tempfile
import types
import warnings
from unittest import mock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
WMF_PY = "/root/reference/wmf/wmf.py"

TRACE: list = []


def snap(v):
    """Deep copy of an argument / result for the trace (arrays by value)."""
    if isinstance(v, np.ndarray):
        return np.array(v, copy=True)
    if isinstance(v, (list, tuple)):
        return type(v)(snap(x) for x in v)
    if isinstance(v, (np.generic,)):
        return v.item()
    return v


class Recorder:
    """Stands where the f2py module object stood.  Attribute reads return the backend's live value (so wmf.py's in-place
    edits such as `models.h_coef[pos] = vec`, wmf.py:3882, reach it); before every call the arrays that changed since the
    last event are written to the trace as assignments."""

    def __init__(self, name, backend, arrays=()):
        object.__setattr__(self, "_n", name)
        object.__setattr__(self, "_b", backend)
        object.__setattr__(self, "_seen", {})
        object.__setattr__(self, "_arrays", set(arrays))

    def _flush(self):
        b, seen = self._b, self._seen
        for k in sorted(self._arrays):
            v = getattr(b, k, None)
            if isinstance(v, np.ndarray):
                key = (v.shape, v.dtype.str, v.tobytes())
                if seen.get(k) != key:
                    seen[k] = key
                    TRACE.append(("set", self._n, k, snap(v)))

    def __setattr__(self, k, v):
        setattr(self._b, k, v)
        cur = getattr(self._b, k, None)
        if isinstance(cur, np.ndarray):
            self._arrays.add(k)
            self._seen[k] = (cur.shape, cur.dtype.str, cur.tobytes())
            TRACE.append(("set", self._n, k, snap(v)))          # what wmf.py assigned (before the f2py cast)
        else:
            TRACE.append(("set", self._n, k, snap(v)))

    def __getattr__(self, k):
        v = getattr(self._b, k)
        if callable(v):
            def call(*args, **kw):
                assert not kw, f"{self._n}.{k} called with keywords {list(kw)}"
                for r in RECORDERS:
                    r._flush()
                res = v(*args)
                TRACE.append(("call", self._n, k, snap(args), snap(res)))
                return res
            return call
        if isinstance(v, np.ndarray):
            self._arrays.add(k)
        return v


RECORDERS: list = []


# ---------------------------------------------------------------------------------------------------------------------
# f2py-shaped fronts of the oracle: optional trailing dimension arguments accepted (wmf.py passes them positionally)
# ---------------------------------------------------------------------------------------------------------------------
def make_cu_backend():
    import inspect
    import oracle

    class F2pyCu(oracle.Cu):
        pass

    def widen(fn, extra_ok):
        n_named = len([p for p in inspect.signature(fn).parameters.values() if p.name != "self"])

        def w(self, *args):
            assert len(args) <= n_named + extra_ok, (fn.__name__, len(args))
            return fn(self, *args[:n_named])
        w.__name__ = fn.__name__
        return w

    for name, fn in inspect.getmembers(oracle.Cu, predicate=inspect.isfunction):
        if not name.startswith("_"):
            setattr(F2pyCu, name, widen(fn, 4))
    return F2pyCu()


class OracleModels:
    """Module `models` on top of the oracle: a bag of module variables with f2py's cast-on-assignment, `shia_v1` with the
    f2py positional signature (modelsmodule.c:375-408), `rain_idw`, and the direct-access file helpers."""
    _INT = {"drena", "unit_type", "control", "control_h", "guarda_cond", "guarda_vfluxes", "speed_type"}

    def __init__(self):
        d = self.__dict__
        d["_a"] = {"speed_type": np.ones(3, np.int32)}
        d["_s"] = dict(dt=60.0, dxp=30.0, retorno_gr=0.0, retorno_aq=0.0, storage_constant=0.001, calc_niter=5, verbose=0,
                       sim_sediments=0, sim_slides=0, sim_floods=0, save_storage=0, save_speed=0, save_vfluxes=0, save_rc=0,
                       save_retorno=0, separate_fluxes=0, separate_rain=0, show_storage=0, show_speed=0, show_mean_speed=0,
                       show_mean_retorno=0, rain_first_point=1, nceldas=0, sed_factor=1.0, sl_fs=1.0, sl_gammaw=9.8,
                       sl_gullienogullie=0)
        from wmf_b200 import _models as shim          # every scalar module variable of modelsmodule.c:3721-3856, at its default
        for table in (shim._SCALARS_F, shim._SCALARS_I):
            for k, v in table.items():
                d["_s"].setdefault(k, v)

    def __getattr__(self, k):
        d = self.__dict__
        if k in d["_a"]:
            return d["_a"][k]
        if k in d["_s"]:
            return d["_s"][k]
        raise AttributeError(k)

    def __setattr__(self, k, v):
        d = self.__dict__
        if isinstance(v, (np.ndarray, list, tuple)) and np.ndim(v) >= 1:
            a = np.asfortranarray(np.asarray(v), dtype=np.int32 if k in self._INT else np.float32)
            if a.ndim == 1 and k not in ("speed_type", "evpserie", "guarda_cond", "guarda_vfluxes", "wi", "diametro"):
                a = np.asfortranarray(a.reshape(1, -1))
            d["_a"][k] = a
        else:
            d["_s"][k] = v

    # -- file helpers (same formats as the Fortran's direct-access records) --
    def shia_v1(self, ruta_bin, ruta_hdr, calib, n_cont, n_conth, n_reg, stoin=None, hspeedin=None, n_cel=None, *paths):
        import oracle
        from wmf_b200._models import read_rain_hdr
        a, s = self.__dict__["_a"], self.__dict__["_s"]
        N, T = int(n_cel), int(n_reg)
        _, pos = read_rain_hdr(str(ruta_hdr).strip(), T, int(s["rain_first_point"]))
        n_rec = os.path.getsize(str(ruta_bin).strip()) // (4 * N)
        rain = np.fromfile(str(ruta_bin).strip(), dtype=np.int32, count=n_rec * N).reshape(n_rec, N)
        inp = {k: v for k, v in a.items()}
        inp.update({k: v for k, v in s.items()})
        inp.update(calib=np.asarray(calib, np.float32), n_reg=T, rain=rain, pos_evento=np.ascontiguousarray(pos, np.int32),
                   evpserie=np.asarray(a["evpserie"], np.float32)[:T], nceldas=N)
        st = np.asarray(stoin, np.float32) if stoin is not None else None
        if st is not None and st[0, 0] > 0:
            inp["stoin"] = np.asfortranarray(st)
        hs = np.asarray(hspeedin, np.float32) if hspeedin is not None else None
        if hs is not None and hs[0, 0] > 0:
            inp["hspeedin"] = np.asfortranarray(hs)
        o = oracle.shia_v1(inp)
        d = self.__dict__
        d["_a"]["mean_rain"] = np.asfortranarray(o["mean_rain"].reshape(1, -1))
        d["_a"]["acum_rain"] = np.asfortranarray(o["acum_rain"].reshape(1, -1))
        nc, nh = int(n_cont), int(n_conth)
        z = lambda *sh: np.zeros(sh, np.float32, order="F")
        return (o["q"], o.get("qsed", z(nc, 3, T)), o.get("qseparated", z(nc, 3, T)), o["hum"], o["st1"], o["st3"], o["balance"],
                o["speed"], o["areacontrol"], o["stoout"], o.get("qsep_byrain", z(nc, 2, T)))

    def rain_idw(self, xy_basin, coord, rain, pp, nhills, ruta, umbral, maskvector, *dims):
        import oracle
        mask = None
        if maskvector is not None and int(np.asarray(maskvector).sum()) != np.asarray(xy_basin).shape[1]:
            mask = maskvector
        r = oracle.rain_idw(xy_basin, coord, rain, pp, umbral, mask, int(nhills) if mask is not None else 0)
        with open(str(ruta).strip(), "wb") as f:
            f.write(np.ascontiguousarray(r["table"], np.int32).tobytes())
        return r["mean_rain"], r["pos_ids"]


def stub_modules():
    """Everything wmf.py imports that is not installed here and that the modelling workflow never touches."""
    for name in ("matplotlib", "matplotlib.path", "matplotlib.ticker", "matplotlib.pyplot", "pylab", "mpl_toolkits",
                 "mpl_toolkits.axes_grid1", "osgeo", "osgeo.ogr", "osgeo.osr", "gdal", "netCDF4", "deap", "deap.base",
                 "deap.creator", "deap.tools", "pysheds", "pysheds.grid", "cartopy", "rasterio"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                m = mock.MagicMock(name=name)
                m.__path__ = []
                sys.modules[name] = m


def load_wmf(cu_rec, models_rec):
    stub_modules()
    mcu, mmod = types.ModuleType("cu"), types.ModuleType("models")
    mcu.cu, mmod.models = cu_rec, models_rec
    sys.modules["cu"], sys.modules["models"] = mcu, mmod
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spec = importlib.util.spec_from_file_location("wmf_reference", WMF_PY)
        w = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(w)
    return w


class InternedStr(str):
    """wmf.py compares `self.modelType[0] is 'c'` (wmf.py:4022, 4403 ...): true on the CPython 3.7 it was written for, false
    on 3.12 where indexing a str no longer returns the interned one-character object.  Handing SimuBasin a str whose items
    are interned runs the unmodified source as its author's interpreter did."""

    def __getitem__(self, i):
        return sys.intern(str.__getitem__(self, i))


def main(out_path=None):
    from wmf_b200 import synth
    out_path = out_path or (sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "simubasin_trace.pkl.gz"))
    cu_b, mod_b = make_cu_backend(), OracleModels()
    cu_r, mod_r = Recorder("cu", cu_b), Recorder("models", mod_b)
    RECORDERS.extend([cu_r, mod_r])
    w = load_wmf(cu_r, mod_r)
    w.Global_EPSG = 32618
    # what read_map_raster leaves in module cu (wmf.py:319-333)
    nc, nr, dx = 72, 60, 30.0
    DIR, DEM = synth.make_raster("G2", nc, nr, 3)
    cu_r.ncols, cu_r.nrows, cu_r.xll, cu_r.yll, cu_r.dx, cu_r.dy, cu_r.dxp, cu_r.nodata = nc, nr, 0.0, 0.0, dx, dx, dx, synth.NODATA
    x, y = synth.outlet_xy(nc, nr, 0.0, 0.0, dx)
    notes = ["wmf.py:4022,4024,4033,4038,4403,4405 compare `self.modelType[0] is 'c'`: false on CPython 3.12 (no interned "
             "one-character result of str indexing); modelType is passed as a str subclass whose items are interned"]
    with mock.patch.object(w.Basin, "__GetBasinPolygon__", lambda self: None, create=True):
        sb = w.SimuBasin(x, y, np.asfortranarray(DEM), np.asfortranarray(DIR), name="trace", threshold=60, dt=3600.0, noData=synth.NODATA,
                         modelType=InternedStr("cells"), controlNodos=True)
    N = sb.ncells
    sb.set_Geomorphology(thresholdes=[30, 500])
    rng = np.random.default_rng(0)
    inp = synth.make_shia_inputs(sb.structure, sb.CellAcum, sb.CellLong, sb.CellSlope, dxp=dx, n_reg=48, seed=0)
    for k, name in enumerate(("v_coef", "v_coef", "v_coef", "v_coef")):
        sb.set_PhysicVariables(name, inp["v_coef"][k], k)
    for k in range(4):
        sb.set_PhysicVariables("h_coef", inp["h_coef"][k], k)
        sb.set_PhysicVariables("h_exp", inp["h_exp"][k], k)
    sb.set_PhysicVariables("capilar", inp["max_capilar"][0], 0)
    sb.set_PhysicVariables("gravit", inp["max_gravita"][0], 0)
    sb.set_PhysicVariables("aquifer", inp["max_aquifer"][0], 0)
    for k in range(5):
        sb.set_Storage(inp["storage"][k], k)
    # rain gauges -> record file through the reference's own interpolation front end
    coord, gauges = synth.make_stations(nc, nr, dx, 9, 48, 0)
    import pandas as pd
    tmp = tempfile.mkdtemp(prefix="wmf_trace_")
    rain_path = os.path.join(tmp, "rain.bin")
    idx = pd.date_range("2020-01-01", periods=48, freq="h")
    reg = pd.DataFrame(4.0 * gauges.T, index=idx)            # a storm strong enough to saturate the capillary tank
    try:
        sb.rain_interpolate_idw(coord, reg, rain_path, p=2, threshold=0.0)
    except Exception as e:                                   # noted in the trace header, the run below needs the files
        notes.append(f"rain_interpolate_idw failed under this Python / pandas: {type(e).__name__}: {e}")
        raise
    notes.append("wmf.py:3217 (__GetEVP_Serie__, the default EvpVariable='sun' of run_shia) calls pd.date_range(freq='1H'), which "
                 "pandas >= 3 rejects ('H' was removed): run_shia is called with EvpVariable=False (evpserie = 1)")
    res = sb.run_shia(inp["calib"], rain_path, 47, 1, EvpVariable=False)      # (read_mean_rain slices [start:start+N], wmf.py:599)
    qsim = np.asarray(res[0]["Qsim"] if isinstance(res, tuple) else res["Qsim"])
    assert np.isfinite(qsim).all() and qsim[0].max() > 0, "the reference workflow produced no discharge"
    for r in RECORDERS:
        r._flush()
    files = {}
    for fn in sorted(os.listdir(tmp)):                       # rain.bin (written by models.rain_idw) and rain.hdr (written by wmf.py)
        with open(os.path.join(tmp, fn), "rb") as fh:
            files[fn] = fh.read()
    with gzip.open(out_path, "wb", compresslevel=9) as f:
        pickle.dump({"trace": TRACE, "notes": notes, "files": files, "n_cells": int(N), "python": sys.version.split()[0], "pandas": pd.__version__,
                     "workflow": "SimuBasin -> set_Geomorphology -> set_PhysicVariables -> set_Storage -> rain_interpolate_idw -> run_shia",
                     "rain_dir": tmp}, f, protocol=4)
    calls = [e for e in TRACE if e[0] == "call"]
    print(f"{len(TRACE)} events ({len(calls)} calls: {sorted({e[1] + '.' + e[2] for e in calls})}) -> {out_path}")
    print(f"basin of {N} cells, Qsim peak {qsim[0].max():.3f} m3/s")
    for n in notes:
        print("note:", n)


if __name__ == "__main__":
    main()
