"""`models` -- host-side mirror of the reference's f2py module object ``models`` (wmf/modelosv2.f90).

Semantics kept from f2py:
  * attribute names are lower-case (``models.max_capilar``, ``models.evpserie``);
  * assigning an array *copies and casts* it into a persistent Fortran-ordered buffer
    (``models.control = np.zeros((1,N))`` becomes int32, wmf.py:3045);
  * reading an attribute returns that same buffer, so ``models.h_coef[pos] = vec``
    (wmf.py:3882) and ``models.storage[p] = vec`` (wmf.py:3952) mutate module state in place;
  * ``shia_v1`` has the reference's positional signature and 11-tuple return
    (modelsmodule.c:375-408), called positionally from ``SimuBasin.run_shia`` (wmf.py:4512-4531).

The time loop itself runs in libwmf_b200.so (``wmf_shia_run``).  The rainfall ``.bin/.hdr`` pair
is parsed here (formats: modelosv2.f90:911-926, 966-1002; writer wmf.py:3349-3362) and handed to
the kernels in memory; ``set_rain`` supplies it without touching the file system.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import ShiaCfg, ShiaInputs, ShiaOutputs, is_torch_cuda, ptr

# name -> (dtype, leading dimension K of the (K,N) array; 0 = 1-D)
_ARRAYS = {
    "drena": (np.int32, 3), "unit_type": (np.int32, 1),
    "hill_long": (np.float32, 1), "hill_slope": (np.float32, 1), "stream_long": (np.float32, 1),
    "stream_slope": (np.float32, 1), "stream_width": (np.float32, 1), "elem_area": (np.float32, 1),
    "v_coef": (np.float32, 4), "v_exp": (np.float32, 4), "h_coef": (np.float32, 4), "h_exp": (np.float32, 4),
    "max_capilar": (np.float32, 1), "max_gravita": (np.float32, 1), "max_aquifer": (np.float32, 1),
    "storage": (np.float32, 5), "control": (np.int32, 1), "control_h": (np.int32, 1),
    "evpserie": (np.float32, 0), "guarda_cond": (np.int32, 0), "guarda_vfluxes": (np.int32, 0),
    "krus": (np.float32, 1), "crus": (np.float32, 1), "prus": (np.float32, 1), "parliac": (np.float32, 3),
    "vd": (np.float32, 3), "vs": (np.float32, 3), "vsc": (np.float32, 3), "vdc": (np.float32, 3),
    "sl_zs": (np.float32, 1), "sl_gammas": (np.float32, 1), "sl_cohesion": (np.float32, 1),
    "sl_frictionangle": (np.float32, 1), "sl_radslope": (np.float32, 1),
    "flood_w": (np.float32, 1), "flood_d50": (np.float32, 1), "flood_slope": (np.float32, 1),
    "flood_sections": (np.float32, -1), "flood_sec_cells": (np.float32, -1),      # (flood_sec_tam, N)
    # module arrays the hot path does not read but f2py exposes (modelsmodule.c:3721-3856): kept so that assignments and reads behave
    "flood_hand": (np.float32, 1), "flood_aquien": (np.int32, 1), "flood_eval": (np.int32, 1), "flood_loc_hand": (np.float32, 1),
    "idevento": (np.int32, 0), "posevento": (np.int32, 0), "posconv": (np.int32, 0), "posstra": (np.int32, 0),
    "speed_map": (np.float32, -1), "fluxes": (np.float32, 3), "vfluxes": (np.float32, 4), "ero": (np.float32, 0), "dep": (np.float32, 0),
    "speed_type": (np.int32, 0), "wi": (np.float32, 0), "diametro": (np.float32, 0),
}
_SCALARS_F = {"g": 9.8, "qskr": 0.0, "qlin_sed": 0.0, "flood_sec_tam": 0.0, "flood_rdf": 0.0, "flood_area": 0.0, "flood_diff": 0.0, "flood_dw": 1000.0, "flood_dsed": 2600.0, "flood_av": 1.0 / 200.0, "flood_cmax": 0.75, "flood_umbral": 3.0,
              "flood_step": 1.0, "flood_hmax": 5.0, "dt": 60.0, "dxp": 1.0, "retorno_gr": 0.0, "retorno_aq": 0.0, "storage_constant": 0.0,
              "sed_factor": 1.0, "sl_fs": 1.0, "sl_gammaw": 9.8, "xll": 0.0, "yll": 0.0, "nodata": -9999.0, "dx": 1.0}
_SCALARS_I = {"flood_max_iter": 10, "nceldas": 0, "ncols": 0, "nrows": 0, "verbose": 0, "rain_first_point": 1, "calc_niter": 5,
              "sim_sediments": 0, "sim_slides": 0, "sim_floods": 0, "save_storage": 0, "save_speed": 0,
              "save_retorno": 0, "save_vfluxes": 0, "save_rc": 0, "show_storage": 0, "show_speed": 0,
              "show_mean_speed": 0, "show_mean_retorno": 0, "show_area": 0, "separate_fluxes": 0,
              "separate_rain": 0, "sl_gullienogullie": 0}
# module arrays shia_v1 allocates and fills (read back by wmf.py:4548-4601); None until a run has produced them
_RESULTS = {"retorned", "mean_rain", "acum_rain", "mean_storage", "mean_speed", "mean_retorno", "mean_vfluxes", "rc_coef",
            "storage_conv", "storage_stra", "volero", "voldepo", "erot", "dept", "sl_riskvector", "sl_slideocurrence",
            "sl_slideacumulate", "sl_slidencelltime", "sl_zcrit", "sl_zmin", "sl_zmax", "sl_bo", "flood_q", "flood_qsed",
            "flood_h", "flood_flood", "flood_speed", "flood_ufr", "flood_cr", "flood_profundidad"}
# switches whose sub-models are outside the hot path built here
_UNSUPPORTED = ()


def read_rain_hdr(ruta_hdr: str, n_reg: int, rain_first_point: int = 1):
    """rain_read_ascii_table, modelosv2.f90:966-1002 -> (idEvento, posEvento)."""
    with open(ruta_hdr) as f:
        lines = f.read().splitlines()
    ntotal = int(lines[2].replace(":", " ").split()[-1])
    if n_reg + rain_first_point > ntotal + 1:
        raise ValueError(f"rain header {ruta_hdr}: {n_reg} intervals from point {rain_first_point} exceed {ntotal} records")
    body = lines[6:]
    if rain_first_point > 1:          # the Fortran skips rain_first_point lines, not rain_first_point-1 (:989-993)
        body = body[rain_first_point:]
    ids, pos = [], []
    for ln in body[:n_reg]:
        parts = ln.replace("\t", " ").split(",")
        ids.append(int(parts[0]))
        pos.append(int(parts[1]))
    if len(pos) < n_reg:
        raise ValueError(f"rain header {ruta_hdr}: only {len(pos)} rows for {n_reg} intervals")
    return np.array(ids, np.int32), np.array(pos, np.int32)


def read_int_basin(ruta: str, record: int, n_cel: int):
    """vect,res = read_int_basin(ruta,record,n_cel)   -- modelosv2.f90:911-926 (direct access, RECL=4*N)."""
    try:
        with open(ruta, "rb") as f:
            f.seek(4 * n_cel * (int(record) - 1))
            v = np.fromfile(f, dtype=np.int32, count=n_cel)
        return (v, 0) if v.size == n_cel else (np.zeros(n_cel, np.int32), 1)
    except OSError:
        return np.zeros(n_cel, np.int32), 1


def read_float_basin_ncol(ruta: str, record: int, n_cel: int, n_col: int):
    """vect,res = read_float_basin_ncol(ruta,record,n_cel,n_col)   -- modelosv2.f90:893-909."""
    try:
        with open(ruta, "rb") as f:
            f.seek(4 * n_cel * n_col * (int(record) - 1))
            v = np.fromfile(f, dtype=np.float32, count=n_cel * n_col)
        if v.size != n_cel * n_col:
            return np.zeros((n_col, n_cel), np.float32, order="F"), 1
        return np.asfortranarray(v.reshape((n_col, n_cel), order="F")), 0
    except OSError:
        return np.zeros((n_col, n_cel), np.float32, order="F"), 1


def read_float_basin(ruta: str, record: int, n_cel: int):
    """vect,res = read_float_basin(ruta,record,n_cel)   -- modelosv2.f90:875-890 (one value per cell)."""
    v, res = read_float_basin_ncol(ruta, record, n_cel, 1)
    return np.ascontiguousarray(v.reshape(-1)), res


def write_int_basin(ruta: str, vect, record: int, n_cel=None, n_col=None):
    """write_int_basin(ruta,vect,record,[n_cel,n_col])   -- modelosv2.f90:945-959."""
    a = np.asfortranarray(vect, dtype=np.int32)
    mode = "wb" if int(record) == 1 or not os.path.exists(ruta) else "r+b"
    with open(ruta, mode) as f:
        f.seek(4 * a.size * (int(record) - 1))
        f.write(a.tobytes(order="F"))


def write_float_basin(ruta: str, vect, record: int, n_cel=None, n_col=None):
    """write_float_basin(ruta,vect,record,[n_cel,n_col])   -- modelosv2.f90:928-944 (record <= 0 writes nothing)."""
    if int(record) <= 0:
        return
    a = np.asfortranarray(vect, dtype=np.float32)
    mode = "wb" if int(record) == 1 or not os.path.exists(ruta) else "r+b"
    with open(ruta, mode) as f:
        f.seek(4 * a.size * (int(record) - 1))
        f.write(a.tobytes(order="F"))


def write_rain_files(path_bin: str, path_hdr: str, rain, pos_evento, dates=None, n_hills=0):
    """Write a rain ``.bin``/``.hdr`` pair in the reference's format (wmf.py:3349-3362): direct-access records
    of N int32 (mm*1000), record 1 all zeros; six header lines then ``id, record, mean, date`` rows."""
    rain = np.ascontiguousarray(rain, dtype=np.int32)
    pos = np.asarray(pos_evento, dtype=np.int64)
    with open(path_bin, "wb") as f:
        f.write(rain.tobytes())
    with open(path_hdr, "w") as f:
        f.write("Numero de celdas: %d \n" % rain.shape[1])
        f.write("Numero de laderas: %d \n" % n_hills)
        f.write("Numero de registros: %d \n" % pos.size)
        f.write("Numero de campos no cero: %d \n" % pos.max())
        f.write("Tipo de interpolacion: TIN\n")
        f.write("IDfecha, Record, Lluvia, Fecha \n")
        for c, p in enumerate(pos):
            mean = rain[p - 1].mean() / 1000.0
            d = dates[c] if dates is not None else "2000-01-01-00:00"
            f.write("%d, \t %d, \t %.2f, %s \n" % (c + 1, p, mean, d))


def _pinned_empty(shape, dtype, order="F"):
    """ndarray backed by page-locked host memory when a CUDA device is present (torch's caching pinned allocator), so
    that the library's host<->device staging copies are true DMA transfers; plain numpy otherwise."""
    n = int(np.prod(shape))
    if n * np.dtype(dtype).itemsize >= (64 << 10):
        try:
            import torch
            if torch.cuda.is_available():
                tdt = {np.dtype(np.float32): torch.float32, np.dtype(np.int32): torch.int32}[np.dtype(dtype)]
                return torch.empty(n, dtype=tdt, pin_memory=True).numpy().reshape(shape, order=order)
        except Exception:
            pass
    return np.empty(shape, dtype=dtype, order=order)


class ModelsModule:
    """Drop-in for the Fortran module object ``models``."""

    def __init__(self, ctx: _lib.Context | None = None):
        d = object.__getattribute__(self, "__dict__")
        d["_ctx"] = ctx
        d["_a"] = {}                       # persistent arrays
        d["_s"] = dict(_SCALARS_F)
        d["_s"].update(_SCALARS_I)
        d["_rain_mem"] = None
        d["_rain_types_mem"] = None
        d["_a"]["speed_type"] = np.ones(3, np.int32)
        d["_a"]["wi"] = np.zeros(3, np.float32)
        d["_a"]["diametro"] = np.zeros(3, np.float32)

    # ---- f2py attribute semantics ------------------------------------------------------------
    def __getattr__(self, name):
        d = object.__getattribute__(self, "__dict__")
        if name in d["_a"]:
            return d["_a"][name]
        if name in d["_s"]:
            v = d["_s"][name]
            return np.int32(v) if name in _SCALARS_I else np.float32(v)
        if name in _ARRAYS:
            return None                    # f2py: unallocated allocatable reads as None
        if name in _RESULTS:
            return None                    # a result array no run has produced yet: unallocated
        if name in ("rute_speed", "rute_storage"):
            return ""                      # character*500 module variables nothing in the hot path reads
        raise AttributeError(name)

    def __setattr__(self, name, value):
        d = object.__getattribute__(self, "__dict__")
        if name in _ARRAYS:
            dtype, K = _ARRAYS[name]
            if value is None:
                d["_a"].pop(name, None)
                return
            src = np.asarray(value)
            arr = _pinned_empty(src.shape, dtype)          # f2py copies and casts on assignment
            arr[...] = src
            if K == 0:
                arr = arr.reshape(-1)
                if name in ("speed_type", "wi", "diametro") and arr.size != 3:
                    raise ValueError(f"models.{name} has 3 elements")
            else:
                if arr.ndim == 1:
                    arr = np.asfortranarray(arr.reshape(1, -1))
                if arr.ndim != 2:
                    raise ValueError(f"models.{name} must be a rank-2 array")
            d["_a"][name] = arr
        elif name in _SCALARS_I:
            d["_s"][name] = int(value)
        elif name in _SCALARS_F:
            d["_s"][name] = float(np.float32(value))
        else:
            d[name] = value

    @property
    def ctx(self) -> _lib.Context:
        d = object.__getattribute__(self, "__dict__")
        if d["_ctx"] is None:
            d["_ctx"] = _lib.default_context()
        return d["_ctx"]

    # ---- helpers kept from the f2py export list ----------------------------------------------------
    read_int_basin = staticmethod(read_int_basin)
    read_float_basin_ncol = staticmethod(read_float_basin_ncol)
    read_float_basin = staticmethod(read_float_basin)
    write_int_basin = staticmethod(write_int_basin)
    write_float_basin = staticmethod(write_float_basin)

    def calc_speed(self, sm, coef, expo, elem_long, speed):
        """area = calc_speed(sm,coef,expo,elem_long,speed)   -- modelosv2.f90:1278-1293.
        f2py returns only ``area`` (speed is intent(inout) on a scalar); both are returned here."""
        c = self.ctx
        s, a = C.c_float(float(speed)), C.c_float()
        c.check(c.L.wmf_calc_speed(c.h, float(sm), float(coef), float(expo), float(elem_long), float(self.dt),
                                   int(self.calc_niter), C.byref(s), C.byref(a)))
        return a.value, s.value

    def set_rain(self, rain, pos_evento):
        """Extension: hold the rainfall in memory instead of the ``.bin/.hdr`` pair.  ``rain`` is
        (n_records,N) int32 mm*1000 (numpy or torch CUDA), record 1 must be the all-zero record,
        ``pos_evento`` (n_reg) 1-based.  ``shia_v1(None, None, ...)`` then uses it."""
        object.__getattribute__(self, "__dict__")["_rain_mem"] = (rain, np.ascontiguousarray(pos_evento, np.int32))

    def set_rain_types(self, conv, stra, pos_conv, pos_stra):
        """Extension (separate_rain = 1): hold the convective / stratiform maps in memory instead of the
        ``ruta_binConv/ruta_hdrConv/ruta_binStra/ruta_hdrStra`` files.  ``conv`` / ``stra`` are (n_records,N) int32 like the
        rain table, ``pos_conv`` / ``pos_stra`` (n_reg) 1-based record numbers (modelosv2.f90:452-456)."""
        object.__getattribute__(self, "__dict__")["_rain_types_mem"] = (
            np.ascontiguousarray(conv, np.int32), np.ascontiguousarray(stra, np.int32),
            np.ascontiguousarray(pos_conv, np.int32), np.ascontiguousarray(pos_stra, np.int32))

    def _load_rain_types(self, ruta_binconv, ruta_binstra, ruta_hdrconv, ruta_hdrstra, n_reg, n_cel):
        """posConv / posStra from the two headers (rain_read_ascii_table_separate, modelosv2.f90:1004-1065: the ``id, record,
        ...`` rows of the rain header format) and the records they name from the two direct-access binaries."""
        d = object.__getattribute__(self, "__dict__")
        if not ruta_binconv and d["_rain_types_mem"] is not None:
            conv, stra, pc, ps = d["_rain_types_mem"]
            if pc.size < n_reg or ps.size < n_reg:
                raise ValueError("set_rain_types: pos_conv / pos_stra shorter than n_reg")
            return conv, stra, pc[:n_reg].copy(), ps[:n_reg].copy()
        if not (ruta_binconv and ruta_binstra and ruta_hdrconv and ruta_hdrstra):
            raise ValueError("models.separate_rain = 1 needs ruta_binconv, ruta_binstra, ruta_hdrconv, ruta_hdrstra "
                             "(modelosv2.f90:346-359) or set_rain_types()")
        out = []
        for rb, rh in ((ruta_binconv, ruta_hdrconv), (ruta_binstra, ruta_hdrstra)):
            # the separate reader wants Nintervals + rain_first_point <= Ntotal (:1026), one row more than the rain reader
            _, pos = read_rain_hdr(str(rh).strip(), n_reg + 1, int(self.rain_first_point))
            pos = pos[:n_reg]
            n_rec = os.path.getsize(str(rb).strip()) // (4 * n_cel)
            table = np.fromfile(str(rb).strip(), dtype=np.int32, count=n_rec * n_cel).reshape(n_rec, n_cel)
            out.append((table, pos))
        return out[0][0], out[1][0], out[0][1], out[1][1]

    def make_resident(self, enable: bool = True):
        """Extension: upload every allocated module array to the GPU once (torch owns the buffers) so that repeated
        runs -- calibration ensembles, benchmarks -- read their static maps in place.  Call again after changing an
        array; ``make_resident(False)`` drops the device copies."""
        d = object.__getattribute__(self, "__dict__")
        if not enable:
            d["_dev"] = {}
            return
        import torch
        dev = {}
        for name, arr in d["_a"].items():
            if isinstance(arr, np.ndarray) and arr.size:
                flat = np.ascontiguousarray(arr.ravel(order="F"))
                # keyed by attribute name and holding the very ndarray the copy was made from: a device copy is used only
                # while models.<name> still IS that array (assignment replaces the ndarray, see __setattr__)
                dev[name] = (arr, torch.from_numpy(flat).to(f"cuda:{self.ctx.device}"))
        d["_dev"] = dev

    def _load_rain(self, ruta_bin, ruta_hdr, n_reg, n_cel):
        d = object.__getattribute__(self, "__dict__")
        if ruta_bin is None and d["_rain_mem"] is not None:
            rain, pos = d["_rain_mem"]
            if pos.size < n_reg:
                raise ValueError("set_rain: pos_evento shorter than n_reg")
            return rain, pos[:n_reg].copy()
        ruta_bin, ruta_hdr = str(ruta_bin).strip(), str(ruta_hdr).strip()
        _, pos = read_rain_hdr(ruta_hdr, n_reg, int(self.rain_first_point))
        uniq = np.unique(pos[pos != 1])
        table = np.zeros((uniq.size + 1, n_cel), np.int32)       # row 0 = the dry record
        with open(ruta_bin, "rb") as f:
            for k, r in enumerate(uniq):
                f.seek(4 * n_cel * (int(r) - 1))
                rec = np.fromfile(f, dtype=np.int32, count=n_cel)
                if rec.size != n_cel:
                    raise ValueError(f"rain binary {ruta_bin}: record {r} is out of range")
                table[k + 1] = rec
        remap = np.ones(n_reg, np.int32)
        wet = pos != 1
        remap[wet] = (np.searchsorted(uniq, pos[wet]) + 2).astype(np.int32)
        return table, remap

    # ---- rainfall interpolation -----------------------------------------------------------------
    def _rain_interp(self, which, xy_basin, coord, rain, extra, ruta, umbral, maskvector, resident, nhills=0):
        c = self.ctx
        xy = np.asfortranarray(xy_basin, dtype=np.float32)
        co = np.asfortranarray(coord, dtype=np.float32)
        r = np.asfortranarray(rain, dtype=np.float32)
        if xy.ndim != 2 or xy.shape[0] != 2 or co.ndim != 2 or co.shape[0] != 2 or r.ndim != 2 or r.shape[0] != co.shape[1]:
            raise ValueError("xy_basin must be (2,nceldas), coord (2,ncoord), rain (ncoord,nreg)")
        N, ncoord, T = xy.shape[1], co.shape[1], r.shape[1]
        mask, n_hills = None, 0
        if maskvector is not None and int(np.asarray(maskvector).sum()) != N:      # the hills model (:1210-1214)
            if which != "idw":
                raise NotImplementedError("rain_mit by hills: the reference takes that branch only when sum(maskVector) == nhills "
                                          "(modelosv2.f90:1126), which no hill table satisfies")
            mask = np.ascontiguousarray(np.asarray(maskvector).reshape(-1), np.int32)
            n_hills = int(nhills)
            if mask.size != N or n_hills < 1 or mask.min() < 0 or mask.max() > n_hills:
                raise ValueError("maskVector must give the hill (1..nhills) of every cell")
        width = n_hills if n_hills else N
        mean, pos = _pinned_empty((T,), np.float32), _pinned_empty((T,), np.int32)
        nrec = C.c_int32(0)

        def call(table_ptr, cap):
            if which == "idw":
                return c.L.wmf_rain_idw(c.h, ptr(xy), ptr(co), ptr(r), float(extra), N, ncoord, T, float(umbral), table_ptr, cap,
                                        ptr(mean), ptr(pos), C.byref(nrec), ptr(mask), n_hills)
            tin, perte = extra
            return c.L.wmf_rain_mit(c.h, ptr(xy), ptr(co), ptr(r), ptr(tin), ptr(perte), N, ncoord, tin.shape[1], T,
                                    float(umbral), table_ptr, cap, ptr(mean), ptr(pos), C.byref(nrec))

        # rows needed while interpolating: record 1 + every step with a positive reading (the others are dry by construction)
        cap = 1 + (int(np.count_nonzero((r > 0).any(axis=0))) if float(umbral) >= 0 else T)
        if resident:
            import torch
            table = torch.empty((cap, width), dtype=torch.int32, device=f"cuda:{c.device}")
            c.check(call(ptr(table), cap))
            self.set_rain(table[:int(nrec.value)], pos)
        else:
            table = _pinned_empty((width, cap), np.int32)             # F-order (values, records) = one record after the other
            c.check(call(ptr(table), cap))
            if ruta:
                with open(str(ruta).strip(), "wb") as f:
                    f.write(table[:, :int(nrec.value)].tobytes(order="F"))
        return mean, pos

    def rain_pre_mit(self, xy_basin, tin, coord, nceldas=None, ntin=None, ncoord=None):
        """tin_perte = rain_pre_mit(xy_basin,tin,coord,[nceldas,ntin,ncoord])   -- modelosv2.f90:1068-1101"""
        c = self.ctx
        xy, co = np.asfortranarray(xy_basin, dtype=np.float32), np.asfortranarray(coord, dtype=np.float32)
        tn = np.asfortranarray(tin, dtype=np.int32)
        if xy.ndim != 2 or xy.shape[0] != 2 or co.ndim != 2 or co.shape[0] != 2 or tn.ndim != 2 or tn.shape[0] != 3:
            raise ValueError("xy_basin must be (2,nceldas), tin (3,ntin), coord (2,ncoord)")
        out = np.zeros((1, xy.shape[1]), np.int32, order="F")
        c.check(c.L.wmf_rain_pre_mit(c.h, ptr(xy), ptr(tn), ptr(co), xy.shape[1], tn.shape[1], co.shape[1], ptr(out)))
        return out

    def rain_idw(self, xy_basin, coord, rain, pp, nhills, ruta, umbral, maskvector=None, nceldas=None, ncoord=None, nreg=None,
                 resident=False):
        """meanrain,posids = rain_idw(xy_basin,coord,rain,pp,nhills,ruta,umbral,maskvector,[nceldas,ncoord,nreg])
        -- modelosv2.f90:1195-1271 (call site wmf.py:3408-3412).  The wet fields are written to ``ruta`` as the Fortran
        does (record 1 = zeros).  Extension: ``resident=True`` keeps the record table on the GPU and installs it with
        ``set_rain`` so that ``shia_v1(None, None, ...)`` runs on it without a file or a PCIe round trip."""
        return self._rain_interp("idw", xy_basin, coord, rain, pp, ruta, umbral, maskvector, resident, nhills)

    def rain_mit(self, xy_basin, coord, rain, tin, tin_perte, nhills, ruta, umbral, maskvector=None, nceldas=None, ncoord=None,
                 ntin=None, nreg=None, resident=False):
        """meanrain,posids = rain_mit(xy_basin,coord,rain,tin,tin_perte,nhills,ruta,umbral,maskvector,[nceldas,ncoord,ntin,nreg])
        -- modelosv2.f90:1104-1194 (call site wmf.py:3335-3347)."""
        tin = np.asfortranarray(tin, dtype=np.int32)
        perte = np.ascontiguousarray(np.asarray(tin_perte).reshape(-1), np.int32)
        if tin.ndim != 2 or tin.shape[0] != 3:
            raise ValueError("tin must be (3,ntin)")
        if perte.min() < 1 or perte.max() > tin.shape[1]:
            raise ValueError("tin_perte must name a triangle (1..ntin) for every cell (wmf.py:3329)")
        return self._rain_interp("mit", xy_basin, coord, rain, (tin, perte), ruta, umbral, maskvector, resident)

    # ---- the model ------------------------------------------------------------------------------
    def shia_v1(self, ruta_bin, ruta_hdr, calib, n_cont, n_conth, n_reg, stoin=None, hspeedin=None, n_cel=None,
                ruta_storage=None, ruta_speed=None, ruta_vfluxes=None, ruta_binconv=None, ruta_binstra=None,
                ruta_hdrconv=None, ruta_hdrstra=None, ruta_retorno=None, ruta_rc=None):
        """q,qsed,qseparated,hum,st1,st3,balance,speed,areacontrol,stoout,qsep_byrain =
        shia_v1(ruta_bin,ruta_hdr,calib,n_cont,n_conth,n_reg,[stoin,hspeedin,n_cel,...])
        -- modelosv2.f90:191-869, f2py signature modelsmodule.c:375-408."""
        for k in _UNSUPPORTED:
            if int(getattr(self, k)) == 1:
                raise NotImplementedError(f"models.{k} = 1 selects a sub-model that is outside the B200 hot path "
                                          "(SURVEY.md 8f); set it to 0")
        extra = self._switch_passes()
        if extra:
            return self._shia_v1_passes(extra, (ruta_bin, ruta_hdr, calib, n_cont, n_conth, n_reg, stoin, hspeedin, n_cel,
                                                ruta_storage, ruta_speed, ruta_vfluxes, ruta_binconv, ruta_binstra,
                                                ruta_hdrconv, ruta_hdrstra, ruta_retorno, ruta_rc))
        a = object.__getattribute__(self, "__dict__")["_a"]
        if "drena" not in a:
            raise ValueError("models.drena is not allocated")
        N = int(n_cel) if n_cel is not None else a["drena"].shape[1]
        if a["drena"].shape[1] != N:
            raise ValueError(f"n_cel={N} but models.drena has {a['drena'].shape[1]} cells")
        T = int(n_reg)
        calib = np.ascontiguousarray(np.asarray(calib, dtype=np.float32).reshape(-1))
        if calib.size != 11:
            raise ValueError("calib must have 11 elements (modelosv2.f90:201)")
        rain, pos = self._load_rain(ruta_bin, ruta_hdr, T, N)
        save_sto, save_spd = int(self.save_storage) == 1, int(self.save_speed) == 1
        if save_sto and not ruta_storage:
            raise ValueError("models.save_storage = 1 needs ruta_storage (modelosv2.f90:799-803)")
        if save_spd and not ruta_speed:
            raise ValueError("models.save_speed = 1 needs ruta_speed (modelosv2.f90:815-818)")
        save_vfl, save_rc, save_ret = int(self.save_vfluxes) == 1, int(self.save_rc) == 1, int(self.save_retorno) == 1
        if save_vfl and not ruta_vfluxes:
            raise ValueError("models.save_vfluxes = 1 needs ruta_vfluxes (modelosv2.f90:806-813)")
        if save_rc and not ruta_rc:
            raise ValueError("models.save_rc = 1 needs ruta_rc (modelosv2.f90:863-867)")
        if save_ret and not ruta_retorno:
            raise ValueError("models.save_retorno = 1 needs ruta_retorno (modelosv2.f90:820-823)")
        types = None
        if int(self.separate_rain) == 1:
            types = self._load_rain_types(ruta_binconv, ruta_binstra, ruta_hdrconv, ruta_hdrstra, T, N)
        out = self._run(calib[None, :], int(n_cont), int(n_conth), T, rain, pos, stoin, hspeedin, rain_types=types)
        a["posevento"] = np.array(pos, np.int32, copy=True)          # module posEvento as rain_read_ascii_table leaves it
        if types is not None:
            a["posconv"], a["posstra"] = np.array(types[2], np.int32, copy=True), np.array(types[3], np.int32, copy=True)
        # state dumps in the reference's direct-access format: record guarda_cond(t) of ruta_storage holds StoOut(5,N)
        # after step t; record t of ruta_speed holds hspeed(4,N)
        if save_sto:
            g = np.asarray(a["guarda_cond"]).reshape(-1)[:T]
            for k, t in enumerate(np.nonzero(g > 0)[0]):
                write_float_basin(str(ruta_storage).strip(), out["sto_snap"][:, :, k], int(g[t]), N, 5)
        if save_spd:
            for t in range(T):
                write_float_basin(str(ruta_speed).strip(), out["speed_snap"][:, :, t], t + 1, N, 4)
        # record guarda_vfluxes(t) of ruta_vfluxes holds vfluxes(4,N) of step t (:811-813); record t of ruta_retorno the
        # return flow of step t (:820-823); record 1 of ruta_rc the two runoff-coefficient accumulators (:863-867)
        if save_vfl and "vfl_snap" in out:
            g = self._guarda("guarda_vfluxes", T)
            for k, t in enumerate(np.nonzero(g > 0)[0]):
                write_float_basin(str(ruta_vfluxes).strip(), out["vfl_snap"][:, :, k], int(g[t]), N, 4)
        if save_ret:
            for t in range(T):
                write_float_basin(str(ruta_retorno).strip(), out["ret_snap"][:, t][None, :], t + 1, N, 1)
        if save_rc:
            write_float_basin(str(ruta_rc).strip(), out["rc_coef"], 1, N, 2)
        nc, nh = int(n_cont), int(n_conth)
        zeros = lambda *s: np.zeros(s, np.float32, order="F")
        qsed = out.get("qsed", zeros(nc, 3, T))
        return (out["q"], qsed, out.get("qseparated", zeros(nc, 3, T)), out["hum"], out["st1"], out["st3"], out["balance"], out["speed"],
                out["areacontrol"], out["stoout"], out.get("qsep_byrain", zeros(nc, 2, T)))

    # switches one launch sequence of wmf_shia_run carries together ("Switch combinations")
    _BASE_SWITCHES = ("sim_sediments", "sim_slides", "save_storage", "save_speed", "save_vfluxes", "save_rc", "save_retorno",
                      "show_storage", "show_mean_speed", "show_mean_retorno")
    _SOLO_SWITCHES = ("separate_fluxes", "separate_rain", "sim_floods")

    def _switch_passes(self):
        """The reference runs any combination of its switches in one ``shia_v1`` call (wmf.py:4634-4637 returns the
        separated fluxes AND the sediment series of the same run).  One ``wmf_shia_run`` carries sediments, landslides,
        every ``save_*`` dump and the ``show_*`` means together, but ``separate_fluxes``, ``separate_rain`` and
        ``sim_floods`` each have a kernel variant of their own.  All of them are diagnostics computed FROM the hydrology --
        none feeds back into the tanks or the routing -- so a combination is the union of the outputs of runs over the same
        hydrology: the first pass runs the base group, every further switch gets a pass of its own and contributes only the
        outputs it owns.  Returns the switches that need such an extra pass ([] = one pass does it)."""
        on = lambda k: int(getattr(self, k)) == 1
        solo = [k for k in self._SOLO_SWITCHES if on(k)]
        base = [k for k in self._BASE_SWITCHES if on(k) and not (k == "show_mean_retorno" and float(self.retorno_gr) != 1.0)]
        if len(solo) + (1 if base else 0) <= 1:
            return []
        return solo if base else solo[1:]

    def _shia_v1_passes(self, extra, args):
        """shia_v1 for a combination of switches no single kernel variant carries: see ``_switch_passes``."""
        saved = {k: int(getattr(self, k)) for k in self._BASE_SWITCHES + self._SOLO_SWITCHES}
        try:
            for k in extra:
                setattr(self, k, 0)
            ret = list(self.shia_v1(*args))                       # base group (or the first solo switch): Q, storages, dumps, ...
            for k in saved:
                setattr(self, k, 0)
            for k in extra:                                       # one pass per remaining switch, everything else off
                setattr(self, k, 1)
                r = self.shia_v1(*args)
                setattr(self, k, 0)
                if k == "separate_fluxes":
                    ret[2] = r[2]                                 # Qseparated
                elif k == "separate_rain":
                    ret[10] = r[10]                               # Qsep_byrain (Storage_conv / Storage_stra stay in the module)
        finally:
            for k, v in saved.items():
                setattr(self, k, v)
        return tuple(ret)

    def _guarda(self, name, T):
        """guarda_vfluxes as shia_v1 sees it: allocated with >= n_reg entries, else "all zeros" (modelosv2.f90:399-409)."""
        g = object.__getattribute__(self, "__dict__")["_a"].get(name)
        if g is None or g.size < T:
            return np.zeros(T, np.int32)
        return np.ascontiguousarray(g.reshape(-1)[:T], np.int32)

    def _cell(self, name, K, N, dtype, required=True):
        a = object.__getattribute__(self, "__dict__")["_a"]
        if name not in a:
            if required:
                raise ValueError(f"models.{name} is not allocated")
            return None
        arr = a[name]
        if arr.shape != (K, N):
            raise ValueError(f"models.{name} has shape {arr.shape}, expected ({K},{N})")
        return arr

    def _run(self, calibs, n_cont, n_conth, T, rain, pos, stoin=None, hspeedin=None, device_out=False, want=None,
             rain_types=None, ensemble=False):
        """Marshal module state into one wmf_shia_run call.  ``calibs`` is (M,11).  ``ensemble=True`` runs the plain
        hydrological configuration an ensemble launch uses (dumps, separations, floods off; outputs laid out with a
        member axis) whatever M is, so that a population evaluates the same model in every chunk size."""
        c = self.ctx
        a = object.__getattribute__(self, "__dict__")["_a"]
        N = a["drena"].shape[1]
        M = calibs.shape[0]
        cfg = ShiaCfg()
        cfg.n_cel, cfg.n_reg, cfg.n_cont, cfg.n_conth = N, T, n_cont, n_conth
        cfg.dt, cfg.dxp = float(self.dt), float(self.dxp)
        cfg.retorno_gr, cfg.retorno_aq = float(self.retorno_gr), float(self.retorno_aq)
        cfg.storage_constant = float(self.storage_constant)
        cfg.speed_type = (C.c_int32 * 3)(*[int(v) for v in a["speed_type"]])
        cfg.calc_niter = int(self.calc_niter)
        for k in ("sim_sediments", "sim_slides", "show_storage", "show_speed", "show_mean_speed",
                  "show_mean_retorno", "save_retorno", "sl_gullienogullie"):
            setattr(cfg, k, int(getattr(self, k)))
        cfg.sed_factor = float(self.sed_factor)
        cfg.wi = (C.c_float * 3)(*[float(v) for v in a["wi"]])
        cfg.diametro = (C.c_float * 3)(*[float(v) for v in a["diametro"]])
        cfg.sl_fs, cfg.sl_gammaw = float(self.sl_fs), float(self.sl_gammaw)
        cfg.n_members = M
        single = (M == 1) and not ensemble
        cfg.save_storage = int(self.save_storage) if single else 0
        cfg.save_speed = int(self.save_speed) if single else 0
        cfg.separate_fluxes = int(self.separate_fluxes) if single else 0
        cfg.save_vfluxes = int(self.save_vfluxes) if single else 0
        cfg.save_rc = int(self.save_rc) if single else 0
        if not single:
            cfg.save_retorno = 0
            cfg.show_storage = cfg.show_mean_speed = cfg.show_mean_retorno = 0
        cfg.separate_rain = 1 if (rain_types is not None and single) else 0
        cfg.sim_floods = int(self.sim_floods) if single else 0
        keep = []
        si = ShiaInputs()

        dev = object.__getattribute__(self, "__dict__").get("_dev") or {}

        def put(field, arr, name=None):
            # a device-resident copy made by make_resident() is used in place (no host->device traffic); only module
            # arrays are looked up, by name, and only while the module attribute is still the array that was uploaded
            if name is not None and arr is not None:
                ent = dev.get(name)
                if ent is not None and ent[0] is arr:
                    arr = ent[1]
            keep.append(arr)
            setattr(si, field, ptr(arr))

        put("drena", self._cell("drena", 3, N, np.int32), "drena")
        si.drena_stride = 3
        put("unit_type", self._cell("unit_type", 1, N, np.int32), "unit_type")
        for k in ("hill_long", "stream_long", "elem_area", "max_capilar", "max_gravita", "max_aquifer"):
            put(k, self._cell(k, 1, N, np.float32), k)
        for k in ("hill_slope", "stream_slope"):
            put(k, self._cell(k, 1, N, np.float32, required=False), k)
        for k in ("v_coef", "h_coef", "h_exp"):
            put(k, self._cell(k, 4, N, np.float32), k)
        put("storage", self._cell("storage", 5, N, np.float32), "storage")
        put("control", self._cell("control", 1, N, np.int32, required=False), "control")
        put("control_h", self._cell("control_h", 1, N, np.int32, required=False), "control_h")
        evp = a.get("evpserie")
        if evp is None or evp.size < T:
            raise ValueError("models.evpserie must hold n_reg values (wmf.py:4487-4490)")
        put("evpserie", np.ascontiguousarray(evp[:T], np.float32))
        put("calib", np.ascontiguousarray(calibs, np.float32))
        if stoin is not None:
            s = np.asfortranarray(stoin, dtype=np.float32)
            if s.shape != (5, N):
                raise ValueError(f"stoin must be (5,{N})")
            put("stoin", s)
        if hspeedin is not None:
            h = np.asfortranarray(hspeedin, dtype=np.float32)
            if h.shape != (4, N):
                raise ValueError(f"hspeedin must be (4,{N})")
            put("hspeedin", h)
        if is_torch_cuda(rain):
            assert rain.dim() == 2 and rain.shape[1] == N and rain.is_contiguous()
            n_records = rain.shape[0]
        else:
            rain = np.ascontiguousarray(rain, dtype=np.int32)
            if rain.ndim != 2 or rain.shape[1] != N:
                raise ValueError(f"rain table must be (n_records,{N})")
            n_records = rain.shape[0]
        put("rain", rain)
        si.n_records = int(n_records)
        pos = np.ascontiguousarray(pos, np.int32)
        if pos.ndim == 2:                     # storm scenarios: member m follows its own sequence of rain records, pos[m, t]
            if pos.shape != (M, T) or M == 1:
                raise ValueError(f"per-member pos_evento must be ({M},{T}) with more than one member")
            cfg.pos_per_member = 1
        put("pos_evento", pos)
        n_snap = 0
        if cfg.save_storage == 1:
            g = a.get("guarda_cond")
            if g is None or g.size < T:
                raise ValueError("models.guarda_cond must hold n_reg values when save_storage = 1 (wmf.py:4497-4505)")
            g = np.ascontiguousarray(g.reshape(-1)[:T], np.int32)
            n_snap = int(np.count_nonzero(g > 0))
            put("guarda_cond", g)
        if cfg.separate_rain == 1:
            conv, stra, pc, ps = rain_types
            for tab, nm in ((conv, "conv"), (stra, "stra")):
                if tab.ndim != 2 or tab.shape[1] != N:
                    raise ValueError(f"{nm} table must be (n_records,{N})")
            put("conv", conv); put("stra", stra); put("pos_conv", pc); put("pos_stra", ps)
            si.n_conv_records, si.n_stra_records = int(conv.shape[0]), int(stra.shape[0])
        if cfg.sim_floods == 1:              # HydroFlash (modelosv2.f90:728-743, 1711-1822)
            for k in ("flood_dw", "flood_dsed", "flood_av", "flood_cmax", "flood_umbral", "flood_step", "flood_hmax"):
                setattr(cfg, k, float(getattr(self, k)))
            for k in ("flood_w", "flood_d50", "flood_slope"):
                put(k, self._cell(k, 1, N, np.float32), k)
            sec = a.get("flood_sections")
            if sec is None or sec.ndim != 2 or sec.shape[1] != N:
                raise ValueError(f"models.flood_sections must be (flood_sec_tam,{N})")
            cfg.flood_sec_tam = int(sec.shape[0])
            put("flood_sections", sec)
            put("flood_sec_cells", self._cell("flood_sec_cells", sec.shape[0], N, np.float32))
        n_vsnap = 0
        if cfg.save_vfluxes == 1:
            g = self._guarda("guarda_vfluxes", T)
            n_vsnap = int(np.count_nonzero(g > 0))
            put("guarda_vfluxes", g)
        sed, sl = cfg.sim_sediments == 1, cfg.sim_slides == 1
        if sed:
            for k in ("krus", "crus", "prus"):
                put(k, self._cell(k, 1, N, np.float32), k)
            put("parliac", self._cell("parliac", 3, N, np.float32), "parliac")
            if "vd" in a and not getattr(self, "_reset_vd", False):
                put("vd_init", self._cell("vd", 3, N, np.float32))
        if sl:
            for k in ("sl_zs", "sl_gammas", "sl_cohesion", "sl_frictionangle", "sl_radslope"):
                put(k, self._cell(k, 1, N, np.float32), k)

        so = ShiaOutputs()
        res = {}

        def out(name, shape, dtype=np.float32):
            if want is not None and name not in want:
                return
            if device_out:
                import torch
                tdt = torch.float32 if dtype == np.float32 else torch.int32
                arr = torch.empty(tuple(reversed(shape)), dtype=tdt, device=f"cuda:{c.device}")
            else:
                arr = _pinned_empty(shape, dtype)
            res[name] = arr
            setattr(so, name, ptr(arr))

        if single:
            out("q", (n_cont, T)); out("hum", (n_conth, T)); out("st1", (n_conth, T)); out("st3", (n_conth, T))
            out("balance", (T,)); out("speed", (n_cont, T)); out("areacontrol", (n_cont, T))
            out("stoout", (5, N)); out("hspeed_out", (4, N)); out("mean_rain", (1, T)); out("acum_rain", (1, N))
            if cfg.separate_fluxes == 1:
                out("qseparated", (n_cont, 3, T))
            if cfg.separate_rain == 1:
                out("qsep_byrain", (n_cont, 2, T)); out("storage_conv", (5, N)); out("storage_stra", (5, N))
            if cfg.sim_floods == 1:
                out("flood_flood", (1, N), np.int32)
                for k in ("flood_h", "flood_ufr", "flood_cr", "flood_q", "flood_qsed", "flood_speed"):
                    out(k, (1, N))
                out("flood_profundidad", (cfg.flood_sec_tam, N)); out("flood_scalars", (3,))
            if cfg.save_storage == 1:
                out("sto_snap", (5, N, max(n_snap, 1)))
            if cfg.save_speed == 1:
                out("speed_snap", (4, N, T))
            if cfg.save_vfluxes == 1:
                out("mean_vfluxes", (4, T))
                if n_vsnap:
                    out("vfl_snap", (4, N, n_vsnap))
            if cfg.save_rc == 1:
                out("rc_coef", (2, N))
            if cfg.save_retorno == 1:
                out("ret_snap", (N, T))
            if cfg.show_storage:
                out("mean_storage", (5, T))
            if cfg.show_mean_speed:
                out("mean_speed", (4, T))
            if cfg.retorno_gr == 1.0:
                out("retorned", (1, N))
                if cfg.show_mean_retorno:
                    out("mean_retorno", (T,))
            if sed:
                out("qsed", (n_cont, 3, T))
                for k in ("vs", "vd", "vsc", "vdc"):
                    out(k, (3, N))
                out("volero", (N,)); out("voldepo", (N,)); out("erot", (3,)); out("dept", (3,))
            if sl:
                for k in ("sl_zmin", "sl_zcrit", "sl_zmax", "sl_bo", "sl_riskvector", "sl_slideocurrence"):
                    out(k, (1, N))
                out("sl_slideacumulate", (1, N), np.int32)
                out("sl_slidencelltime", (T,), np.int32)
        else:
            out("q", (n_cont, T, M))
            out("stoout", (5, N, M))
            out("hspeed_out", (4, N, M))
            if sed:                           # the member axis comes last (slowest in memory), as for q
                out("qsed", (n_cont, 3, T, M))
                for k in ("vs", "vd", "vsc", "vdc"):
                    out(k, (3, N, M))
                out("volero", (N, M)); out("voldepo", (N, M)); out("erot", (3, M)); out("dept", (3, M))
            if sl:
                for k in ("sl_zmin", "sl_zcrit", "sl_zmax", "sl_bo", "sl_riskvector"):
                    out(k, (1, N))
                out("sl_slideocurrence", (N, M))
                out("sl_slideacumulate", (N, M), np.int32)
                out("sl_slidencelltime", (T, M), np.int32)
        c.check(c.L.wmf_shia_run(c.h, C.byref(cfg), C.byref(si), C.byref(so)))
        # module-global results the reference leaves behind (read back at wmf.py:4548-4601)
        if single and not device_out:
            d = object.__getattribute__(self, "__dict__")
            for k in ("mean_rain", "acum_rain", "mean_storage", "mean_speed", "mean_retorno", "mean_vfluxes", "rc_coef",
                      "storage_conv", "storage_stra", "flood_flood", "flood_h", "flood_ufr", "flood_cr", "flood_q", "flood_qsed",
                      "flood_speed", "flood_profundidad",
                      "retorned", "vs", "vd",
                      "vsc", "vdc", "volero", "voldepo", "sl_zmin", "sl_zcrit", "sl_zmax", "sl_bo", "sl_riskvector",
                      "sl_slideocurrence", "sl_slideacumulate", "sl_slidencelltime"):
                if k in res:
                    if k in _ARRAYS:
                        d["_a"][k] = res[k]
                    else:
                        d[k] = res[k]
            if "erot" in res:
                d["erot"], d["dept"] = res["erot"], res["dept"]
            if "flood_scalars" in res:
                d["_s"]["flood_rdf"], d["_s"]["flood_diff"] = float(res["flood_scalars"][0]), float(res["flood_scalars"][2])
        return res
