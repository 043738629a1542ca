"""THE REFERENCE ITSELF -- test infrastructure, NOT product code.

ctypes front-end of ``oracle/_ref/libwmf_ref.so``: the reference's own compiled Fortran (the COFF objects of its last
f2py build, ``/root/reference/build/temp.win-amd64-3.7/Release/wmf/{cuencas,modelosv2}.o``, gfortran 5.3.0 ``-O3
-funroll-loops``) loaded and relocated by ``oracle/ref_loader.c``.  ``RefCu`` and ``shia_v1`` have the same Python
surface as ``oracle.Cu`` / ``oracle.shia_v1`` so that a test can run the same inputs through the reference, the C
restatement and the CUDA path.  What the f2py wrappers did (set module variables, allocate the module arrays, pass every
argument by reference plus the hidden character lengths) is done here by hand.

``libwmf_ref.so`` can only be (re)built where ``/root/reference`` exists (``make -C oracle ref``); the built file is
git-ignored and travels to the GPU box.  ``available()`` says whether it is there.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libwmf_ref.so")
REF_OBJ_DIR = "/root/reference/build/temp.win-amd64-3.7/Release/wmf"
_L = None
CU, MODELS = 0, 1


def build() -> bool:
    """Build oracle/_ref/libwmf_ref.so when the reference tree is present.  Returns availability."""
    if os.path.exists(os.path.join(REF_OBJ_DIR, "modelosv2.o")):
        subprocess.run(["make", "-C", _HERE, "ref"], check=True, capture_output=True)
    return os.path.exists(LIB_PATH)


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib() -> C.CDLL:
    global _L
    if _L is None:
        if not available() and not build():
            raise RuntimeError("oracle/_ref/libwmf_ref.so is not built (needs /root/reference: make -C oracle ref)")
        L = C.CDLL(LIB_PATH)
        L.wref_error.restype = C.c_char_p
        L.wref_sym.restype = C.c_void_p
        L.wref_sym.argtypes = [C.c_int, C.c_char_p]
        L.wref_call.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint64)]
        L.wref_set_allocatable.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_void_p]
        L.wref_free_allocatable.argtypes = [C.c_void_p]
        L.wref_get_allocatable.restype = C.c_void_p
        L.wref_get_allocatable.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        if L.wref_init() != 0:
            raise RuntimeError("reference loader: " + L.wref_error().decode())
        _L = L
    return _L


def _sym(obj: int, name: str) -> int:
    a = lib().wref_sym(obj, name.encode())
    if not a:
        raise KeyError(name)
    return a


def _modvar(obj: int, name: str) -> int:
    return _sym(obj, ("__cu_MOD_" if obj == CU else "__models_MOD_") + name.lower())


def _call(obj: int, routine: str, *args):
    """Call a module procedure; every argument is an address (or a hidden character length)."""
    vals = []
    for a in args:
        if isinstance(a, np.ndarray):
            vals.append(a.ctypes.data)
        elif isinstance(a, int):
            vals.append(a)
        elif a is None:
            vals.append(0)                      # absent optional argument
        else:
            vals.append(C.addressof(a))
    arr = (C.c_uint64 * len(vals))(*vals)
    if lib().wref_call(_modvar(obj, routine), len(vals), arr) != 0:
        raise RuntimeError("wref_call failed")


def _set_scalar(obj, name, value, ctype):
    ctype.from_address(_modvar(obj, name)).value = value


def _get_scalar(obj, name, ctype):
    return ctype.from_address(_modvar(obj, name)).value


def _set_alloc(obj, name, arr, dtype):
    """``module.name = arr`` for an allocatable of rank arr.ndim (1 or 2), Fortran order, copied."""
    a = np.asfortranarray(arr, dtype=dtype)
    n0, n1 = (a.shape[0], 1) if a.ndim == 1 else a.shape
    flat = np.ascontiguousarray(a.ravel(order="F"))
    lib().wref_set_allocatable(_modvar(obj, name), a.ndim, 1 if dtype == np.int32 else 3, n0, n1, flat.ctypes.data)


def _free_alloc(obj, name):
    lib().wref_free_allocatable(_modvar(obj, name))


def _get_alloc(obj, name, rank, dtype):
    n0, n1 = C.c_int64(), C.c_int64()
    p = lib().wref_get_allocatable(_modvar(obj, name), rank, C.byref(n0), C.byref(n1))
    if not p:
        return None
    count = n0.value * n1.value
    buf = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float if dtype == np.float32 else C.c_int32)), shape=(count,))
    out = np.array(buf, dtype=dtype, copy=True)
    return out if rank == 1 else out.reshape((n0.value, n1.value), order="F")


def _f(a):
    return np.asfortranarray(a, dtype=np.float32)


def _i(a):
    return np.asfortranarray(a, dtype=np.int32)


i32, f32 = C.c_int32, C.c_float


class RefCu:
    """The reference's module ``cu`` (cuencas.f90), same Python surface as ``oracle.Cu``."""

    def __init__(self):
        lib()
        self.ncols = 0
        self.nrows = 0
        self.xll = 0.0
        self.yll = 0.0
        self.dx = 1.0
        self.dy = 1.0
        self.dxp = 1.0
        self.nodata = -9999.0
        self.area = self.pend_media = self.elevacion = self.centrox = self.centroy = 0.0

    def _push(self):
        _set_scalar(CU, "ncols", int(self.ncols), i32)
        _set_scalar(CU, "nrows", int(self.nrows), i32)
        for k in ("xll", "yll", "dx", "dy", "dxp", "nodata"):
            _set_scalar(CU, k, float(getattr(self, k)), f32)

    def _pull(self):
        for k in ("area", "pend_media", "elevacion", "centrox", "centroy"):
            setattr(self, k, _get_scalar(CU, k, f32))

    def coord2fil_col(self, x, y):
        self._push()
        col, fil = i32(), i32()
        _call(CU, "coord2fil_col", f32(x), f32(y), col, fil)
        return col.value, fil.value

    def basin_find(self, x, y, dir_map):
        self._push()
        d = _i(dir_map)
        nc, nr = d.shape
        n = i32()
        # basin_temp is allocated only when it is not yet (cuencas.f90:539): a second call on a LARGER raster overruns
        # it, and basin_cut slices to the end of whatever size it has.  Start from the state of a freshly imported module.
        _free_alloc(CU, "basin_temp")
        _call(CU, "basin_find", f32(x), f32(y), d, i32(nc), i32(nr), n)
        return n.value

    def basin_cut(self, nceldas):
        self._push()
        out = np.zeros((3, nceldas), np.int32, order="F")
        _call(CU, "basin_cut", i32(nceldas), out)
        return out

    def basin_acum(self, basin_f):
        self._push()
        b = _i(basin_f)
        acum = np.zeros(b.shape[1], np.int32)
        _call(CU, "basin_acum", b, i32(b.shape[1]), acum)
        self._pull()
        return acum

    def basin_acum_var(self, drain, var):
        d = np.ascontiguousarray(np.asarray(drain).reshape(-1), np.int32)
        v = np.ascontiguousarray(np.asarray(var).reshape(-1), np.float32)
        out = np.zeros((1, d.size), np.float32, order="F")
        _call(CU, "basin_acum_var", d, i32(d.size), v, out)
        return out

    def basin_basics(self, basin_f, dem, dir_vec):
        self._push()
        b = _i(basin_f)
        n = b.shape[1]
        dem = np.ascontiguousarray(dem, np.float32)
        dv = np.ascontiguousarray(dir_vec, np.int32)
        acum = np.zeros(n, np.int32)
        lng, pend, elev = (np.zeros(n, np.float32) for _ in range(3))
        _call(CU, "basin_basics", b, dem, dv, i32(n), acum, lng, pend, elev)
        self._pull()
        return acum, lng, pend, elev

    def basin_coordxy(self, basin_f):
        self._push()
        b = _i(basin_f)
        n = b.shape[1]
        x, y = np.zeros(n, np.float32), np.zeros(n, np.float32)
        _call(CU, "basin_coordxy", b, x, y, i32(n))
        return x, y

    def basin_map2basin(self, basin_f, mapa, xllm, yllm, dxm, dym, nodatam):
        self._push()
        b = _i(basin_f)
        n = b.shape[1]
        m = _f(mapa)
        vec = np.zeros(n, np.float32)
        _call(CU, "basin_map2basin", b, i32(n), m, f32(xllm), f32(yllm), i32(m.shape[0]), i32(m.shape[1]), f32(dxm),
              f32(dym), f32(nodatam), vec)
        return vec

    def basin_2map_find(self, basin):
        b = _i(basin)
        a, c = i32(), i32()
        _call(CU, "basin_2map_find", b, a, c, i32(b.shape[1]))
        return a.value, c.value

    def basin_2map(self, basin, var, map_ncols, map_nrows):
        self._push()
        b = _i(basin)
        v = np.ascontiguousarray(var, np.float32)
        mapa = np.zeros((map_ncols, map_nrows), np.float32, order="F")
        xll, yll = f32(), f32()
        _call(CU, "basin_2map", b, v, mapa, i32(map_ncols), i32(map_nrows), xll, yll, i32(b.shape[1]))
        return mapa, xll.value, yll.value

    def basin_float_var2map(self, basin_f, vect, nc, nf):
        self._push()
        b = _i(basin_f)
        v = np.ascontiguousarray(vect, np.float32)
        mapa = np.zeros((nc, nf), np.float32, order="F")
        _call(CU, "basin_float_var2map", b, v, mapa, i32(nc), i32(nf), i32(b.shape[1]))
        return mapa

    def basin_stream_nod(self, basin_f, acum, umbral):
        b = _i(basin_f)
        n = b.shape[1]
        a = np.ascontiguousarray(acum, np.int32)
        cauce, nodos, traz = (np.zeros(n, np.int32) for _ in range(3))
        nn, ncau = i32(), i32()
        _call(CU, "basin_stream_nod", b, a, i32(n), i32(int(umbral)), cauce, nodos, traz, nn, ncau)
        return cauce, nodos, traz, nn.value, ncau.value

    def basin_stream_type(self, basin_f, acum, umbrales):
        a = np.ascontiguousarray(acum, np.int32)
        u = np.ascontiguousarray(umbrales, np.int32)
        drain = np.ascontiguousarray(np.asarray(basin_f).reshape(3, -1, order="F")[0] if np.asarray(basin_f).ndim == 2
                                     else basin_f, np.int32)
        out = np.zeros(a.size, np.int32)
        _call(CU, "basin_stream_type", drain, a, u, out, i32(u.size), i32(a.size))
        return out

    def geo_acum_to_cauce(self, acum, umbral):
        a = _i(acum)
        out = np.zeros(a.shape, np.int32, order="F")
        _call(CU, "geo_acum_to_cauce", a, out, i32(int(umbral)), i32(a.shape[0]), i32(a.shape[1]))
        return out

    def geo_hand(self, basin_f, elev, lng, cauce):
        b = _i(basin_f)
        n = b.shape[1]
        e, l = np.ascontiguousarray(elev, np.float32), np.ascontiguousarray(lng, np.float32)
        c = np.ascontiguousarray(cauce, np.int32)
        hand, hdnd, aq = np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.int32)
        _call(CU, "geo_hand", b, e, l, c, i32(n), hand, hdnd, aq)
        return hand, hdnd, aq

    def geo_hand_global(self, dem, dir_map, red):
        self._push()
        d, di, r = _f(dem), _i(dir_map), _i(red)
        nc, nr = d.shape
        hand = np.zeros((nc, nr), np.float32, order="F")
        _call(CU, "geo_hand_global", d, di, r, hand, i32(nc), i32(nr))
        return hand

    def basin_time_to_out(self, basin_f, lng, speed):
        b = _i(basin_f)
        n = b.shape[1]
        l, s = np.ascontiguousarray(lng, np.float32), np.ascontiguousarray(speed, np.float32)
        t = np.zeros(n, np.float32)
        _call(CU, "basin_time_to_out", b, l, s, t, i32(n))
        return t

    def basin_propagate(self, basin_f, var):
        b = _i(basin_f)
        n = b.shape[1]
        v = np.ascontiguousarray(var, np.float32)
        out = np.zeros(n, np.float32)
        _call(CU, "basin_propagate", b, v, out, i32(n))
        return out


    def basin_point2var(self, basin_f, id_coord, xy_coord):
        self._push()
        b = _i(basin_f)
        n = b.shape[1]
        ids, xy = np.ascontiguousarray(id_coord, np.int32), _f(xy_coord)
        res, pts = np.zeros(ids.size, np.int32), np.zeros(n, np.int32)
        _call(CU, "basin_point2var", b, ids, xy, res, pts, i32(ids.size), i32(n))
        return res, pts

    def basin_stream_point2stream(self, basin_f, cauce, id_coord, xy_coord):
        self._push()
        b = _i(basin_f)
        n = b.shape[1]
        ca = np.ascontiguousarray(cauce, np.int32)
        ids, xy = np.ascontiguousarray(id_coord, np.int32), _f(xy_coord)
        res, pts = np.zeros(ids.size, np.int32), np.zeros(n, np.int32)
        xyn = np.zeros((2, ids.size), np.float32, order="F")
        _call(CU, "basin_stream_point2stream", b, ca, ids, xy, res, pts, xyn, i32(ids.size), i32(n))
        return res, pts, xyn

    def basin_stream_sections(self, basin_f, cauces, directions, dem, num_celdas):
        b = _i(basin_f)
        n = b.shape[1]
        ca, di = np.ascontiguousarray(cauces, np.int32), np.ascontiguousarray(directions, np.int32)
        d = _f(dem)
        S = 2 * int(num_celdas) + 1
        sec, cel = np.zeros((S, n), np.float32, order="F"), np.zeros((S, n), np.int32, order="F")
        _call(CU, "basin_stream_sections", b, ca, di, d, i32(int(num_celdas)), i32(n), i32(d.shape[0]), i32(d.shape[1]), sec, cel)
        return sec, cel

    def stream_find_to_corr(self, x, y, dem, dir_map, cauce):
        self._push()
        de, d, ca = _f(dem), _i(dir_map), _i(cauce)
        lat, lon = f32(), f32()
        _call(CU, "stream_find_to_corr", f32(x), f32(y), de, d, ca, i32(d.shape[0]), i32(d.shape[1]), lat, lon)
        return lat.value, lon.value

    # ---- sub-basin (hills) topology, cuencas.f90:1380-1438, 1835-2118 ----
    def basin_stream_slope(self, basin_f, elev, lng, nodos, n_cauce):
        self._push()
        b = _i(basin_f)
        n = b.shape[1]
        e, l = np.ascontiguousarray(elev, np.float32), np.ascontiguousarray(lng, np.float32)
        nd = np.ascontiguousarray(nodos, np.int32)
        s, ln = np.zeros(n, np.float32), np.zeros(n, np.float32)
        _call(CU, "basin_stream_slope", b, e, l, nd, i32(int(n_cauce)), s, ln, i32(n))
        return s, ln

    def basin_subbasin_nod(self, basin_f, acum, umbral):
        b = _i(basin_f)
        n = b.shape[1]
        a = np.ascontiguousarray(acum, np.int32)
        cauce, nodos = np.zeros(n, np.int32), np.zeros(n, np.int32)
        nn = i32()
        _call(CU, "basin_subbasin_nod", b, a, i32(n), i32(int(umbral)), cauce, nodos, nn)
        return cauce, nodos, nn.value

    def basin_subbasin_cut(self, n_nodos):
        out = np.zeros((2, n_nodos), np.int32, order="F")
        _call(CU, "basin_subbasin_cut", i32(n_nodos), out)
        return out

    def basin_subbasin_find(self, basin_f, nodos, n_nodos):
        b = _i(basin_f)
        n = b.shape[1]
        nd = np.ascontiguousarray(nodos, np.int32)
        sub_pert, sub_basin = np.zeros(n, np.int32), np.zeros(n_nodos, np.int32)
        _call(CU, "basin_subbasin_find", b, nd, sub_pert, sub_basin, i32(n_nodos), i32(n))
        return sub_pert, sub_basin

    def basin_subbasin_horton(self, sub_basins, nceldas):
        sb = _i(sub_basins)
        nn = sb.shape[1]
        sh, nh = np.zeros(nn, np.int32), np.zeros(nceldas, np.int32)
        _call(CU, "basin_subbasin_horton", sb, sh, nh, i32(nn), i32(int(nceldas)))
        return sh, nh

    def basin_subbasin_long(self, sub_pert, cauce, lng, sub_basin, sub_horton):
        sp, ca = np.ascontiguousarray(sub_pert, np.int32), np.ascontiguousarray(cauce, np.int32)
        li = np.ascontiguousarray(np.asarray(lng), np.float32).astype(np.int32)        # f2py casts to the INTEGER dummy
        sb, sh = _i(sub_basin), np.ascontiguousarray(sub_horton, np.int32)
        nn = sb.shape[1]
        out = np.zeros(nn, np.float32)
        mx, nmx = f32(), i32()
        _call(CU, "basin_subbasin_long", sp, ca, li, sb, sh, out, mx, nmx, i32(nn), i32(sp.size))
        return out, mx.value, nmx.value

    def basin_subbasin_stream_prop(self, sub_pert, cauce, lng, slope, n_nodos):
        sp, ca = np.ascontiguousarray(sub_pert, np.int32), np.ascontiguousarray(cauce, np.int32)
        l, s = np.ascontiguousarray(lng, np.float32), np.ascontiguousarray(slope, np.float32)
        ss, sl = np.zeros(n_nodos, np.float32), np.zeros(n_nodos, np.float32)
        _call(CU, "basin_subbasin_stream_prop", sp, ca, l, s, ss, sl, i32(sp.size), i32(int(n_nodos)))
        return ss, sl

    def basin_subbasin_map2subbasin(self, sub_pert, var, n_nodos, cauce, sum_mean):
        sp, ca = np.ascontiguousarray(sub_pert, np.int32), np.ascontiguousarray(cauce, np.int32)
        v = np.ascontiguousarray(var, np.float32)
        out = np.zeros(n_nodos, np.float32)
        _call(CU, "basin_subbasin_map2subbasin", sp, v, out, i32(int(n_nodos)), i32(sp.size), ca, i32(int(sum_mean)))
        return out


# ---------------------------------------------------------------------------------------------- module models
_CELL_F = {"hill_long": 1, "hill_slope": 1, "stream_long": 1, "stream_slope": 1, "elem_area": 1, "v_coef": 4, "h_coef": 4,
           "h_exp": 4, "max_capilar": 1, "max_gravita": 1, "max_aquifer": 1, "storage": 5}
_SED_F = {"krus": 1, "crus": 1, "prus": 1, "parliac": 3}
_SL_F = {"sl_zs": 1, "sl_gammas": 1, "sl_cohesion": 1, "sl_frictionangle": 1, "sl_radslope": 1}
_RUN_ALLOCATED = ("vd", "vs", "vsc", "vdc", "volero", "voldepo", "retorned", "mean_retorno", "mean_storage", "mean_speed",
                  "mean_vfluxes", "vfluxes", "rc_coef", "fluxes", "guarda_cond", "guarda_vfluxes")


def _cell2(a, K, N, dtype):
    a = np.asarray(a)
    if a.ndim == 1:
        a = a.reshape(1, -1)
    assert a.shape == (K, N), (a.shape, K, N)
    return np.asfortranarray(a, dtype=dtype)


def _ruta(path, n=500):
    b = (path or "").encode()
    assert len(b) < n
    return np.frombuffer(b + b" " * (n - len(b)), dtype=np.uint8).copy()


def write_rain(path_bin, path_hdr, rain, pos):
    """The rain ``.bin`` / ``.hdr`` pair shia_v1 reads (format: wmf.py:3349-3362, modelosv2.f90:966-1002)."""
    rain = np.ascontiguousarray(rain, np.int32)
    rain.tofile(path_bin)
    with open(path_hdr, "w") as f:
        f.write("Numero de celdas: %d \n" % rain.shape[1])
        f.write("Numero de laderas: %d \n" % 0)
        f.write("Numero de registros: %d \n" % len(pos))
        f.write("Numero de campos no cero: %d \n" % int(np.max(pos)))
        f.write("Tipo de interpolacion: TIN\n")
        f.write("IDfecha, Record, Lluvia, Fecha \n")
        for c, p in enumerate(pos):
            f.write("%d, \t %d, \t %.2f, %s \n" % (c + 1, int(p), 0.0, "2020-01-01-00:00"))


def shia_v1(inp: dict, workdir: str | None = None, dumps: dict | None = None, rain_files: tuple | None = None) -> dict:
    """Run the reference's ``shia_v1`` (modelosv2.f90:191-869) on the input dict ``oracle.shia_v1`` takes.  The rain
    table is written to a ``.bin`` / ``.hdr`` pair the Fortran reads record by record, as in production
    (``rain_files`` = an already written pair, so that a timed call does not include the writing).  ``dumps`` may
    give ``ruta_storage / ruta_speed / ruta_vfluxes / ruta_retorno / ruta_rc`` paths for the save_* switches."""
    drena = np.asarray(inp["drena"])
    drena = np.ascontiguousarray(drena[0] if drena.ndim == 2 else drena, np.int32)
    N, T = drena.size, int(inp["n_reg"])
    control = _cell2(inp.get("control", np.zeros(N)), 1, N, np.int32)
    control_h = _cell2(inp.get("control_h", np.zeros(N)), 1, N, np.int32)
    n_cont = int(np.count_nonzero(control)) + 1
    n_conth = max(1, int(np.count_nonzero(control_h)))
    for k in _RUN_ALLOCATED:
        _free_alloc(MODELS, k)
    d3 = np.zeros((3, N), np.int32, order="F")
    d3[0] = drena
    _set_alloc(MODELS, "drena", d3, np.int32)        # wmf.py:3031 assigns the (3,N) structure; shia_v1 reads row 1
    _set_alloc(MODELS, "unit_type", _cell2(inp["unit_type"], 1, N, np.int32), np.int32)
    for k, K in _CELL_F.items():
        _set_alloc(MODELS, k, _cell2(inp[k], K, N, np.float32), np.float32)
    _set_alloc(MODELS, "evpserie", np.ascontiguousarray(inp.get("evpserie", np.ones(T)), np.float32)[:T], np.float32)
    _set_alloc(MODELS, "control", control, np.int32)
    _set_alloc(MODELS, "control_h", control_h, np.int32)
    for k, ct in (("dt", f32), ("dxp", f32), ("retorno_gr", f32), ("retorno_aq", f32), ("storage_constant", f32)):
        _set_scalar(MODELS, k, float(inp.get(k, {"dxp": 1.0}.get(k, 0.0))), ct)
    _set_scalar(MODELS, "nceldas", N, i32)
    _set_scalar(MODELS, "verbose", 0, i32)
    _set_scalar(MODELS, "rain_first_point", 1, i32)
    _set_scalar(MODELS, "calc_niter", int(inp.get("calc_niter", 5)), i32)
    for k in ("sim_sediments", "sim_slides", "sim_floods", "save_storage", "save_speed", "save_retorno", "save_vfluxes",
              "save_rc", "show_storage", "show_speed", "show_mean_speed", "show_mean_retorno", "show_area",
              "separate_fluxes", "separate_rain"):
        _set_scalar(MODELS, k, int(inp.get(k, 0)), i32)
    (C.c_int32 * 3).from_address(_modvar(MODELS, "speed_type"))[:] = [int(v) for v in inp.get("speed_type", [1, 1, 1])]
    for k in ("guarda_cond", "guarda_vfluxes"):
        if inp.get(k) is not None:
            _set_alloc(MODELS, k, np.ascontiguousarray(inp[k], np.int32)[:T], np.int32)
    sed, sl = int(inp.get("sim_sediments", 0)) == 1, int(inp.get("sim_slides", 0)) == 1
    if sed:
        _set_scalar(MODELS, "sed_factor", float(inp.get("sed_factor", 1.0)), f32)
        (C.c_float * 3).from_address(_modvar(MODELS, "wi"))[:] = [float(v) for v in inp["wi"]]
        (C.c_float * 3).from_address(_modvar(MODELS, "diametro"))[:] = [float(v) for v in inp["diametro"]]
        for k, K in _SED_F.items():
            _set_alloc(MODELS, k, _cell2(inp[k], K, N, np.float32), np.float32)
        if inp.get("vd") is not None:
            _set_alloc(MODELS, "vd", _cell2(inp["vd"], 3, N, np.float32), np.float32)
    if sl:
        _set_scalar(MODELS, "sl_gullienogullie", int(inp.get("sl_gullienogullie", 0)), i32)
        _set_scalar(MODELS, "sl_fs", float(inp.get("sl_fs", 1.0)), f32)
        _set_scalar(MODELS, "sl_gammaw", float(inp.get("sl_gammaw", 9.8)), f32)
        for k, K in _SL_F.items():
            _set_alloc(MODELS, k, _cell2(inp[k], K, N, np.float32), np.float32)

    floods = int(inp.get("sim_floods", 0)) == 1
    _FLOOD_OUT = ("flood_q", "flood_qsed", "flood_speed", "flood_h", "flood_ufr", "flood_cr", "flood_eval", "flood_flood",
                  "flood_loc_hand", "flood_slope", "flood_profundidad")
    if floods:                           # HydroFlash: flood_allocate keeps whatever is allocated -> start from a fresh module
        for k in _FLOOD_OUT:
            _free_alloc(MODELS, k)
        for k, mod in (("flood_dw", "flood_dw"), ("flood_dsed", "flood_dsed"), ("flood_av", "flood_av"), ("flood_cmax", "flood_cmax"),
                       ("flood_umbral", "flood_umbral"), ("flood_step", "flood_step"), ("flood_hmax", "flood_hmax")):
            _set_scalar(MODELS, mod, float(inp[k]), f32)
        _set_scalar(MODELS, "flood_max_iter", int(inp.get("flood_max_iter", 10)), i32)
        for k in ("flood_w", "flood_d50", "flood_slope"):
            _set_alloc(MODELS, k, _cell2(inp[k], 1, N, np.float32), np.float32)
        _set_alloc(MODELS, "flood_sections", np.asfortranarray(inp["flood_sections"], dtype=np.float32), np.float32)
        _set_alloc(MODELS, "flood_sec_cells", np.asfortranarray(inp["flood_sec_cells"], dtype=np.float32), np.float32)
    tmp = None
    blank_path = ""
    if rain_files is not None:
        pb, ph = rain_files
    else:
        if workdir is None:
            tmp = tempfile.TemporaryDirectory(prefix="wref_")
            workdir = tmp.name
        pb, ph = os.path.join(workdir, "rain.bin"), os.path.join(workdir, "rain.hdr")
        write_rain(pb, ph, inp["rain"], np.asarray(inp["pos_evento"])[:T])
    conv_files = [blank_path] * 4
    if int(inp.get("separate_rain", 0)) == 1:            # ruta_binConv, ruta_binStra, ruta_hdrConv, ruta_hdrStra
        if tmp is None and workdir is None:
            tmp = tempfile.TemporaryDirectory(prefix="wref_")
            workdir = tmp.name
        wd = workdir or tmp.name
        names = {}
        for k in ("conv", "stra"):
            names[k] = (os.path.join(wd, k + ".bin"), os.path.join(wd, k + ".hdr"))
            # one spare header row: the separate reader wants Nintervals + rain_first_point <= Ntotal (:1026)
            pk = np.concatenate([np.asarray(inp["pos_" + k])[:T], [1]])
            write_rain(names[k][0], names[k][1], inp[k], pk)
        conv_files = [names["conv"][0], names["stra"][0], names["conv"][1], names["stra"][1]]
        for k in ("storage_conv", "storage_stra"):
            _free_alloc(MODELS, k)
    dumps = dumps or {}
    rutas = {k: _ruta(dumps.get(k)) for k in ("ruta_storage", "ruta_speed", "ruta_vfluxes", "ruta_retorno", "ruta_rc")}
    blank = _ruta("")
    calib = np.ascontiguousarray(inp["calib"], np.float32)
    assert calib.size == 11
    stoin = _cell2(inp["stoin"], 5, N, np.float32) if inp.get("stoin") is not None else np.zeros((5, N), np.float32, order="F")
    hsin = _cell2(inp["hspeedin"], 4, N, np.float32) if inp.get("hspeedin") is not None else np.zeros((4, N), np.float32, order="F")
    def z(*s):
        # one spare column: StoOut(4, N_cel+1) is written at a channel outlet (modelosv2.f90:468, :617-618)
        return np.zeros(s[:-1] + (s[-1] + 1,), np.float32, order="F")[..., :s[-1]]

    res = dict(q=z(n_cont, T), qsed=z(n_cont, 3, T), qseparated=z(n_cont, 3, T), hum=z(n_conth, T), st1=z(n_conth, T),
               st3=z(n_conth, T), balance=z(T), speed=z(n_cont, T), areacontrol=z(n_cont, T), stoout=z(5, N),
               qsep_byrain=z(n_cont, 2, T))
    L500 = 500
    _call(MODELS, "shia_v1", _ruta(pb), _ruta(ph), calib, stoin, hsin, i32(N), i32(n_cont), i32(n_conth), i32(T),
          res["q"], res["qsed"], res["qseparated"], res["hum"], res["st1"], res["st3"], res["balance"], res["speed"],
          res["areacontrol"], res["stoout"], rutas["ruta_storage"], rutas["ruta_speed"], rutas["ruta_vfluxes"],
          _ruta(conv_files[0]), _ruta(conv_files[1]), _ruta(conv_files[2]), _ruta(conv_files[3]),
          res["qsep_byrain"], rutas["ruta_retorno"], rutas["ruta_rc"],
          *([L500] * 11))
    # module-level results (read back by wmf.py:4548-4601)
    for k, rank in (("mean_rain", 2), ("acum_rain", 2), ("mean_storage", 2), ("mean_speed", 2), ("mean_retorno", 1),
                    ("retorned", 2), ("mean_vfluxes", 2), ("vfluxes", 2), ("rc_coef", 2)):
        v = _get_alloc(MODELS, k, rank, np.float32)
        if v is not None:
            res[k] = v.reshape(-1) if (rank == 2 and v.shape[0] == 1) else v
    if floods:
        res["flood_flood"] = _get_alloc(MODELS, "flood_flood", 2, np.int32).reshape(-1)
        for k in ("flood_h", "flood_ufr", "flood_cr", "flood_q", "flood_qsed", "flood_speed"):
            res[k] = _get_alloc(MODELS, k, 2, np.float32).reshape(-1)
        res["flood_profundidad"] = _get_alloc(MODELS, "flood_profundidad", 2, np.float32)
        res["flood_scalars"] = np.array([_get_scalar(MODELS, "flood_rdf", f32), _get_scalar(MODELS, "flood_area", f32),
                                         _get_scalar(MODELS, "flood_diff", f32)], np.float32)
    if int(inp.get("separate_rain", 0)) == 1:
        res["storage_conv"] = _get_alloc(MODELS, "storage_conv", 2, np.float32)
        res["storage_stra"] = _get_alloc(MODELS, "storage_stra", 2, np.float32)
    if sed:
        for k in ("vs", "vd", "vsc", "vdc"):
            res[k] = _get_alloc(MODELS, k, 2, np.float32)
        res["volero"] = _get_alloc(MODELS, "volero", 1, np.float32)
        res["voldepo"] = _get_alloc(MODELS, "voldepo", 1, np.float32)
        res["erot"] = np.array((C.c_float * 3).from_address(_modvar(MODELS, "erot")), np.float32)
        res["dept"] = np.array((C.c_float * 3).from_address(_modvar(MODELS, "dept")), np.float32)
    if sl:
        for k in ("sl_zmin", "sl_zcrit", "sl_zmax", "sl_bo", "sl_riskvector", "sl_slideocurrence"):
            res[k] = _get_alloc(MODELS, k, 2, np.float32).reshape(-1)
        res["sl_slideacumulate"] = _get_alloc(MODELS, "sl_slideacumulate", 2, np.int32).reshape(-1)
        res["sl_slidencelltime"] = _get_alloc(MODELS, "sl_slidencelltime", 1, np.int32)
    if tmp is not None:
        tmp.cleanup()
    return res


def calc_speed(sm, coef, expo, elem_long, speed, dt, calc_niter=5):
    """``area = calc_speed(sm, coef, expo, elem_long, speed)`` (modelosv2.f90:1278-1293); returns (area, speed)."""
    _set_scalar(MODELS, "dt", float(dt), f32)
    _set_scalar(MODELS, "calc_niter", int(calc_niter), i32)
    s, area = f32(speed), f32()
    _call(MODELS, "calc_speed", f32(sm), f32(coef), f32(expo), f32(elem_long), s, area)
    return area.value, s.value


def _read_records(path, n):
    raw = np.fromfile(path, np.int32)
    return raw.reshape(-1, n)


def rain_idw(xy_basin, coord, rain, pp, umbral, workdir=None, mask=None, nhills=0):
    """The reference's ``rain_idw`` (modelosv2.f90:1195-1271), cells model: returns what it writes to ``ruta`` as a record
    table plus meanRain / posIds."""
    xy, co, r = _f(xy_basin), _f(coord), _f(rain)
    n, nc, T = xy.shape[1], co.shape[1], r.shape[1]
    tmp = tempfile.TemporaryDirectory(prefix="wref_") if workdir is None else None
    path = os.path.join(workdir or tmp.name, "idw.bin")
    mean, pos = np.zeros(T, np.float32), np.zeros(T, np.int32)
    mask = np.ones(n, np.int32) if not nhills else np.ascontiguousarray(mask, np.int32)
    _call(MODELS, "rain_idw", xy, co, r, f32(pp), i32(n), i32(nc), i32(T), i32(int(nhills)), _ruta(path, 255), f32(umbral), mean, pos,
          mask, 255)
    out = dict(table=_read_records(path, nhills if nhills else n), mean_rain=mean, pos_ids=pos)
    if tmp is not None:
        tmp.cleanup()
    return out


def rain_mit(xy_basin, coord, rain, tin, tin_perte, umbral, workdir=None, mask=None, nhills=0):
    """The reference's ``rain_mit`` (modelosv2.f90:1104-1194), cells model."""
    xy, co, r = _f(xy_basin), _f(coord), _f(rain)
    tn = _i(tin)
    tp = np.asfortranarray(np.asarray(tin_perte, np.int32).reshape(1, -1))
    n, nc, T = xy.shape[1], co.shape[1], r.shape[1]
    tmp = tempfile.TemporaryDirectory(prefix="wref_") if workdir is None else None
    path = os.path.join(workdir or tmp.name, "mit.bin")
    mean, pos = np.zeros(T, np.float32), np.zeros(T, np.int32)
    mask = np.ones(n, np.int32) if not nhills else np.ascontiguousarray(mask, np.int32)
    _call(MODELS, "rain_mit", xy, co, r, tn, tp, i32(n), i32(int(nhills)), i32(nc), i32(tn.shape[1]), i32(T), _ruta(path, 255),
          f32(umbral), mean, pos, mask, 255)
    out = dict(table=_read_records(path, nhills if nhills else n), mean_rain=mean, pos_ids=pos)
    if tmp is not None:
        tmp.cleanup()
    return out


def rain_pre_mit(xy_basin, tin, coord):
    """``tin_perte = rain_pre_mit(xy_basin, tin, coord)`` (modelosv2.f90:1068-1101): the triangle holding every cell."""
    xy, co, tn = _f(xy_basin), _f(coord), _i(tin)
    n = xy.shape[1]
    out = np.zeros((1, n), np.int32, order="F")
    _call(MODELS, "rain_pre_mit", out, xy, tn, co, i32(n), i32(tn.shape[1]), i32(co.shape[1]))
    return out.reshape(-1)
