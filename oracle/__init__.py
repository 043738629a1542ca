"""CPU ORACLE -- test infrastructure, NOT product code (parity pinned against the reference's own compiled Fortran, see wmf_oracle.h).

ctypes front-end of the C restatement of the reference's Fortran
(wmf/cuencas.f90 module ``cu``; wmf/modelosv2.f90 module ``models``).  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package.

Conventions follow f2py: rasters are ``(ncols, nrows)`` arrays, basin tables
``(3, N)``, per-cell parameter maps ``(K, N)``; everything is converted to
Fortran order fp32 / int32 before the call.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS: dict = {}

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int32)
f64p = C.POINTER(C.c_double)


def build(force: bool = False) -> None:
    """Compile the oracle (parity and timing builds) with the committed Makefile."""
    need = force or not all(
        os.path.exists(os.path.join(_HERE, n)) for n in ("libwmf_oracle.so", "libwmf_oracle_fast.so")
    )
    if not need:
        srcs = [os.path.join(_HERE, n) for n in ("wmf_oracle_cu.c", "wmf_oracle_models.c", "wmf_oracle.h")]
        newest = max(os.path.getmtime(s) for s in srcs)
        need = any(
            os.path.getmtime(os.path.join(_HERE, n)) < newest for n in ("libwmf_oracle.so", "libwmf_oracle_fast.so")
        )
    if need:
        subprocess.run(["make", "-C", _HERE, "-B", "all"], check=True, capture_output=True)


class Geo(C.Structure):
    _fields_ = [
        ("ncols", C.c_int32), ("nrows", C.c_int32),
        ("xll", C.c_float), ("yll", C.c_float), ("dx", C.c_float), ("dy", C.c_float),
        ("dxp", C.c_float), ("nodata", C.c_float),
    ]


class BasinScalars(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("area", "pend_media", "elevacion", "centrox", "centroy")]


class ShiaIn(C.Structure):
    _fields_ = [
        ("n_cel", C.c_int32), ("n_reg", C.c_int32), ("n_cont", C.c_int32), ("n_conth", C.c_int32),
        ("dt", C.c_float), ("dxp", C.c_float), ("retorno_gr", C.c_float), ("retorno_aq", C.c_float),
        ("storage_constant", C.c_float),
        ("speed_type", C.c_int32 * 3), ("calc_niter", C.c_int32),
        ("sim_sediments", C.c_int32), ("sim_slides", C.c_int32),
        ("show_storage", C.c_int32), ("show_speed", C.c_int32), ("show_mean_speed", C.c_int32),
        ("show_mean_retorno", C.c_int32), ("save_retorno", C.c_int32),
        ("save_vfluxes", C.c_int32), ("save_rc", C.c_int32),
        ("separate_rain", C.c_int32), ("separate_fluxes", C.c_int32),
        ("calib", C.c_float * 11),
        ("drena", i32p), ("unit_type", i32p),
        ("hill_long", f32p), ("hill_slope", f32p), ("stream_long", f32p), ("stream_slope", f32p),
        ("elem_area", f32p),
        ("v_coef", f32p), ("h_coef", f32p), ("h_exp", f32p),
        ("max_capilar", f32p), ("max_gravita", f32p), ("max_aquifer", f32p),
        ("evpserie", f32p), ("storage", f32p),
        ("control", i32p), ("control_h", i32p),
        ("stoin", f32p), ("hspeedin", f32p),
        ("rain", i32p), ("n_records", C.c_int32), ("pos_evento", i32p),
        ("sed_factor", C.c_float), ("wi", C.c_float * 3), ("diametro", C.c_float * 3),
        ("krus", f32p), ("crus", f32p), ("prus", f32p), ("parliac", f32p), ("vd_init", f32p),
        ("sl_gullienogullie", C.c_int32), ("sl_fs", C.c_float), ("sl_gammaw", C.c_float),
        ("sl_zs", f32p), ("sl_gammas", f32p), ("sl_cohesion", f32p), ("sl_frictionangle", f32p),
        ("sl_radslope", f32p),
        ("conv", i32p), ("stra", i32p), ("n_conv_records", C.c_int32), ("n_stra_records", C.c_int32),
        ("pos_conv", i32p), ("pos_stra", i32p),
        ("sim_floods", C.c_int32), ("flood_sec_tam", C.c_int32),
        ("flood_dw", C.c_float), ("flood_dsed", C.c_float), ("flood_av", C.c_float), ("flood_cmax", C.c_float),
        ("flood_umbral", C.c_float), ("flood_step", C.c_float), ("flood_hmax", C.c_float),
        ("flood_w", f32p), ("flood_d50", f32p), ("flood_slope", f32p), ("flood_sections", f32p), ("flood_sec_cells", f32p),
    ]


class ShiaOut(C.Structure):
    _fields_ = [
        ("q", f32p), ("qsed", f32p), ("qseparated", f32p), ("hum", f32p), ("st1", f32p), ("st3", f32p),
        ("balance", f32p), ("balance64", f64p), ("speed", f32p), ("areacontrol", f32p),
        ("stoout", f32p), ("hspeed_out", f32p), ("mean_rain", f32p), ("acum_rain", f32p),
        ("mean_storage", f32p), ("mean_speed", f32p), ("mean_retorno", f32p), ("retorned", f32p),
        ("mean_vfluxes", f32p), ("rc_coef", f32p),
        ("mean_storage64", f64p), ("mean_speed64", f64p), ("mean_retorno64", f64p),
        ("mean_vfluxes64", f64p), ("vfluxes_last", f32p), ("retorned_last", f32p),
        ("vs", f32p), ("vd", f32p), ("vsc", f32p), ("vdc", f32p), ("volero", f32p), ("voldepo", f32p),
        ("erot", f32p), ("dept", f32p), ("erot64", f64p), ("dept64", f64p),
        ("sl_zmin", f32p), ("sl_zcrit", f32p), ("sl_zmax", f32p), ("sl_bo", f32p),
        ("sl_riskvector", f32p), ("sl_slideocurrence", f32p),
        ("sl_slideacumulate", i32p), ("sl_slidencelltime", i32p),
        ("qsep_byrain", f32p), ("storage_conv", f32p), ("storage_stra", f32p),
        ("flood_flood", i32p), ("flood_h", f32p), ("flood_ufr", f32p), ("flood_cr", f32p), ("flood_q", f32p),
        ("flood_qsed", f32p), ("flood_speed", f32p), ("flood_profundidad", f32p), ("flood_scalars", f32p),
    ]


def lib(fast: bool = False) -> C.CDLL:
    key = "fast" if fast else "parity"
    if key not in _LIBS:
        build()
        L = C.CDLL(os.path.join(_HERE, "libwmf_oracle_fast.so" if fast else "libwmf_oracle.so"))
        L.wo_basin_find.restype = C.c_int32
        L.wo_shia_v1.restype = C.c_int
        L.wo_calc_speed.restype = C.c_float
        L.wo_calc_speed.argtypes = [C.c_float] * 4 + [f32p, C.c_float, C.c_int32]
        _LIBS[key] = L
    return _LIBS[key]


def _f(a, shape=None):
    a = np.asfortranarray(a, dtype=np.float32)
    return a


def _i(a):
    return np.asfortranarray(a, dtype=np.int32)


def _p(a, typ):
    return a.ctypes.data_as(typ)


class Cu:
    """Oracle twin of the f2py module object ``cu`` (cuencas.f90)."""

    def __init__(self, fast: bool = False):
        self._L = lib(fast)
        self.ncols = 0
        self.nrows = 0
        self.xll = 0.0
        self.yll = 0.0
        self.dx = 1.0
        self.dy = 1.0
        self.dxp = 1.0
        self.nodata = -9999.0
        self._basin_temp = None
        self.area = self.pend_media = self.elevacion = self.centrox = self.centroy = 0.0

    def _geo(self) -> Geo:
        return Geo(int(self.ncols), int(self.nrows), float(self.xll), float(self.yll), float(self.dx),
                   float(self.dy), float(self.dxp), float(self.nodata))

    # cuencas.f90:77-90
    def coord2fil_col(self, x, y):
        col, fil = C.c_int32(), C.c_int32()
        g = self._geo()
        self._L.wo_coord2fil_col(C.byref(g), C.c_float(x), C.c_float(y), C.byref(col), C.byref(fil))
        return col.value, fil.value

    # cuencas.f90:525-591
    def basin_find(self, x, y, dir_map):
        d = _i(dir_map)
        nc, nr = d.shape
        assert (nc, nr) == (self.ncols, self.nrows)
        self._basin_temp = np.empty(3 * nc * nr, dtype=np.int32)
        g = self._geo()
        n = self._L.wo_basin_find(C.byref(g), C.c_float(x), C.c_float(y), _p(d, i32p), nc, nr,
                                  _p(self._basin_temp, i32p))
        return int(n)

    # cuencas.f90:592-607
    def basin_cut(self, nceldas):
        out = np.empty((3, nceldas), dtype=np.int32, order="F")
        g = self._geo()
        self._L.wo_basin_cut(C.byref(g), _p(self._basin_temp, i32p), int(nceldas), _p(out, i32p))
        return out

    def basin_acum(self, basin_f):
        b = _i(basin_f)
        n = b.shape[1]
        acum = np.empty(n, dtype=np.int32)
        area = C.c_float()
        g = self._geo()
        self._L.wo_basin_acum(C.byref(g), _p(b, i32p), n, _p(acum, i32p), C.byref(area))
        self.area = area.value
        return acum

    def basin_acum_var(self, drain, var):
        d = np.ascontiguousarray(np.asarray(drain).reshape(-1), dtype=np.int32)
        v = np.ascontiguousarray(np.asarray(var).reshape(-1), dtype=np.float32)
        out = np.empty((1, d.size), dtype=np.float32, order="F")
        self._L.wo_basin_acum_var(_p(d, i32p), d.size, _p(v, f32p), _p(out, f32p))
        return out

    def basin_basics(self, basin_f, dem, dir_vec):
        b = _i(basin_f)
        n = b.shape[1]
        dem = np.ascontiguousarray(dem, dtype=np.float32)
        dv = np.ascontiguousarray(dir_vec, dtype=np.int32)
        acum = np.empty(n, np.int32)
        lng, pend, elev = (np.empty(n, np.float32) for _ in range(3))
        sc = BasinScalars()
        g = self._geo()
        self._L.wo_basin_basics(C.byref(g), _p(b, i32p), _p(dem, f32p), _p(dv, i32p), n, _p(acum, i32p),
                                _p(lng, f32p), _p(pend, f32p), _p(elev, f32p), C.byref(sc))
        self.area, self.pend_media, self.elevacion = sc.area, sc.pend_media, sc.elevacion
        self.centrox, self.centroy = sc.centrox, sc.centroy
        return acum, lng, pend, elev

    def basin_coordxy(self, basin_f):
        b = _i(basin_f)
        n = b.shape[1]
        x, y = np.empty(n, np.float32), np.empty(n, np.float32)
        g = self._geo()
        self._L.wo_basin_coordxy(C.byref(g), _p(b, i32p), n, _p(x, f32p), _p(y, f32p))
        return x, y

    def basin_map2basin(self, basin_f, mapa, xllm, yllm, dxm, dym, nodatam):
        b = _i(basin_f)
        n = b.shape[1]
        m = _f(mapa)
        ncm, nrm = m.shape
        vec = np.empty(n, np.float32)
        g = self._geo()
        self._L.wo_basin_map2basin(C.byref(g), _p(b, i32p), n, _p(m, f32p), C.c_float(xllm), C.c_float(yllm),
                                   ncm, nrm, C.c_float(dxm), C.c_float(dym), C.c_float(nodatam), _p(vec, f32p))
        return vec

    def basin_2map_find(self, basin):
        b = _i(basin)
        a, c = C.c_int32(), C.c_int32()
        self._L.wo_basin_2map_find(_p(b, i32p), b.shape[1], C.byref(a), C.byref(c))
        return a.value, c.value

    def basin_2map(self, basin, var, map_ncols, map_nrows):
        b = _i(basin)
        v = np.ascontiguousarray(var, dtype=np.float32)
        mapa = np.empty((map_ncols, map_nrows), np.float32, order="F")
        xll, yll = C.c_float(), C.c_float()
        g = self._geo()
        self._L.wo_basin_2map(C.byref(g), _p(b, i32p), _p(v, f32p), b.shape[1], map_ncols, map_nrows,
                              _p(mapa, f32p), C.byref(xll), C.byref(yll))
        return mapa, xll.value, yll.value

    def basin_float_var2map(self, basin_f, vect, nc, nf):
        b = _i(basin_f)
        v = np.ascontiguousarray(vect, dtype=np.float32)
        mapa = np.empty((nc, nf), np.float32, order="F")
        g = self._geo()
        self._L.wo_basin_float_var2map(C.byref(g), _p(b, i32p), _p(v, f32p), b.shape[1], nc, nf, _p(mapa, f32p))
        return mapa

    def basin_stream_nod(self, basin_f, acum, umbral):
        b = _i(basin_f)
        n = b.shape[1]
        a = np.ascontiguousarray(acum, dtype=np.int32)
        cauce, nodos, traz = (np.empty(n, np.int32) for _ in range(3))
        nn, ncau = C.c_int32(), C.c_int32()
        self._L.wo_basin_stream_nod(_p(b, i32p), _p(a, i32p), n, int(umbral), _p(cauce, i32p), _p(nodos, i32p),
                                    _p(traz, i32p), C.byref(nn), C.byref(ncau))
        return cauce, nodos, traz, nn.value, ncau.value

    def basin_stream_type(self, basin_f, acum, umbrales):
        a = np.ascontiguousarray(acum, dtype=np.int32)
        u = np.ascontiguousarray(umbrales, dtype=np.int32)
        out = np.empty(a.size, np.int32)
        self._L.wo_basin_stream_type(_p(a, i32p), a.size, _p(u, i32p), u.size, _p(out, i32p))
        return out

    def geo_acum_to_cauce(self, acum, umbral):
        a = _i(acum)
        out = np.empty(a.shape, np.int32, order="F")
        self._L.wo_geo_acum_to_cauce(_p(a, i32p), a.size, int(umbral), _p(out, i32p))
        return out

    def geo_hand(self, basin_f, elev, lng, cauce):
        b = _i(basin_f)
        n = b.shape[1]
        e = np.ascontiguousarray(elev, np.float32)
        l = np.ascontiguousarray(lng, np.float32)
        c = np.ascontiguousarray(cauce, np.int32)
        hand, hdnd, aq = np.empty(n, np.float32), np.empty(n, np.float32), np.empty(n, np.int32)
        self._L.wo_geo_hand(_p(b, i32p), _p(e, f32p), _p(l, f32p), _p(c, i32p), n, _p(hand, f32p),
                            _p(hdnd, f32p), _p(aq, i32p))
        return hand, hdnd, aq

    def geo_hand_global(self, dem, dir_map, red):
        d, di, r = _f(dem), _i(dir_map), _i(red)
        nc, nr = d.shape
        hand = np.empty((nc, nr), np.float32, order="F")
        g = self._geo()
        self._L.wo_geo_hand_global(C.byref(g), _p(d, f32p), _p(di, i32p), _p(r, i32p), nc, nr, _p(hand, f32p))
        return hand

    def basin_time_to_out(self, basin_f, lng, speed):
        b = _i(basin_f)
        n = b.shape[1]
        l = np.ascontiguousarray(lng, np.float32)
        s = np.ascontiguousarray(speed, np.float32)
        t = np.empty(n, np.float32)
        self._L.wo_basin_time_to_out(_p(b, i32p), _p(l, f32p), _p(s, f32p), n, _p(t, f32p))
        return t

    def basin_propagate(self, basin_f, var):
        b = _i(basin_f)
        n = b.shape[1]
        v = np.ascontiguousarray(var, np.float32)
        out = np.empty(n, np.float32)
        self._L.wo_basin_propagate(_p(b, i32p), _p(v, f32p), n, _p(out, f32p))
        return out


    def basin_point2var(self, basin_f, id_coord, xy_coord):
        """res_coord,basin_pts = basin_point2var(basin_f,id_coord,xy_coord,[ncoord,nceldas])   -- cuencas.f90:1230-1261"""
        b = _i(basin_f)
        n = b.shape[1]
        ids, xy = np.ascontiguousarray(id_coord, np.int32), _f(xy_coord)
        res, pts = np.zeros(ids.size, np.int32), np.zeros(n, np.int32)
        g = self._geo()
        self._L.wo_basin_point2var(C.byref(g), _p(b, i32p), _p(ids, i32p), _p(xy, f32p), ids.size, n, _p(res, i32p), _p(pts, i32p))
        return res, pts

    def basin_stream_point2stream(self, basin_f, cauce, id_coord, xy_coord):
        """res_coord,basin_pts,xy_new = basin_stream_point2stream(basin_f,cauce,id_coord,xy_coord,[ncoord,nceldas])
        -- cuencas.f90:1524-1577"""
        b = _i(basin_f)
        n = b.shape[1]
        ca = np.ascontiguousarray(cauce, np.int32)
        ids, xy = np.ascontiguousarray(id_coord, np.int32), _f(xy_coord)
        res, pts = np.zeros(ids.size, np.int32), np.zeros(n, np.int32)
        xyn = np.zeros((2, ids.size), np.float32, order="F")
        g = self._geo()
        self._L.wo_basin_stream_point2stream(C.byref(g), _p(b, i32p), _p(ca, i32p), _p(ids, i32p), _p(xy, f32p), ids.size, n,
                                             _p(res, i32p), _p(pts, i32p), _p(xyn, f32p))
        return res, pts, xyn

    def basin_stream_sections(self, basin_f, cauces, directions, dem, num_celdas):
        """secciones,secciones_cel = basin_stream_sections(basin_f,cauces,directions,dem,num_celdas,[nceldas,ncols,nrows])
        -- cuencas.f90:1464-1523"""
        b = _i(basin_f)
        n = b.shape[1]
        ca, di = np.ascontiguousarray(cauces, np.int32), np.ascontiguousarray(directions, np.int32)
        d = _f(dem)
        S = 2 * int(num_celdas) + 1
        sec, cel = np.zeros((S, n), np.float32, order="F"), np.zeros((S, n), np.int32, order="F")
        self._L.wo_basin_stream_sections(_p(b, i32p), _p(ca, i32p), _p(di, i32p), _p(d, f32p), int(num_celdas), n, d.shape[0],
                                         d.shape[1], _p(sec, f32p), _p(cel, i32p))
        return sec, cel

    def stream_find_to_corr(self, x, y, dem, dir_map, cauce):
        """lat,lon = stream_find_to_corr(x,y,dem,dir,cauce,[nc,nf])   -- cuencas.f90:372-417"""
        d, ca = _i(dir_map), _i(cauce)
        lat, lon = C.c_float(), C.c_float()
        g = self._geo()
        self._L.wo_stream_find_to_corr(C.byref(g), C.c_float(x), C.c_float(y), _p(d, i32p), _p(ca, i32p), d.shape[0], d.shape[1],
                                       C.byref(lat), C.byref(lon))
        return lat.value, lon.value

    # ---- sub-basin (hills) topology, cuencas.f90:1380-1438, 1835-2118 ----
    def basin_stream_slope(self, basin_f, elev, lng, nodos, n_cauce):
        b = _i(basin_f)
        n = b.shape[1]
        e, l = np.ascontiguousarray(elev, np.float32), np.ascontiguousarray(lng, np.float32)
        nd = np.ascontiguousarray(nodos, np.int32)
        s, ln = np.empty(n, np.float32), np.empty(n, np.float32)
        g = self._geo()
        self._L.wo_basin_stream_slope(C.byref(g), _p(b, i32p), _p(e, f32p), _p(l, f32p), _p(nd, i32p), int(n_cauce), n,
                                      _p(s, f32p), _p(ln, f32p))
        return s, ln

    def basin_subbasin_nod(self, basin_f, acum, umbral):
        b = _i(basin_f)
        n = b.shape[1]
        a = np.ascontiguousarray(acum, np.int32)
        cauce, nodos = np.empty(n, np.int32), np.empty(n, np.int32)
        nn = C.c_int32()
        self._sub_basins = np.zeros(2 * n, np.int32)
        self._L.wo_basin_subbasin_nod(_p(b, i32p), _p(a, i32p), n, int(umbral), _p(cauce, i32p), _p(nodos, i32p), C.byref(nn),
                                      _p(self._sub_basins, i32p))
        return cauce, nodos, nn.value

    def basin_subbasin_cut(self, n_nodos):
        return np.array(self._sub_basins[:2 * n_nodos].reshape(2, n_nodos, order="F"), order="F")

    def basin_subbasin_find(self, basin_f, nodos, n_nodos):
        b = _i(basin_f)
        n = b.shape[1]
        nd = np.ascontiguousarray(nodos, np.int32)
        sub_pert = np.empty(n, np.int32)
        self._L.wo_basin_subbasin_find(_p(b, i32p), _p(nd, i32p), n, _p(sub_pert, i32p))
        return sub_pert, np.zeros(n_nodos, np.int32)

    def basin_subbasin_horton(self, sub_basins, nceldas):
        sb = _i(sub_basins)
        nn = sb.shape[1]
        sh, nh = np.empty(nn, np.int32), np.empty(nceldas, np.int32)
        self._L.wo_basin_subbasin_horton(_p(sb, i32p), nn, int(nceldas), _p(sh, i32p), _p(nh, i32p))
        return sh, nh

    def basin_subbasin_long(self, sub_pert, cauce, lng, sub_basin, sub_horton):
        sp, ca = np.ascontiguousarray(sub_pert, np.int32), np.ascontiguousarray(cauce, np.int32)
        li = np.ascontiguousarray(np.asarray(lng), np.float32).astype(np.int32)        # f2py casts to the INTEGER dummy
        sb, sh = _i(sub_basin), np.ascontiguousarray(sub_horton, np.int32)
        nn = sb.shape[1]
        out = np.empty(nn, np.float32)
        mx, nmx = C.c_float(), C.c_int32()
        self._L.wo_basin_subbasin_long(_p(sp, i32p), _p(ca, i32p), _p(li, i32p), _p(sb, i32p), _p(sh, i32p), nn, sp.size,
                                       _p(out, f32p), C.byref(mx), C.byref(nmx))
        return out, mx.value, nmx.value

    def basin_subbasin_stream_prop(self, sub_pert, cauce, lng, slope, n_nodos):
        sp, ca = np.ascontiguousarray(sub_pert, np.int32), np.ascontiguousarray(cauce, np.int32)
        l, s = np.ascontiguousarray(lng, np.float32), np.ascontiguousarray(slope, np.float32)
        ss, sl = np.empty(n_nodos, np.float32), np.empty(n_nodos, np.float32)
        self._L.wo_basin_subbasin_stream_prop(_p(sp, i32p), _p(ca, i32p), _p(l, f32p), _p(s, f32p), sp.size, int(n_nodos),
                                              _p(ss, f32p), _p(sl, f32p))
        return ss, sl

    def basin_subbasin_map2subbasin(self, sub_pert, var, n_nodos, cauce, sum_mean):
        sp, ca = np.ascontiguousarray(sub_pert, np.int32), np.ascontiguousarray(cauce, np.int32)
        v = np.ascontiguousarray(var, np.float32)
        out = np.empty(n_nodos, np.float32)
        self._L.wo_basin_subbasin_map2subbasin(_p(sp, i32p), _p(v, f32p), int(n_nodos), sp.size, _p(ca, i32p), int(sum_mean),
                                               _p(out, f32p))
        return out


# keys of the SHIA input dict: name -> (ctype pointer, K)  (array of shape (K,N), or (N,) / (1,N) for K=1)
_CELL_F = {
    "hill_long": 1, "hill_slope": 1, "stream_long": 1, "stream_slope": 1, "elem_area": 1,
    "v_coef": 4, "h_coef": 4, "h_exp": 4, "max_capilar": 1, "max_gravita": 1, "max_aquifer": 1,
    "storage": 5,
}
_SED_F = {"krus": 1, "crus": 1, "prus": 1, "parliac": 3}
_SL_F = {"sl_zs": 1, "sl_gammas": 1, "sl_cohesion": 1, "sl_frictionangle": 1, "sl_radslope": 1}


def _cellarr(a, K, N, dtype):
    a = np.asarray(a)
    if a.ndim == 1:
        assert K == 1 and a.shape[0] == N, (a.shape, K, N)
        return np.ascontiguousarray(a, dtype=dtype)
    assert a.shape == (K, N), (a.shape, K, N)
    return np.ascontiguousarray(np.asfortranarray(a, dtype=dtype).ravel(order="F"))


def shia_v1(inp: dict, fast: bool = False) -> dict:
    """Run the oracle's shia_v1 (modelosv2.f90:191-869) on an input dict.

    ``inp`` uses the lower-case ``models`` attribute names (``drena``, ``unit_type``, ``v_coef`` ...)
    plus ``calib`` (11), ``n_reg``, ``rain`` (n_records, N) int32, ``pos_evento`` (n_reg) 1-based,
    and the scalar switches.  Returns a dict of fp32 arrays in f2py shapes.
    """
    L = lib(fast)
    keep = []
    drena = np.asarray(inp["drena"])
    drena = np.ascontiguousarray(drena[0] if drena.ndim == 2 else drena, dtype=np.int32)
    N = drena.size
    T = int(inp["n_reg"])
    control = _cellarr(inp.get("control", np.zeros(N)), 1, N, np.int32)
    control_h = _cellarr(inp.get("control_h", np.zeros(N)), 1, N, np.int32)
    n_cont = int(np.count_nonzero(control)) + 1
    n_conth = max(1, int(np.count_nonzero(control_h)))
    si = ShiaIn()
    si.n_cel, si.n_reg, si.n_cont, si.n_conth = N, T, n_cont, n_conth
    si.dt = float(inp["dt"])
    si.dxp = float(inp.get("dxp", 1.0))
    si.retorno_gr = float(inp.get("retorno_gr", 0))
    si.retorno_aq = float(inp.get("retorno_aq", 0))
    si.storage_constant = float(inp.get("storage_constant", 0.0))
    si.speed_type = (C.c_int32 * 3)(*[int(v) for v in inp.get("speed_type", [1, 1, 1])])
    si.calc_niter = int(inp.get("calc_niter", 5))
    for k in ("sim_sediments", "sim_slides", "show_storage", "show_speed", "show_mean_speed",
              "show_mean_retorno", "save_retorno", "save_vfluxes", "save_rc", "separate_fluxes", "separate_rain"):
        setattr(si, k, int(inp.get(k, 0)))
    si.calib = (C.c_float * 11)(*[float(v) for v in inp["calib"]])

    def put(name, arr, typ):
        keep.append(arr)
        setattr(si, name, _p(arr, typ))

    put("drena", drena, i32p)
    put("unit_type", _cellarr(inp["unit_type"], 1, N, np.int32), i32p)
    for k, K in _CELL_F.items():
        put(k, _cellarr(inp[k], K, N, np.float32), f32p)
    put("evpserie", np.ascontiguousarray(inp.get("evpserie", np.ones(T)), dtype=np.float32), f32p)
    put("control", control, i32p)
    put("control_h", control_h, i32p)
    if inp.get("stoin") is not None:
        put("stoin", _cellarr(inp["stoin"], 5, N, np.float32), f32p)
    if inp.get("hspeedin") is not None:
        put("hspeedin", _cellarr(inp["hspeedin"], 4, N, np.float32), f32p)
    rain = np.ascontiguousarray(inp["rain"], dtype=np.int32)
    assert rain.ndim == 2 and rain.shape[1] == N
    put("rain", rain, i32p)
    si.n_records = rain.shape[0]
    pos = np.ascontiguousarray(inp["pos_evento"], dtype=np.int32)
    assert pos.size >= T and pos.min() >= 1 and pos.max() <= rain.shape[0]
    put("pos_evento", pos, i32p)
    if si.separate_rain == 1:            # rain-type tables (n_records, N) int32 and their record of every step (1-based)
        for k in ("conv", "stra"):
            tab = np.ascontiguousarray(inp[k], dtype=np.int32)
            assert tab.ndim == 2 and tab.shape[1] == N
            put(k, tab, i32p)
            setattr(si, f"n_{k}_records", tab.shape[0])
            pk = np.ascontiguousarray(inp["pos_" + k], dtype=np.int32)
            assert pk.size >= T
            put("pos_" + k, pk, i32p)
    floods = int(inp.get("sim_floods", 0)) == 1
    if floods:                           # HydroFlash (modelosv2.f90:728-743, 1711-1822)
        si.sim_floods = 1
        for k in ("flood_dw", "flood_dsed", "flood_av", "flood_cmax", "flood_umbral", "flood_step", "flood_hmax"):
            setattr(si, k, float(inp[k]))
        for k in ("flood_w", "flood_d50", "flood_slope"):
            put(k, _cellarr(inp[k], 1, N, np.float32), f32p)
        sec = np.asfortranarray(inp["flood_sections"], dtype=np.float32)
        si.flood_sec_tam = sec.shape[0]
        put("flood_sections", np.ascontiguousarray(sec.ravel(order="F")), f32p)
        put("flood_sec_cells", np.ascontiguousarray(np.asfortranarray(inp["flood_sec_cells"], dtype=np.float32).ravel(order="F")), f32p)
    sed = si.sim_sediments == 1
    sl = si.sim_slides == 1
    if sed:
        si.sed_factor = float(inp.get("sed_factor", 1.0))
        si.wi = (C.c_float * 3)(*[float(v) for v in inp["wi"]])
        si.diametro = (C.c_float * 3)(*[float(v) for v in inp["diametro"]])
        for k, K in _SED_F.items():
            put(k, _cellarr(inp[k], K, N, np.float32), f32p)
        if inp.get("vd") is not None:
            put("vd_init", _cellarr(inp["vd"], 3, N, np.float32), f32p)
    if sl:
        si.sl_gullienogullie = int(inp.get("sl_gullienogullie", 0))
        si.sl_fs = float(inp.get("sl_fs", 1.0))
        si.sl_gammaw = float(inp.get("sl_gammaw", 9.8))
        for k, K in _SL_F.items():
            put(k, _cellarr(inp[k], K, N, np.float32), f32p)

    so = ShiaOut()
    res = {}

    def out(name, shape, dtype=np.float32, typ=f32p, zero=True):
        a = (np.zeros if zero else np.empty)(shape, dtype=dtype, order="F")
        res[name] = a
        setattr(so, name, _p(a, typ))

    out("q", (n_cont, T))
    out("qsed", (n_cont, 3, T))
    out("qseparated", (n_cont, 3, T))
    out("hum", (n_conth, T))
    out("st1", (n_conth, T))
    out("st3", (n_conth, T))
    out("balance", (T,))
    out("balance64", (T,), np.float64, f64p)
    out("speed", (n_cont, T))
    out("areacontrol", (n_cont, T))
    out("stoout", (5, N))
    out("hspeed_out", (4, N))
    out("mean_rain", (T,))
    out("acum_rain", (N,))
    out("mean_storage", (5, T))
    out("mean_speed", (4, T))
    out("mean_retorno", (T,))
    out("retorned", (N,))
    out("mean_vfluxes", (4, T))
    out("rc_coef", (2, N))
    out("mean_storage64", (5, T), np.float64, f64p)
    out("mean_speed64", (4, T), np.float64, f64p)
    out("mean_retorno64", (T,), np.float64, f64p)
    out("mean_vfluxes64", (4, T), np.float64, f64p)
    out("vfluxes_last", (4, N))
    if floods:
        out("flood_flood", (N,), np.int32, i32p)
        for k in ("flood_h", "flood_ufr", "flood_cr", "flood_q", "flood_qsed", "flood_speed"):
            out(k, (N,))
        out("flood_profundidad", (si.flood_sec_tam, N))
        out("flood_scalars", (3,))
    if si.separate_rain == 1:
        out("qsep_byrain", (n_cont, 2, T))
        out("storage_conv", (5, N))
        out("storage_stra", (5, N))
    out("retorned_last", (N,))
    if sed:
        for k in ("vs", "vd", "vsc", "vdc"):
            out(k, (3, N))
        out("volero", (N,))
        out("voldepo", (N,))
        out("erot", (3,))
        out("dept", (3,))
        out("erot64", (3,), np.float64, f64p)
        out("dept64", (3,), np.float64, f64p)
    if sl:
        for k in ("sl_zmin", "sl_zcrit", "sl_zmax", "sl_bo", "sl_riskvector", "sl_slideocurrence"):
            out(k, (N,))
        out("sl_slideacumulate", (N,), np.int32, i32p)
        out("sl_slidencelltime", (T,), np.int32, i32p)
    rc = L.wo_shia_v1(C.byref(si), C.byref(so))
    if rc != 0:
        raise ValueError(f"oracle shia_v1 rejected its inputs (rc={rc})")
    return res


def calc_speed(sm, coef, expo, elem_long, speed, dt, calc_niter=5):
    s = C.c_float(speed)
    area = lib().wo_calc_speed(C.c_float(sm), C.c_float(coef), C.c_float(expo), C.c_float(elem_long),
                               C.byref(s), C.c_float(dt), int(calc_niter))
    return float(area), float(s.value)


def _rain_common(L, fn, args, n, T, mask, nhills):
    cap = T + 1
    width = nhills if nhills else n
    table = np.zeros((cap, width), np.int32)
    mean, pos, mean64 = np.zeros(T, np.float32), np.zeros(T, np.int32), np.zeros(T, np.float64)
    mk = np.ascontiguousarray(mask, np.int32) if nhills else None
    fn.restype = C.c_int32
    nrec = fn(*args, _p(table, i32p), cap, _p(mean, f32p), _p(pos, i32p), _p(mean64, f64p),
              _p(mk, i32p) if nhills else None, int(nhills))
    if nrec < 1:
        raise ValueError("oracle rain interpolation failed")
    return dict(table=table[:nrec].copy(), mean_rain=mean, pos_ids=pos, mean_rain64=mean64)


def rain_idw(xy_basin, coord, rain, pp, umbral, mask=None, nhills=0):
    """modelosv2.f90:1195-1271.  xy_basin (2,N), coord (2,ncoord), rain (ncoord,nreg); returns the record table the Fortran
    writes to ``ruta`` (row 0 = record 1 = zeros), meanRain, posIds (+ an fp64 twin of meanRain).  ``mask`` / ``nhills``:
    the hills model (records of nhills values)."""
    L = lib()
    xy, co, r = _f(xy_basin), _f(coord), _f(rain)
    n, nc, T = xy.shape[1], co.shape[1], r.shape[1]
    assert xy.shape[0] == 2 and co.shape[0] == 2 and r.shape[0] == nc
    args = (_p(xy, f32p), _p(co, f32p), _p(r, f32p), C.c_float(pp), n, nc, T, C.c_float(umbral))
    return _rain_common(L, L.wo_rain_idw, args, n, T, mask, nhills)


def rain_mit(xy_basin, coord, rain, tin, tin_perte, umbral, mask=None, nhills=0):
    """modelosv2.f90:1104-1194.  tin (3,ntin) 1-based station ids, tin_perte (N) 1-based triangle of each cell."""
    L = lib()
    xy, co, r = _f(xy_basin), _f(coord), _f(rain)
    tn = _i(tin)
    tp = np.ascontiguousarray(np.asarray(tin_perte).reshape(-1), np.int32)
    n, nc, T = xy.shape[1], co.shape[1], r.shape[1]
    args = (_p(xy, f32p), _p(co, f32p), _p(r, f32p), _p(tn, i32p), _p(tp, i32p), n, nc, tn.shape[1], T, C.c_float(umbral))
    return _rain_common(L, L.wo_rain_mit, args, n, T, mask, nhills)


def rain_pre_mit(xy_basin, tin, coord):
    """tin_perte = rain_pre_mit(xy_basin,tin,coord)   -- modelosv2.f90:1068-1101"""
    L = lib()
    xy, co, tn = _f(xy_basin), _f(coord), _i(tin)
    out = np.zeros(xy.shape[1], np.int32)
    L.wo_rain_pre_mit(_p(xy, f32p), _p(tn, i32p), _p(co, f32p), xy.shape[1], tn.shape[1], _p(out, i32p))
    return out
