"""`cu` -- host-side mirror of the reference's f2py module object ``cu`` (wmf/cuencas.f90).

Same attribute names (``cu.ncols``, ``cu.dxp`` ... set by ``wmf.read_map_raster``,
wmf.py:319-333), same positional signatures and return shapes as the f2py wrappers
(build/src.win-amd64-3.7/cumodule.c), so ``wmf.py`` call sites such as

    self.ncells    = cu.basin_find(lat, lon, DIR, cu.ncols, cu.nrows)      # wmf.py:3010
    self.structure = cu.basin_cut(self.ncells)                              # wmf.py:3012
    acum           = cu.basin_acum(self.structure, self.ncells)             # wmf.py:3018

work unchanged.  Every function is one call into libwmf_b200.so (CUDA, sm_100a).

Arrays may be numpy arrays (host: copied to the GPU and back inside the call, results are
numpy) or torch CUDA tensors laid out like the f2py buffers (device resident: used in
place, results are torch CUDA tensors).  A resident ``(3,N)`` table is a ``(N,3)`` int32
tensor; a resident raster ``(ncols,nrows)`` is a ``(nrows,ncols)`` tensor.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import is_torch_cuda, ptr

_GEO_F = ("xll", "yll", "nodata", "dx", "dy", "dxp")
_GEO_I = ("ncols", "nrows")
_RESULTS = ("area", "pend_media", "elevacion", "centrox", "centroy")


def _f2py_error(msg):
    return ValueError(msg)


class CuModule:
    """Drop-in for the Fortran module object ``cu``."""

    def __init__(self, ctx: _lib.Context | None = None):
        object.__setattr__(self, "_ctx", ctx)
        object.__setattr__(self, "_g", {"ncols": 0, "nrows": 0, "xll": 0.0, "yll": 0.0, "nodata": -9999.0,
                                        "dx": 1.0, "dy": 1.0, "dxp": 1.0})
        object.__setattr__(self, "_res", dict.fromkeys(_RESULTS, 0.0))

    # ---- module scalars --------------------------------------------------------------------
    def __getattr__(self, name):
        g = object.__getattribute__(self, "_g")
        if name in g:
            v = g[name]
            return np.int32(v) if name in _GEO_I else np.float32(v)
        res = object.__getattribute__(self, "_res")
        if name in res:
            return np.float32(res[name])
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name in _GEO_I:
            self._g[name] = int(value)
        elif name in _GEO_F:
            self._g[name] = float(np.float32(value))
        elif name in _RESULTS:
            self._res[name] = float(np.float32(value))
        else:
            object.__setattr__(self, name, value)

    # ---- plumbing --------------------------------------------------------------------------
    @property
    def ctx(self) -> _lib.Context:
        if self._ctx is None:
            object.__setattr__(self, "_ctx", _lib.default_context())
        return self._ctx

    def _push_geo(self):
        g = self._g
        c = self.ctx
        c.check(c.L.wmf_set_geo(c.h, g["ncols"], g["nrows"], g["xll"], g["yll"], g["dx"], g["dy"], g["dxp"], g["nodata"]))
        return c

    def _pull_scalars(self, c):
        buf = (C.c_float * 5)()
        c.check(c.L.wmf_get_basin_scalars(c.h, buf))
        for k, n in enumerate(_RESULTS):
            self._res[n] = float(buf[k])

    @staticmethod
    def _table(basin_f):
        """(3,N) basin table -> flat f2py buffer, N, resident?"""
        if is_torch_cuda(basin_f):
            assert basin_f.dim() == 2 and basin_f.shape[1] == 3 and basin_f.is_contiguous(), "resident table must be (N,3) int32"
            return basin_f, basin_f.shape[0], True
        b = np.asarray(basin_f)
        if b.ndim != 2 or b.shape[0] != 3:
            raise _f2py_error(f"basin_f must have shape (3,nceldas), got {b.shape}")
        return np.asfortranarray(b, dtype=np.int32), b.shape[1], False

    @staticmethod
    def _vec(v, n, dtype, name):
        if is_torch_cuda(v):
            assert v.numel() == n and v.is_contiguous()
            return v
        a = np.ascontiguousarray(np.asarray(v).reshape(-1), dtype=dtype)
        if a.size != n:
            raise _f2py_error(f"{name} must have {n} elements, got {a.size}")
        return a

    @staticmethod
    def _raster(m, dtype, name):
        """(ncols,nrows) raster -> flat f2py buffer, ncols, nrows"""
        if is_torch_cuda(m):
            assert m.dim() == 2 and m.is_contiguous()
            return m, m.shape[1], m.shape[0]
        a = np.asarray(m)
        if a.ndim != 2:
            raise _f2py_error(f"{name} must be a 2-D (ncols,nrows) array")
        return np.asfortranarray(a, dtype=dtype), a.shape[0], a.shape[1]

    @staticmethod
    def _new(like_resident, shape, dtype, device=None):
        if like_resident:
            import torch
            tdt = {np.float32: torch.float32, np.int32: torch.int32}[dtype]
            return torch.empty(shape, dtype=tdt, device=device)
        return np.empty(shape, dtype=dtype)

    # ---- functions (cumodule.c doc strings) ---------------------------------------------------
    def coord2fil_col(self, x, y):
        """col,fil = coord2fil_col(x,y)   -- cuencas.f90:77-90"""
        c = self._push_geo()
        col, fil = C.c_int32(), C.c_int32()
        c.check(c.L.wmf_coord2fil_col(c.h, float(x), float(y), C.byref(col), C.byref(fil)))
        return col.value, fil.value

    def basin_find(self, x, y, dir, nc=None, nr=None):
        """nceldas = basin_find(x,y,dir,[nc,nr])   -- cuencas.f90:525-591"""
        d, dnc, dnr = self._raster(dir, np.int32, "dir")
        if nc is not None and (int(nc), int(nr)) != (dnc, dnr):
            raise _f2py_error(f"dir is ({dnc},{dnr}) but nc,nr = ({nc},{nr})")
        c = self._push_geo()
        n = C.c_int32()
        c.check(c.L.wmf_basin_find(c.h, float(x), float(y), ptr(d), dnc, dnr, C.byref(n)))
        return int(n.value)

    def basin_find_depth(self):
        c = self.ctx
        d = C.c_int32()
        c.check(c.L.wmf_basin_find_depth(c.h, C.byref(d)))
        return d.value

    def basin_cut(self, nceldas, resident=False):
        """basin_f = basin_cut(nceldas)   -- cuencas.f90:592-607"""
        c = self.ctx
        n = int(nceldas)
        if resident:
            import torch
            out = torch.empty((n, 3), dtype=torch.int32, device=f"cuda:{c.device}")
            c.check(c.L.wmf_basin_cut(c.h, n, ptr(out)))
            return out
        out = np.empty((3, n), dtype=np.int32, order="F")
        c.check(c.L.wmf_basin_cut(c.h, n, ptr(out)))
        return out

    def basin_acum(self, basin_f, nceldas=None):
        """acum = basin_acum(basin_f,[nceldas])   -- cuencas.f90:659-680"""
        b, n, res = self._table(basin_f)
        c = self._push_geo()
        acum = self._new(res, (n,), np.int32, getattr(b, "device", None))
        c.check(c.L.wmf_basin_acum(c.h, ptr(b), n, ptr(acum)))
        self._pull_scalars(c)
        return acum

    def basin_acum_var(self, basin_f, var_to_acum, nceldas=None):
        """acum = basin_acum_var(basin_f,var_to_acum,[nceldas])   -- cuencas.f90:681-701 (basin_f is row 1)"""
        res = is_torch_cuda(basin_f)
        n = basin_f.numel() if res else np.asarray(basin_f).size
        d = self._vec(basin_f, n, np.int32, "basin_f")
        v = self._vec(var_to_acum, n, np.float32, "var_to_acum")
        c = self.ctx
        out = self._new(res, (1, n), np.float32, getattr(d, "device", None))
        if not res:
            out = np.asfortranarray(out)
        c.check(c.L.wmf_basin_acum_var(c.h, ptr(d), n, ptr(v), ptr(out)))
        return out

    def basin_basics(self, basin_f, dem, dir, nceldas=None):
        """acum,long_bn,pend,elev = basin_basics(basin_f,dem,dir,[nceldas])   -- cuencas.f90:608-658"""
        b, n, res = self._table(basin_f)
        dem = self._vec(dem, n, np.float32, "dem")
        dr = self._vec(dir, n, np.int32, "dir")
        c = self._push_geo()
        dev = getattr(b, "device", None)
        acum = self._new(res, (n,), np.int32, dev)
        lng, pend, elev = (self._new(res, (n,), np.float32, dev) for _ in range(3))
        c.check(c.L.wmf_basin_basics(c.h, ptr(b), ptr(dem), ptr(dr), n, ptr(acum), ptr(lng), ptr(pend), ptr(elev)))
        self._pull_scalars(c)
        return acum, lng, pend, elev

    def basin_coordxy(self, basin_f, nceldas=None):
        """x,y = basin_coordxy(basin_f,[nceldas])   -- cuencas.f90:797-808"""
        b, n, res = self._table(basin_f)
        c = self._push_geo()
        dev = getattr(b, "device", None)
        x, y = self._new(res, (n,), np.float32, dev), self._new(res, (n,), np.float32, dev)
        c.check(c.L.wmf_basin_coordxy(c.h, ptr(b), n, ptr(x), ptr(y)))
        return x, y

    def basin_map2basin(self, basin_f, mapa, xllm, yllm, dxm, dym, nodatam, nceldas=None, ncolsm=None, nrowsm=None):
        """vec = basin_map2basin(basin_f,mapa,xllm,yllm,dxm,dym,nodatam,[nceldas,ncolsm,nrowsm])   -- cuencas.f90:1144-1184"""
        b, n, res = self._table(basin_f)
        m, ncm, nrm = self._raster(mapa, np.float32, "mapa")
        c = self._push_geo()
        vec = self._new(res, (n,), np.float32, getattr(b, "device", None))
        c.check(c.L.wmf_basin_map2basin(c.h, ptr(b), n, ptr(m), float(xllm), float(yllm), ncm, nrm, float(dxm),
                                        float(dym), float(nodatam), ptr(vec)))
        return vec

    def basin_2map_find(self, basin, nceldas=None):
        """map_ncols,map_nrows = basin_2map_find(basin,[nceldas])   -- cuencas.f90:1185-1201"""
        b, n, _ = self._table(basin)
        c = self.ctx
        a, r = C.c_int32(), C.c_int32()
        c.check(c.L.wmf_basin_2map_find(c.h, ptr(b), n, C.byref(a), C.byref(r)))
        return a.value, r.value

    def basin_2map(self, basin, var, map_ncols, map_nrows, nceldas=None):
        """mapa,map_xll,map_yll = basin_2map(basin,var,map_ncols,map_nrows,[nceldas])   -- cuencas.f90:1202-1229"""
        b, n, res = self._table(basin)
        v = self._vec(var, n, np.float32, "var")
        c = self._push_geo()
        if res:
            mapa = self._new(True, (int(map_nrows), int(map_ncols)), np.float32, b.device)
        else:
            mapa = np.empty((int(map_ncols), int(map_nrows)), np.float32, order="F")
        xll, yll = C.c_float(), C.c_float()
        c.check(c.L.wmf_basin_2map(c.h, ptr(b), ptr(v), n, int(map_ncols), int(map_nrows), ptr(mapa), C.byref(xll), C.byref(yll)))
        return mapa, xll.value, yll.value

    def basin_float_var2map(self, basin_f, vect, nc, nf, nceldas=None):
        """mapa = basin_float_var2map(basin_f,vect,nc,nf,[nceldas])   -- cuencas.f90:1037-1051"""
        b, n, res = self._table(basin_f)
        v = self._vec(vect, n, np.float32, "vect")
        c = self._push_geo()
        if res:
            mapa = self._new(True, (int(nf), int(nc)), np.float32, b.device)
        else:
            mapa = np.empty((int(nc), int(nf)), np.float32, order="F")
        c.check(c.L.wmf_basin_float_var2map(c.h, ptr(b), ptr(v), n, int(nc), int(nf), ptr(mapa)))
        return mapa

    def basin_stream_mask(self, acum, umbral):
        """cauce,n_cauce: the channel-mask part of basin_stream_nod (cuencas.f90:1351-1353, acum >= umbral)"""
        res = is_torch_cuda(acum)
        n = acum.numel() if res else np.asarray(acum).size
        a = self._vec(acum, n, np.int32, "acum")
        c = self.ctx
        cauce = self._new(res, (n,), np.int32, getattr(a, "device", None))
        cnt = C.c_int32()
        c.check(c.L.wmf_basin_stream_mask(c.h, ptr(a), n, int(umbral), ptr(cauce), C.byref(cnt)))
        return cauce, cnt.value

    def basin_stream_nod(self, basin_f, acum, umbral, nceldas=None):
        """cauce,nodos,trazado,n_nodos,n_cauce = basin_stream_nod(basin_f,acum,umbral,[nceldas])   -- cuencas.f90:1340-1379"""
        b, n, res = self._table(basin_f)
        a = self._vec(acum, n, np.int32, "acum")
        c = self.ctx
        dev = getattr(b, "device", None)
        cauce, nodos, traz = (self._new(res, (n,), np.int32, dev) for _ in range(3))
        nn, ncau = C.c_int32(), C.c_int32()
        c.check(c.L.wmf_basin_stream_nod(c.h, ptr(b), ptr(a), n, int(umbral), ptr(cauce), ptr(nodos), ptr(traz),
                                         C.byref(nn), C.byref(ncau)))
        return cauce, nodos, traz, nn.value, ncau.value

    def basin_point2var(self, basin_f, id_coord, xy_coord, ncoord=None, nceldas=None):
        """res_coord,basin_pts = basin_point2var(basin_f,id_coord,xy_coord,[ncoord,nceldas])   -- cuencas.f90:1230-1261"""
        b, n, _ = self._table(basin_f)
        c = self._push_geo()
        ids = np.ascontiguousarray(np.asarray(id_coord).reshape(-1), np.int32)
        xy = np.asfortranarray(xy_coord, dtype=np.float32)
        if xy.ndim != 2 or xy.shape != (2, ids.size):
            raise _f2py_error(f"xy_coord must have shape (2,{ids.size}), got {xy.shape}")
        res, pts = np.empty(ids.size, np.int32), np.empty(n, np.int32)
        c.check(c.L.wmf_basin_point2var(c.h, ptr(b), ptr(ids), ptr(xy), ids.size, n, ptr(res), ptr(pts)))
        return res, pts

    def basin_stream_point2stream(self, basin_f, cauce, id_coord, xy_coord, ncoord=None, nceldas=None):
        """res_coord,basin_pts,xy_new = basin_stream_point2stream(basin_f,cauce,id_coord,xy_coord,[ncoord,nceldas])
        -- cuencas.f90:1524-1577 (set_Control -> Points_Points2Stream, wmf.py:2096-2103)"""
        b, n, _ = self._table(basin_f)
        c = self._push_geo()
        ca = self._vec(cauce, n, np.int32, "cauce")
        ids = np.ascontiguousarray(np.asarray(id_coord).reshape(-1), np.int32)
        xy = np.asfortranarray(xy_coord, dtype=np.float32)
        if xy.ndim != 2 or xy.shape != (2, ids.size):
            raise _f2py_error(f"xy_coord must have shape (2,{ids.size}), got {xy.shape}")
        res, pts = np.empty(ids.size, np.int32), np.empty(n, np.int32)
        xyn = np.empty((2, ids.size), np.float32, order="F")
        c.check(c.L.wmf_basin_stream_point2stream(c.h, ptr(b), ptr(ca), ptr(ids), ptr(xy), ids.size, n, ptr(res), ptr(pts), ptr(xyn)))
        return res, pts, xyn

    def basin_stream_sections(self, basin_f, cauces, directions, dem, num_celdas, nceldas=None, ncols=None, nrows=None):
        """secciones,secciones_cel = basin_stream_sections(basin_f,cauces,directions,dem,num_celdas,[nceldas,ncols,nrows])
        -- cuencas.f90:1464-1523 (GetGeo_Sections, wmf.py:1585-1587)"""
        b, n, _ = self._table(basin_f)
        c = self.ctx
        ca, di = self._vec(cauces, n, np.int32, "cauces"), self._vec(directions, n, np.int32, "directions")
        d, ncc, nrr = self._raster(dem, np.float32, "dem")
        S = 2 * int(num_celdas) + 1
        sec, cel = np.empty((S, n), np.float32, order="F"), np.empty((S, n), np.int32, order="F")
        c.check(c.L.wmf_basin_stream_sections(c.h, ptr(b), ptr(ca), ptr(di), ptr(d), int(num_celdas), n, ncc, nrr, ptr(sec), ptr(cel)))
        return sec, cel

    def stream_find_to_corr(self, x, y, dem, dir, cauce, nc=None, nf=None):
        """lat,lon = stream_find_to_corr(x,y,dem,dir,cauce,[nc,nf])   -- cuencas.f90:372-417 (wmf.py:2998-3000)"""
        c = self._push_geo()
        di, ncc, nrr = self._raster(dir, np.int32, "dir")
        ca, _, _ = self._raster(cauce, np.int32, "cauce")
        lat, lon = C.c_float(), C.c_float()
        c.check(c.L.wmf_stream_find_to_corr(c.h, float(x), float(y), ptr(di), ptr(ca), ncc, nrr, C.byref(lat), C.byref(lon)))
        return lat.value, lon.value

    # ---- sub-basin (hills) topology: called by SimuBasin.__init__ / set_Geomorphology (wmf.py:3019-3023, 3664-3676) ----
    def basin_subbasin_nod(self, basin_f, acum, umbral, nceldas=None):
        """cauce,nodos_fin,n_nodos = basin_subbasin_nod(basin_f,acum,umbral,[nceldas])   -- cuencas.f90:1835-1920"""
        b, n, res = self._table(basin_f)
        a = self._vec(acum, n, np.int32, "acum")
        c = self.ctx
        dev = getattr(b, "device", None)
        cauce, nodos = (self._new(res, (n,), np.int32, dev) for _ in range(2))
        nn = C.c_int32()
        c.check(c.L.wmf_basin_subbasin_nod(c.h, ptr(b), ptr(a), n, int(umbral), ptr(cauce), ptr(nodos), C.byref(nn)))
        return cauce, nodos, nn.value

    def basin_subbasin_cut(self, n_nodos):
        """sub_basins = basin_subbasin_cut(n_nodos)   -- cuencas.f90:1921-1930"""
        c = self.ctx
        out = np.empty((2, int(n_nodos)), np.int32, order="F")
        c.check(c.L.wmf_basin_subbasin_cut(c.h, int(n_nodos), ptr(out)))
        return out

    def basin_subbasin_find(self, basin_f, nodos, n_nodos, nceldas=None):
        """sub_pert,sub_basin = basin_subbasin_find(basin_f,nodos,n_nodos,[nceldas])   -- cuencas.f90:1979-2010 (the Fortran
        never assigns sub_basin: zeros)"""
        b, n, res = self._table(basin_f)
        nd = self._vec(nodos, n, np.int32, "nodos")
        c = self.ctx
        sub_pert = self._new(res, (n,), np.int32, getattr(b, "device", None))
        c.check(c.L.wmf_basin_subbasin_find(c.h, ptr(b), ptr(nd), n, ptr(sub_pert)))
        return sub_pert, np.zeros(int(n_nodos), np.int32)

    def basin_subbasin_horton(self, sub_basins, nceldas, n_nodos=None):
        """sub_horton,nod_horton = basin_subbasin_horton(sub_basins,nceldas,[n_nodos])   -- cuencas.f90:1931-1978"""
        sb = np.asfortranarray(sub_basins, dtype=np.int32)
        if sb.ndim != 2 or sb.shape[0] != 2:
            raise _f2py_error(f"sub_basins must have shape (2,n_nodos), got {sb.shape}")
        c = self.ctx
        sh, nh = np.empty(sb.shape[1], np.int32), np.empty(int(nceldas), np.int32)
        c.check(c.L.wmf_basin_subbasin_horton(c.h, ptr(sb), sb.shape[1], int(nceldas), ptr(sh), ptr(nh)))
        return sh, nh

    def basin_subbasin_long(self, sub_pert, cauce, long, sub_basin, sub_horton, n_nodos=None, nceldas=None):
        """sub_basin_long,max_long,nodo_max_long = basin_subbasin_long(sub_pert,cauce,long,sub_basin,sub_horton,[n_nodos,nceldas])
        -- cuencas.f90:2011-2051"""
        sp = np.ascontiguousarray(np.asarray(sub_pert).reshape(-1), np.int32)
        n = sp.size
        ca, lg = self._vec(cauce, n, np.int32, "cauce"), self._vec(long, n, np.float32, "long")
        sb = np.asfortranarray(sub_basin, dtype=np.int32)
        shn = self._vec(sub_horton, sb.shape[1], np.int32, "sub_horton")
        c = self.ctx
        out = np.empty(sb.shape[1], np.float32)
        mx, nmx = C.c_float(), C.c_int32()
        c.check(c.L.wmf_basin_subbasin_long(c.h, ptr(sp), ptr(ca), ptr(lg), ptr(sb), ptr(shn), sb.shape[1], n, ptr(out),
                                            C.byref(mx), C.byref(nmx)))
        return out, mx.value, nmx.value

    def basin_subbasin_stream_prop(self, sub_pert, cauce, long, slope, n_nodos, nceldas=None):
        """stream_slope,stream_long = basin_subbasin_stream_prop(sub_pert,cauce,long,slope,n_nodos,[nceldas])
        -- cuencas.f90:2093-2115"""
        sp = np.ascontiguousarray(np.asarray(sub_pert).reshape(-1), np.int32)
        n = sp.size
        ca = self._vec(cauce, n, np.int32, "cauce")
        lg, sl = self._vec(long, n, np.float32, "long"), self._vec(slope, n, np.float32, "slope")
        c = self.ctx
        ss, slo = np.empty(int(n_nodos), np.float32), np.empty(int(n_nodos), np.float32)
        c.check(c.L.wmf_basin_subbasin_stream_prop(c.h, ptr(sp), ptr(ca), ptr(lg), ptr(sl), n, int(n_nodos), ptr(ss), ptr(slo)))
        return ss, slo

    def basin_subbasin_map2subbasin(self, sub_pert, basin_var, n_nodos, cauce, sum_mean, nceldas=None):
        """subbasin_sum = basin_subbasin_map2subbasin(sub_pert,basin_var,n_nodos,cauce,sum_mean,[nceldas])   -- cuencas.f90:2052-2092"""
        sp = np.ascontiguousarray(np.asarray(sub_pert).reshape(-1), np.int32)
        n = sp.size
        v, ca = self._vec(basin_var, n, np.float32, "basin_var"), self._vec(cauce, n, np.int32, "cauce")
        c = self.ctx
        out = np.empty(int(n_nodos), np.float32)
        c.check(c.L.wmf_basin_subbasin_map2subbasin(c.h, ptr(sp), ptr(v), int(n_nodos), n, ptr(ca), int(sum_mean), ptr(out)))
        return out

    def basin_stream_slope(self, basin_f, basin_elev, basin_long, nodos, n_cauce, nceldas=None):
        """stream_s,stream_l = basin_stream_slope(basin_f,basin_elev,basin_long,nodos,n_cauce,[nceldas])   -- cuencas.f90:1380-1438"""
        b, n, res = self._table(basin_f)
        e, lg = self._vec(basin_elev, n, np.float32, "basin_elev"), self._vec(basin_long, n, np.float32, "basin_long")
        nd = self._vec(nodos, n, np.int32, "nodos")
        c = self.ctx
        self._push_geo()
        dev = getattr(b, "device", None)
        ss, sl = (self._new(res, (n,), np.float32, dev) for _ in range(2))
        c.check(c.L.wmf_basin_stream_slope(c.h, ptr(b), ptr(e), ptr(lg), ptr(nd), n, ptr(ss), ptr(sl)))
        return ss, sl

    def basin_stream_type(self, basin_f, acum, umbrales, numbrales=None, nceldas=None):
        """stream_types = basin_stream_type(basin_f,acum,umbrales,[numbrales,nceldas])   -- cuencas.f90:1439-1463"""
        res = is_torch_cuda(acum)
        n = acum.numel() if res else np.asarray(acum).size
        a = self._vec(acum, n, np.int32, "acum")
        u = np.ascontiguousarray(np.asarray(umbrales).reshape(-1), dtype=np.int32)
        c = self.ctx
        out = self._new(res, (n,), np.int32, getattr(a, "device", None))
        c.check(c.L.wmf_basin_stream_type(c.h, ptr(a), n, ptr(u), u.size, ptr(out)))
        return out

    def geo_acum_to_cauce(self, acum, umbral, ncols=None, nrows=None):
        """cauce = geo_acum_to_cauce(acum,umbral,[ncols,nrows])   -- cuencas.f90:2120-2129 (strict >)"""
        a, nc, nr = self._raster(acum, np.int32, "acum")
        c = self.ctx
        if is_torch_cuda(a):
            out = self._new(True, tuple(a.shape), np.int32, a.device)
        else:
            out = np.empty((nc, nr), np.int32, order="F")
        c.check(c.L.wmf_geo_acum_to_cauce(c.h, ptr(a), nc * nr, int(umbral), ptr(out)))
        return out

    def geo_hand(self, basin_f, basin_elev, basin_long, cauce, nceldas=None):
        """hand_model,hdnd_model,a_quien = geo_hand(basin_f,basin_elev,basin_long,cauce,[nceldas])   -- cuencas.f90:2131-2174"""
        b, n, res = self._table(basin_f)
        e = self._vec(basin_elev, n, np.float32, "basin_elev")
        l = self._vec(basin_long, n, np.float32, "basin_long")
        ca = self._vec(cauce, n, np.int32, "cauce")
        c = self.ctx
        dev = getattr(b, "device", None)
        hand, hdnd = self._new(res, (n,), np.float32, dev), self._new(res, (n,), np.float32, dev)
        aq = self._new(res, (n,), np.int32, dev)
        c.check(c.L.wmf_geo_hand(c.h, ptr(b), ptr(e), ptr(l), ptr(ca), n, ptr(hand), ptr(hdnd), ptr(aq)))
        return hand, hdnd, aq

    def geo_hand_global(self, dem, dir, red, nc=None, nr=None):
        """hand = geo_hand_global(dem,dir,red,[nc,nr])   -- cuencas.f90:2175-2219"""
        d, ncc, nrr = self._raster(dem, np.float32, "dem")
        di, _, _ = self._raster(dir, np.int32, "dir")
        r, _, _ = self._raster(red, np.int32, "red")
        c = self._push_geo()
        if is_torch_cuda(d):
            out = self._new(True, tuple(d.shape), np.float32, d.device)
        else:
            out = np.empty((ncc, nrr), np.float32, order="F")
        c.check(c.L.wmf_geo_hand_global(c.h, ptr(d), ptr(di), ptr(r), ncc, nrr, ptr(out)))
        return out

    def basin_time_to_out(self, basin_f, long_bn, speed, nceldas=None):
        """time = basin_time_to_out(basin_f,long_bn,speed,[nceldas])   -- cuencas.f90:724-756"""
        b, n, res = self._table(basin_f)
        l = self._vec(long_bn, n, np.float32, "long_bn")
        s = self._vec(speed, n, np.float32, "speed")
        c = self.ctx
        t = self._new(res, (n,), np.float32, getattr(b, "device", None))
        c.check(c.L.wmf_basin_time_to_out(c.h, ptr(b), ptr(l), ptr(s), n, ptr(t)))
        return t

    def basin_propagate(self, basin_f, var, nceldas=None):
        """var_prop = basin_propagate(basin_f,var,[nceldas])   -- cuencas.f90:1811-1833"""
        b, n, res = self._table(basin_f)
        v = self._vec(var, n, np.float32, "var")
        c = self.ctx
        out = self._new(res, (n,), np.float32, getattr(b, "device", None))
        c.check(c.L.wmf_basin_propagate(c.h, ptr(b), ptr(v), n, ptr(out)))
        return out

    def dir_reclass_rwatershed(self, mapa, nc=None, nr=None):
        """cuencas.f90:2220-2235 is a nine-entry recoding of r.watershed directions to keypad codes; it is
        I/O pre-processing and stays a numpy lookup on the host."""
        m = np.asarray(mapa)
        lut = {1: 9, 2: 8, 3: 7, 4: 4, 5: 1, 6: 2, 7: 3, 8: 6}
        out = np.full(m.shape, int(self._g["nodata"]), dtype=np.int32)
        for k, v in lut.items():
            out[m == k] = v
        return out

    def dir_reclass_opentopo(self, mapa, nc=None, nr=None):
        """cuencas.f90:2237-2252, same remark as dir_reclass_rwatershed."""
        m = np.asarray(mapa)
        lut = {3: 8, 2: 9, 1: 6, 8: 3, 7: 2, 6: 1, 5: 4, 4: 7}
        out = np.full(m.shape, int(self._g["nodata"]), dtype=np.int32)
        for k, v in lut.items():
            out[m == k] = v
        return out
