"""ctypes binding of libwmf_b200.so (C ABI declared in include/wmf_b200.h).

The library is the product: there is no Python / numpy / CPU fallback.  If the shared
object has not been built (``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C wmf_b200/csrc``) or no CUDA device is present, the first call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# WMF_B200_LIB: load another build of the same library (tuning experiments with different launch bounds)
LIB_PATH = os.environ.get("WMF_B200_LIB") or os.path.join(_HERE, "libwmf_b200.so")

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int32)

WMF_OK, WMF_ERR_ARG, WMF_ERR_CUDA, WMF_ERR_TOPOLOGY, WMF_ERR_STATE, WMF_ERR_NOMEM = range(6)


class WmfError(RuntimeError):
    """A libwmf_b200 call returned a non-zero status (the f2py modules raised <module>.error)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"libwmf_b200 status {code}: {msg}")
        self.code = code


class ShiaCfg(C.Structure):
    _fields_ = [
        ("n_cel", C.c_int32), ("n_reg", C.c_int32), ("n_cont", C.c_int32), ("n_conth", C.c_int32),
        ("dt", C.c_float), ("dxp", C.c_float), ("retorno_gr", C.c_float), ("retorno_aq", C.c_float),
        ("storage_constant", C.c_float),
        ("speed_type", C.c_int32 * 3), ("calc_niter", C.c_int32),
        ("sim_sediments", C.c_int32), ("sim_slides", C.c_int32),
        ("show_storage", C.c_int32), ("show_speed", C.c_int32), ("show_mean_speed", C.c_int32),
        ("show_mean_retorno", C.c_int32), ("save_retorno", C.c_int32),
        ("sed_factor", C.c_float), ("wi", C.c_float * 3), ("diametro", C.c_float * 3),
        ("sl_gullienogullie", C.c_int32), ("sl_fs", C.c_float), ("sl_gammaw", C.c_float),
        ("n_members", C.c_int32),
        ("save_storage", C.c_int32), ("save_speed", C.c_int32), ("separate_fluxes", C.c_int32),
        ("save_vfluxes", C.c_int32), ("save_rc", C.c_int32), ("separate_rain", C.c_int32),
        ("sim_floods", C.c_int32), ("flood_sec_tam", C.c_int32),
        ("flood_dw", C.c_float), ("flood_dsed", C.c_float), ("flood_av", C.c_float), ("flood_cmax", C.c_float),
        ("flood_umbral", C.c_float), ("flood_step", C.c_float), ("flood_hmax", C.c_float),
        ("pos_per_member", C.c_int32),
    ]


class ShiaInputs(C.Structure):
    _fields_ = [
        ("drena", C.c_void_p), ("drena_stride", C.c_int32), ("unit_type", C.c_void_p),
        ("hill_long", C.c_void_p), ("hill_slope", C.c_void_p), ("stream_long", C.c_void_p),
        ("stream_slope", C.c_void_p), ("elem_area", C.c_void_p),
        ("v_coef", C.c_void_p), ("h_coef", C.c_void_p), ("h_exp", C.c_void_p),
        ("max_capilar", C.c_void_p), ("max_gravita", C.c_void_p), ("max_aquifer", C.c_void_p),
        ("evpserie", C.c_void_p), ("storage", C.c_void_p),
        ("control", C.c_void_p), ("control_h", C.c_void_p),
        ("calib", C.c_void_p), ("stoin", C.c_void_p), ("hspeedin", C.c_void_p),
        ("rain", C.c_void_p), ("n_records", C.c_int32), ("pos_evento", C.c_void_p),
        ("krus", C.c_void_p), ("crus", C.c_void_p), ("prus", C.c_void_p), ("parliac", C.c_void_p),
        ("vd_init", C.c_void_p),
        ("sl_zs", C.c_void_p), ("sl_gammas", C.c_void_p), ("sl_cohesion", C.c_void_p),
        ("sl_frictionangle", C.c_void_p), ("sl_radslope", C.c_void_p),
        ("guarda_cond", C.c_void_p), ("guarda_vfluxes", C.c_void_p),
        ("conv", C.c_void_p), ("stra", C.c_void_p), ("n_conv_records", C.c_int32), ("n_stra_records", C.c_int32),
        ("pos_conv", C.c_void_p), ("pos_stra", C.c_void_p),
        ("flood_w", C.c_void_p), ("flood_d50", C.c_void_p), ("flood_slope", C.c_void_p), ("flood_sections", C.c_void_p),
        ("flood_sec_cells", C.c_void_p),
    ]


_OUT_NAMES = [
    "q", "qsed", "hum", "st1", "st3", "balance", "speed", "areacontrol", "stoout", "hspeed_out",
    "mean_rain", "acum_rain", "mean_storage", "mean_speed", "mean_retorno", "retorned",
    "vs", "vd", "vsc", "vdc", "volero", "voldepo", "erot", "dept",
    "sl_zmin", "sl_zcrit", "sl_zmax", "sl_bo", "sl_riskvector", "sl_slideocurrence",
    "sl_slideacumulate", "sl_slidencelltime", "sto_snap", "speed_snap", "qseparated",
    "mean_vfluxes", "vfl_snap", "rc_coef", "ret_snap", "qsep_byrain", "storage_conv", "storage_stra",
    "flood_flood", "flood_h", "flood_ufr", "flood_cr", "flood_q", "flood_qsed", "flood_speed", "flood_profundidad", "flood_scalars",
]


class ShiaOutputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _OUT_NAMES]


# every symbol include/wmf_b200.h declares (tests/test_abi.py checks the header against this list)
EXPORTS = [
    "wmf_abi_version", "wmf_ctx_create", "wmf_ctx_destroy", "wmf_last_error", "wmf_set_stream",
    "wmf_synchronize", "wmf_launch_count", "wmf_set_geo", "wmf_get_basin_scalars",
    "wmf_coord2fil_col", "wmf_basin_find", "wmf_basin_cut", "wmf_basin_find_depth", "wmf_basin_acum",
    "wmf_basin_acum_var", "wmf_basin_basics", "wmf_basin_coordxy", "wmf_basin_map2basin",
    "wmf_basin_2map_find", "wmf_basin_2map", "wmf_basin_float_var2map", "wmf_basin_stream_mask", "wmf_basin_stream_nod",
    "wmf_basin_stream_type", "wmf_geo_acum_to_cauce", "wmf_geo_hand", "wmf_geo_hand_global",
    "wmf_basin_time_to_out", "wmf_basin_propagate", "wmf_shia_run", "wmf_eval_scores",
    "wmf_calc_speed", "wmf_last_kernel_ms", "wmf_last_shia_stats", "wmf_rain_idw", "wmf_rain_mit", "wmf_rain_pre_mit",
    "wmf_basin_subbasin_nod", "wmf_basin_subbasin_cut", "wmf_basin_subbasin_find", "wmf_basin_subbasin_horton",
    "wmf_basin_subbasin_long", "wmf_basin_subbasin_stream_prop", "wmf_basin_subbasin_map2subbasin", "wmf_basin_stream_slope", "wmf_stream_find_to_corr",
    "wmf_basin_point2var", "wmf_basin_stream_point2stream", "wmf_basin_stream_sections",
]

_lib = None


def load() -> C.CDLL:
    """Load libwmf_b200.so; raise (never fall back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is not built. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). wmf_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    L.wmf_last_error.restype = C.c_char_p
    L.wmf_last_error.argtypes = [C.c_void_p]
    L.wmf_launch_count.restype = C.c_int64
    L.wmf_launch_count.argtypes = [C.c_void_p]
    L.wmf_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    L.wmf_ctx_destroy.argtypes = [C.c_void_p]
    L.wmf_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    L.wmf_synchronize.argtypes = [C.c_void_p]
    L.wmf_set_geo.argtypes = [C.c_void_p, C.c_int32, C.c_int32] + [C.c_float] * 6
    L.wmf_get_basin_scalars.argtypes = [C.c_void_p, f32p]
    L.wmf_coord2fil_col.argtypes = [C.c_void_p, C.c_float, C.c_float, i32p, i32p]
    vp = C.c_void_p
    i32, f32, i64 = C.c_int32, C.c_float, C.c_int64
    L.wmf_basin_find.argtypes = [vp, f32, f32, vp, i32, i32, i32p]
    L.wmf_basin_cut.argtypes = [vp, i32, vp]
    L.wmf_basin_find_depth.argtypes = [vp, i32p]
    L.wmf_basin_acum.argtypes = [vp, vp, i32, vp]
    L.wmf_basin_acum_var.argtypes = [vp, vp, i32, vp, vp]
    L.wmf_basin_basics.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp, vp]
    L.wmf_basin_coordxy.argtypes = [vp, vp, i32, vp, vp]
    L.wmf_basin_map2basin.argtypes = [vp, vp, i32, vp, f32, f32, i32, i32, f32, f32, f32, vp]
    L.wmf_basin_2map_find.argtypes = [vp, vp, i32, i32p, i32p]
    L.wmf_basin_2map.argtypes = [vp, vp, vp, i32, i32, i32, vp, f32p, f32p]
    L.wmf_basin_float_var2map.argtypes = [vp, vp, vp, i32, i32, i32, vp]
    L.wmf_basin_stream_mask.argtypes = [vp, vp, i32, i32, vp, i32p]
    L.wmf_basin_stream_nod.argtypes = [vp, vp, vp, i32, i32, vp, vp, vp, i32p, i32p]
    L.wmf_basin_stream_type.argtypes = [vp, vp, i32, vp, i32, vp]
    L.wmf_geo_acum_to_cauce.argtypes = [vp, vp, i64, i32, vp]
    L.wmf_geo_hand.argtypes = [vp, vp, vp, vp, vp, i32, vp, vp, vp]
    L.wmf_geo_hand_global.argtypes = [vp, vp, vp, vp, i32, i32, vp]
    L.wmf_basin_time_to_out.argtypes = [vp, vp, vp, vp, i32, vp]
    L.wmf_basin_propagate.argtypes = [vp, vp, vp, i32, vp]
    L.wmf_shia_run.argtypes = [vp, C.POINTER(ShiaCfg), C.POINTER(ShiaInputs), C.POINTER(ShiaOutputs)]
    L.wmf_eval_scores.argtypes = [vp, vp, i32, i32, i32, i32, vp, vp]
    L.wmf_calc_speed.argtypes = [vp, f32, f32, f32, f32, f32, i32, f32p, f32p]
    L.wmf_last_kernel_ms.argtypes = [vp, f32p]
    L.wmf_last_shia_stats.argtypes = [vp, i32p, i32p]
    L.wmf_basin_subbasin_nod.argtypes = [vp, vp, vp, i32, i32, vp, vp, i32p]
    L.wmf_basin_subbasin_cut.argtypes = [vp, i32, vp]
    L.wmf_basin_subbasin_find.argtypes = [vp, vp, vp, i32, vp]
    L.wmf_basin_subbasin_horton.argtypes = [vp, vp, i32, i32, vp, vp]
    L.wmf_basin_subbasin_long.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, vp, f32p, i32p]
    L.wmf_basin_subbasin_stream_prop.argtypes = [vp, vp, vp, vp, vp, i32, i32, vp, vp]
    L.wmf_basin_subbasin_map2subbasin.argtypes = [vp, vp, vp, i32, i32, vp, i32, vp]
    L.wmf_basin_stream_slope.argtypes = [vp, vp, vp, vp, vp, i32, vp, vp]
    L.wmf_basin_point2var.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp]
    L.wmf_basin_stream_point2stream.argtypes = [vp, vp, vp, vp, vp, i32, i32, vp, vp, vp]
    L.wmf_basin_stream_sections.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp, vp]
    L.wmf_stream_find_to_corr.argtypes = [vp, f32, f32, vp, vp, i32, i32, f32p, f32p]
    L.wmf_rain_pre_mit.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp]
    L.wmf_rain_idw.argtypes = [vp, vp, vp, vp, f32, i32, i32, i32, f32, vp, i32, vp, vp, i32p, vp, i32]
    L.wmf_rain_mit.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, f32, vp, i32, vp, vp, i32p]
    _lib = L
    return L


def is_torch_cuda(x) -> bool:
    return type(x).__module__.startswith("torch") and getattr(x, "is_cuda", False)


def ptr(x) -> int | None:
    """Raw address of a numpy array (host) or a torch CUDA tensor (device); None passes NULL."""
    if x is None:
        return None
    if is_torch_cuda(x):
        assert x.is_contiguous()
        return x.data_ptr()
    assert isinstance(x, np.ndarray), type(x)
    return x.ctypes.data


class Context:
    """Owner of one ``wmf_ctx`` (one GPU, one stream)."""

    def __init__(self, device: int | None = None):
        self.L = load()
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        h = C.c_void_p()
        rc = self.L.wmf_ctx_create(int(device), C.byref(h))
        if rc != WMF_OK or not h.value:
            raise WmfError(rc, f"wmf_ctx_create(device={device}) failed: no usable CUDA device (no CPU fallback)")
        self.h = h
        self.device = int(device)

    def check(self, rc: int) -> None:
        if rc != WMF_OK:
            raise WmfError(rc, self.L.wmf_last_error(self.h).decode(errors="replace"))

    def use_stream(self, cuda_stream: int | None) -> None:
        self.check(self.L.wmf_set_stream(self.h, C.c_void_p(cuda_stream) if cuda_stream else None))

    def synchronize(self) -> None:
        self.check(self.L.wmf_synchronize(self.h))

    @property
    def launches(self) -> int:
        return int(self.L.wmf_launch_count(self.h))

    def last_kernel_ms(self) -> float:
        v = C.c_float()
        self.check(self.L.wmf_last_kernel_ms(self.h, C.byref(v)))
        return float(v.value)

    def last_shia_stats(self):
        a, b = C.c_int32(), C.c_int32()
        self.check(self.L.wmf_last_shia_stats(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def close(self) -> None:
        if getattr(self, "h", None) is not None and self.h.value:
            self.L.wmf_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx: Context | None = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx
