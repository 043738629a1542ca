"""CPU: the C-ABI library loads and exports every symbol the header declares; host-side shim logic (no GPU calls)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    if not os.path.exists(g.LIB):
        g.build()
    from wmf_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "wmf_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(wmf_[a-z0-9_]+)\s*\(", hdr)))
    assert declared == sorted(_lib.EXPORTS)
    L = _lib.load()
    for s in declared:
        assert hasattr(L, s), s
    assert L.wmf_abi_version() == 14


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from wmf_b200 import WmfError, cu
    with pytest.raises(WmfError):
        cu.basin_acum(np.zeros((3, 4), np.int32))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "wmf_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, re.M), f
                assert "wmf_oracle" not in src, f


def test_f2py_attribute_semantics():
    from wmf_b200 import CuModule, ModelsModule
    m = ModelsModule()
    m.control = np.zeros((1, 7))                       # float64 -> int32 copy (wmf.py:3045)
    assert m.control.dtype == np.int32 and m.control.flags["F_CONTIGUOUS"]
    src = np.ones((4, 7))
    m.h_coef = src
    src[:] = 5                                         # assignment copied: later edits of the source do not leak in
    assert m.h_coef[0, 0] == 1 and m.h_coef.dtype == np.float32
    m.h_coef[1] = 3.5                                  # reads alias module memory (wmf.py:3882)
    assert m.h_coef[1, 3] == np.float32(3.5)
    m.dt = 300
    assert isinstance(m.dt, np.float32) and m.dt == 300
    m.calc_niter = 7.0
    assert isinstance(m.calc_niter, np.int32) and m.calc_niter == 7
    assert m.v_coef is None                            # unallocated allocatable
    with pytest.raises(AttributeError):
        m.no_such_thing
    m.speed_type = [2, 2, 1]
    assert m.speed_type.tolist() == [2, 2, 1]
    with pytest.raises(ValueError):
        m.speed_type = [1, 2]
    c = CuModule()
    c.ncols, c.dxp = 12.0, 30
    assert isinstance(c.ncols, np.int32) and c.ncols == 12 and isinstance(c.dxp, np.float32)


def test_unsupported_switches_fail_loudly():
    from wmf_b200 import ModelsModule
    m = ModelsModule()
    m.drena = np.zeros((3, 2), np.int32)
    with pytest.raises(ValueError):
        m.shia_v1("x.bin", "x.hdr", np.ones(10), 1, 1, 1)      # calib(11) required (modelosv2.f90:201)


def test_switch_combinations_are_decomposed_into_passes(monkeypatch):
    """Host logic of `_models._switch_passes` / `_shia_v1_passes` (no GPU): which switches get a launch sequence of their own,
    that every pass runs with exactly the switches it owns, which return values come from which pass, and that the caller's
    switches are restored."""
    from wmf_b200 import ModelsModule
    m = ModelsModule()
    assert m._switch_passes() == []
    m.sim_sediments, m.save_storage, m.show_storage = 1, 1, 1
    assert m._switch_passes() == []                                  # the base group is one launch sequence
    m.separate_fluxes = 1
    assert m._switch_passes() == ["separate_fluxes"]
    m.separate_rain, m.sim_floods = 1, 1
    assert m._switch_passes() == ["separate_fluxes", "separate_rain", "sim_floods"]
    m.sim_sediments = m.save_storage = m.show_storage = 0
    assert m._switch_passes() == ["separate_rain", "sim_floods"]     # no base group: the first switch rides the first pass
    m.separate_rain = m.sim_floods = 0
    assert m._switch_passes() == []                                  # separate_fluxes alone
    # the passes themselves, with the single-pass body replaced by a recorder
    m.sim_slides, m.separate_rain, m.show_mean_speed = 1, 1, 1
    seen = []
    real = ModelsModule.shia_v1

    def fake(self, *a, **k):
        if self._switch_passes():
            return real(self, *a, **k)
        on = tuple(k for k in self._BASE_SWITCHES + self._SOLO_SWITCHES if int(getattr(self, k)) == 1)
        seen.append(on)
        return tuple(f"{i}:{'+'.join(on)}" for i in range(11))

    monkeypatch.setattr(ModelsModule, "shia_v1", fake)
    ret = m.shia_v1(None, None, np.ones(11), 1, 1, 4)
    assert seen == [("sim_slides", "show_mean_speed"), ("separate_fluxes",), ("separate_rain",)]
    assert ret[0] == "0:sim_slides+show_mean_speed" and ret[2] == "2:separate_fluxes" and ret[10] == "10:separate_rain"
    assert ret[1].endswith("sim_slides+show_mean_speed") and ret[9].endswith("sim_slides+show_mean_speed")
    assert (int(m.sim_slides), int(m.separate_fluxes), int(m.separate_rain), int(m.show_mean_speed), int(m.sim_floods)) == (1, 1, 1, 1, 0)


def test_rain_file_formats(tmp_path):
    from wmf_b200 import ModelsModule
    from wmf_b200._models import read_int_basin, read_rain_hdr, write_rain_files
    rng = np.random.default_rng(0)
    rain = rng.integers(0, 5000, (6, 11)).astype(np.int32)
    rain[0] = 0
    pos = np.array([1, 3, 3, 1, 6, 2, 1, 5], np.int32)
    b, h = str(tmp_path / "r.bin"), str(tmp_path / "r.hdr")
    write_rain_files(b, h, rain, pos)
    ids, p = read_rain_hdr(h, 8)
    assert p.tolist() == pos.tolist() and ids.tolist() == list(range(1, 9))
    v, res = read_int_basin(b, 3, 11)
    assert res == 0 and v.tolist() == rain[2].tolist()
    assert read_int_basin(b, 99, 11)[1] == 1
    # rain_first_point > 1 skips rain_first_point lines, not rain_first_point-1 (modelosv2.f90:989-993)
    _, p2 = read_rain_hdr(h, 3, rain_first_point=2)
    assert p2.tolist() == pos[2:5].tolist()
    with pytest.raises(ValueError):
        read_rain_hdr(h, 9)
    m = ModelsModule()
    table, remap = m._load_rain(b, h, 8, 11)
    assert table[0].max() == 0 and np.array_equal(table[remap - 1], rain[pos - 1])


def test_float_basin_records(tmp_path):
    from wmf_b200._models import read_float_basin_ncol, write_float_basin
    p = str(tmp_path / "s.StObin")
    a = np.arange(15, dtype=np.float32).reshape(5, 3, order="F")
    write_float_basin(p, a, 1)
    write_float_basin(p, a + 100, 2)
    write_float_basin(p, a - 1, 0)                         # record <= 0 writes nothing (modelosv2.f90:937)
    v, res = read_float_basin_ncol(p, 2, 3, 5)
    assert res == 0 and np.array_equal(v, a + 100)
    assert os.path.getsize(p) == 2 * 15 * 4


def test_synthetic_generators_are_deterministic_and_acyclic():
    from wmf_b200 import synth
    for kind in ("G1", "G2"):
        d1, e1 = synth.make_raster(kind, 48, 36, 3)
        d2, e2 = synth.make_raster(kind, 48, 36, 3, chunk_rows=5)
        assert np.array_equal(d1, d2) and np.array_equal(e1, e2)
        assert d1.shape == (48, 36) and d1.flags["F_CONTIGUOUS"]
        inner = d1[1:-1, 1:-1]
        assert inner.min() >= 1 and inner.max() <= 9 and not (inner == 5).any()
        assert (d1[0] == -9999).all() and (d1[:, -1] == -9999).all()
        # DEM strictly decreases along the flow for every interior cell except the outlet
        c, r = np.meshgrid(np.arange(1, 47), np.arange(1, 35), indexing="ij")
        code = d1[c, r]
        c2, r2 = c + (code - 1) % 3 - 1, r + 1 - (code - 1) // 3
        oc, orow = synth.outlet_colrow(48, 36)
        notout = ~((c + 1 == oc) & (r + 1 == orow))
        assert np.all(e1[c2, r2][notout] < e1[c, r][notout])


def test_every_export_has_ctypes_argtypes():
    """A C entry point called without argtypes gets its 64-bit pointers truncated to int: every export must declare them."""
    from wmf_b200 import _lib
    L = _lib.load()
    missing = [s for s in _lib.EXPORTS if getattr(L, s).argtypes is None and s != "wmf_abi_version"]
    assert not missing, missing


def test_every_f2py_module_attribute_exists():
    """modelsmodule.c:3721-3856 lists the 140 names the f2py module `models` exposes; each must be readable on the drop-in
    object (None for an unallocated array, as f2py).  Checked against the reference's generated wrapper when the tree is here."""
    import re
    path = "/root/reference/build/src.win-amd64-3.7/modelsmodule.c"
    if not os.path.exists(path):
        pytest.skip("reference tree not present")
    from wmf_b200 import ModelsModule
    lines = open(path, encoding="latin-1").read().splitlines()[3714:3860]
    names = re.findall(r'\{"([a-z_0-9]+)"', "\n".join(lines))
    assert len(names) >= 136
    m = ModelsModule()
    missing = []
    for n in names:
        try:
            getattr(m, n)
        except AttributeError:
            missing.append(n)
    assert not missing, missing
