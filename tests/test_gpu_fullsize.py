"""Parity at BASELINE.json's full sizes through size-independent properties (no oracle: it would take minutes).

Path A (C3-like rasters): the basin table produced on the GPU is checked, cell by cell with torch ops on the device, to be
exactly the reference's FIFO breadth-first order (cuencas.f90:525-591): every cell's DIR code points at its parent, parents
come later, levels never increase along the table, the cells draining into one cell are contiguous and in the 3x3 scan
order (kf outer, kc inner), every raster cell appears once; acum satisfies acum(parent) = 1 + sum(acum(children)).
Path B (C2): different schedules of the same arithmetic (time blocks of 1 and 4 steps, the one-step-per-launch kernel)
must agree bit for bit; mass is conserved every step; a run split in two resumes exactly (StoOut / HspeedIn).
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _check_bfs_table(b, acum, d_dir, nc, nr):
    """b: (N,3) int32 device tensor = basin_f transposed; d_dir: (nr,nc) int32 device raster."""
    import torch
    N = b.shape[0]
    d, col, row = b[:, 0].long(), b[:, 1].long(), b[:, 2].long()
    idx = torch.arange(N, device=b.device)
    assert int(d[-1]) == 0 and bool((d[:-1] > 0).all())
    par = N - d[:-1]                                   # 0-based parent index of cells 0..N-2
    assert bool((par > idx[:-1]).all()) and bool((par < N).all())
    # every raster cell at most once
    lin = (row - 1) * nc + (col - 1)
    assert bool((lin >= 0).all()) and bool((lin < nc * nr).all())
    assert torch.unique(lin).numel() == N
    # DIR code of the cell == keypad code of the move to its parent: code = 5 - 3*dy + dx  (cuencas.f90:445, 2684-2699)
    dirv = d_dir.reshape(-1)[lin[:-1]].long()
    dx = col[par] - col[:-1]
    dy = row[par] - row[:-1]
    assert bool((dx.abs() <= 1).all()) and bool((dy.abs() <= 1).all())
    assert bool((dirv == 5 - 3 * dy + dx).all())
    # levels (hops to the outlet) never increase along the table; computed by pointer doubling
    lev = torch.ones(N, dtype=torch.long, device=b.device)
    lev[-1] = 0
    jmp = torch.cat([par, torch.tensor([N - 1], device=b.device)])
    for _ in range(40):
        lev = lev + torch.where(jmp != N - 1, lev[jmp], torch.zeros_like(lev))
        jmp2 = jmp[jmp]
        if bool((jmp2 == jmp).all()):
            break
        jmp = jmp2
    assert bool((lev[1:] <= lev[:-1]).all())
    # queue order: the table is the FIFO queue reversed, so the parents of consecutive cells never decrease in table
    # index (=> the cells draining into one cell are contiguous and parents are served first-in first-out), and
    # siblings follow the probe order of cuencas.f90:555-571 (kf = rows top to bottom, kc = columns left to right)
    assert bool((par[1:] >= par[:-1]).all())
    same = par[1:] == par[:-1]
    scan = (row[:-1] - row[par] + 1) * 3 + (col[:-1] - col[par] + 1)       # position of the child in the 3x3 window
    assert bool((scan[1:][same] < scan[:-1][same]).all())                   # the table is the queue reversed
    # accumulation: acum(parent) = 1 + sum of acum over its children
    s = torch.zeros(N, dtype=torch.long, device=b.device)
    s.index_add_(0, par, acum[:-1].long())
    assert bool((acum.long() == s + 1).all()) and int(acum[-1]) == N
    return int(lev.max()) + 1


@pytest.mark.parametrize("kind,size", [("G2", 4096), ("G1", 3000)])
def test_path_a_properties_at_scale(kind, size):
    import torch
    from wmf_b200 import CuModule, synth
    cu = CuModule()
    nc = nr = int(os.environ.get("WMF_TEST_PATHA_SIZE", size))
    DIR, _ = synth.make_raster(kind, nc, nr, seed=1)
    cu.ncols, cu.nrows, cu.xll, cu.yll, cu.dx, cu.dy, cu.dxp, cu.nodata = nc, nr, 0.0, 0.0, 30.0, 30.0, 30.0, synth.NODATA
    x, y = synth.outlet_xy(nc, nr, 0.0, 0.0, 30.0)
    d_dir = torch.from_numpy(np.ascontiguousarray(DIR.T)).cuda()
    n = cu.basin_find(x, y, d_dir)
    assert n == (nc - 2) * (nr - 2)
    b = cu.basin_cut(n, resident=True)
    acum = cu.basin_acum(b)
    depth = _check_bfs_table(b, acum, d_dir, nc, nr)
    assert depth == cu.basin_find_depth()


def _c2_inputs(T):
    import bench
    from wmf_b200 import CuModule
    cu = CuModule()
    wl = bench.build_workload("c2", cu)
    inp = dict(wl["inp"])
    inp["n_reg"] = T
    inp["pos_evento"] = inp["pos_evento"][:T]
    inp["evpserie"] = inp["evpserie"][:T]
    return wl, inp


def _run(inp, T, stoin=None, hspeedin=None, device_rain=None):
    from wmf_b200 import ModelsModule, synth
    m = ModelsModule()
    synth.apply_to_models(m, inp)
    n_cont = int(np.count_nonzero(inp["control"])) + 1
    out = m._run(np.asarray(inp["calib"], np.float32)[None, :], n_cont, 1, T, device_rain if device_rain is not None else inp["rain"],
                 inp["pos_evento"][:T], stoin, hspeedin)
    return out


def test_path_b_schedules_agree_bit_for_bit_on_c2(monkeypatch):
    """1 044 484 cells: K=4 time blocks, K=1, and the one-step-per-launch kernel run the same fp32 operations in the same
    order per cell, so Q, StoOut and the final speeds must be identical to the last bit; balance is conserved."""
    import torch
    T = 96
    wl, inp = _c2_inputs(T)
    d_rain = torch.from_numpy(inp["rain"]).cuda()
    ref = _run(inp, T, device_rain=d_rain)
    assert ref["q"].shape[1] == T and np.isfinite(ref["q"]).all() and ref["q"][0].max() > 0
    assert ref["stoout"].min() >= 0.0
    tot = float(ref["stoout"].astype(np.float64).sum())
    assert np.abs(ref["balance"]).max() <= 1e-6 * tot                      # mass conservation, every step
    for k in ("1", "0"):
        monkeypatch.setenv("WMF_SHIA_K", k)
        other = _run(inp, T, device_rain=d_rain)
        np.testing.assert_array_equal(other["q"], ref["q"])
        np.testing.assert_array_equal(other["stoout"], ref["stoout"])
        np.testing.assert_array_equal(other["hspeed_out"], ref["hspeed_out"])
        assert np.abs(other["balance"] - ref["balance"]).max() <= 2e-7 * tot
    monkeypatch.delenv("WMF_SHIA_K")
    # resume: 40 steps, then the remaining 56 from (StoOut, hspeed) == 96 steps in one go
    T1 = 40
    a = dict(inp); a["pos_evento"] = inp["pos_evento"][:T1]; a["evpserie"] = inp["evpserie"][:T1]
    r1 = _run(a, T1, device_rain=d_rain)
    b = dict(inp); b["pos_evento"] = inp["pos_evento"][T1:]; b["evpserie"] = inp["evpserie"][T1:]
    r2 = _run(b, T - T1, stoin=r1["stoout"], hspeedin=r1["hspeed_out"], device_rain=d_rain)
    np.testing.assert_array_equal(np.concatenate([r1["q"], r2["q"]], axis=1), ref["q"])
    np.testing.assert_array_equal(r2["stoout"], ref["stoout"])


# ---------------------------------------------------------------------------------------------------------------------
# Oracle parity at BASELINE.json's sizes (the oracle's -O3 build: same IEEE arithmetic, bit-identical to the parity build
# and to the reference's compiled Fortran, tests/test_reference_pin.py).  CPU side: 10-40 s per test.
# ---------------------------------------------------------------------------------------------------------------------
def _excess(got, want, rtol=1e-5, afloor=1e-7):
    """Worst |got-want| in units of the tolerance rtol*|want| + afloor*max|want| (<= 1 passes)."""
    got, want = np.asarray(got, np.float64).reshape(-1), np.asarray(want, np.float64).reshape(-1)
    assert got.shape == want.shape
    return float((np.abs(got - want) / (rtol * np.abs(want) + afloor * np.abs(want).max() + 1e-300)).max())


def test_c2_basin_200_steps_vs_oracle():
    """C2 (1 044 484 cells) x 200 steps through models.shia_v1 with host buffers against the oracle: Q at the outlet and
    the 16 control points, StoOut of every cell and tank within 1e-5; Acum_rain bit-exact."""
    import oracle
    from wmf_b200 import ModelsModule, synth
    T = 200
    wl, inp = _c2_inputs(T)
    m = ModelsModule()
    synth.apply_to_models(m, inp)
    m.set_rain(inp["rain"], inp["pos_evento"])
    n_cont = int(np.count_nonzero(inp["control"])) + 1
    ret = m.shia_v1(None, None, inp["calib"], n_cont, 1, T, None, None, wl["n"])
    ref = oracle.shia_v1(inp, fast=True)
    assert ref["q"][0].max() > 1.0
    assert _excess(ret[0], ref["q"]) <= 1.0, "Q"
    assert _excess(ret[9], ref["stoout"]) <= 1.0, "StoOut"
    np.testing.assert_array_equal(np.asarray(m.acum_rain).reshape(-1), ref["acum_rain"])


def test_path_a_8192_bit_exact_vs_oracle():
    """find + cut + acum on an 8192 x 8192 raster (67 M cells, 8190 BFS levels): bit-exact against the oracle's FIFO search."""
    import torch
    import oracle
    from wmf_b200 import CuModule, synth
    nc = nr = int(os.environ.get("WMF_TEST_PATHA_ORACLE_SIZE", 8192))
    DIR, _ = synth.make_raster("G2", nc, nr, seed=2)
    cu, oc = CuModule(), oracle.Cu(fast=True)
    for c in (cu, oc):
        c.ncols, c.nrows, c.xll, c.yll, c.dx, c.dy, c.dxp, c.nodata = nc, nr, 0.0, 0.0, 30.0, 30.0, 30.0, synth.NODATA
    x, y = synth.outlet_xy(nc, nr, 0.0, 0.0, 30.0)
    d_dir = torch.from_numpy(np.ascontiguousarray(DIR.T)).cuda()
    n = cu.basin_find(x, y, d_dir)
    b = cu.basin_cut(n, resident=True)
    acum = cu.basin_acum(b)
    n_o = oc.basin_find(x, y, DIR)
    assert n == n_o == (nc - 2) * (nr - 2)
    b_o = oc.basin_cut(n_o)
    assert bool((b.cpu().numpy().T == b_o).all()), "basin_f differs from the oracle"
    a_o = oc.basin_acum(b_o)
    assert bool((acum.cpu().numpy().reshape(-1) == np.asarray(a_o).reshape(-1)).all()), "acum differs from the oracle"


def test_c4_shape_ensemble_64_members_vs_oracle():
    """C4 shape: 64 calibration sets on the 260 100-cell basin x 200 steps in ONE launch sequence (member-minor layout);
    4 members re-run one by one by the oracle, hydrographs at every control point within 1e-5."""
    import torch
    import oracle
    from wmf_b200 import CuModule, ModelsModule, synth
    nc = nr = 512
    T, M = 200, 64
    cu = CuModule()
    DIR, DEM = synth.make_raster("G2", nc, nr, seed=0)
    cu.ncols, cu.nrows, cu.xll, cu.yll, cu.dx, cu.dy, cu.dxp, cu.nodata = nc, nr, 0.0, 0.0, 30.0, 30.0, 30.0, synth.NODATA
    x, y = synth.outlet_xy(nc, nr, 0.0, 0.0, 30.0)
    n = cu.basin_find(x, y, DIR, nc, nr)
    b = cu.basin_cut(n)
    dv = cu.basin_map2basin(b, DIR.astype(np.float32), 0.0, 0.0, 30.0, 30.0, synth.NODATA).astype(np.int32)
    de = cu.basin_map2basin(b, DEM, 0.0, 0.0, 30.0, 30.0, synth.NODATA)
    acum, lng, pend, _ = cu.basin_basics(b, de, dv)
    inp = synth.make_shia_inputs(b, acum, lng, pend, dxp=30.0, dt=3600.0, n_reg=T, seed=0, n_control=16)
    m = ModelsModule()
    synth.apply_to_models(m, inp)
    m.make_resident()
    d_rain = torch.from_numpy(inp["rain"]).cuda()
    n_cont = int(np.count_nonzero(inp["control"])) + 1
    calibs = synth.make_ensemble_calibs(M, seed=0)
    out = m._run(calibs, n_cont, 1, T, d_rain, inp["pos_evento"], device_out=True, want={"q"})
    q = out["q"].cpu().numpy().reshape(M, T, n_cont)
    for k in (0, 17, 40, 63):
        ref = oracle.shia_v1(dict(inp, calib=calibs[k]), fast=True)
        assert ref["q"][0].max() > 0
        assert _excess(q[k].T, ref["q"]) <= 1.0, f"member {k}"


def test_c5_basin_sediments_slides_48_steps_vs_oracle():
    """C5: the 2.09 M-cell 12 m basin, 5-minute steps, kinematic hillslope speeds, SED sediments + landslides, 48 steps:
    Q, StoOut, Qsed, the four sediment stores, VolERO / VolDEPo within 1e-5; landslide maps identical."""
    import bench
    import oracle
    from wmf_b200 import CuModule, ModelsModule, synth
    T = 48
    cu = CuModule()
    wl = bench.build_workload("c5", cu)
    b, n = wl["basin"], wl["n"]
    dv = cu.basin_map2basin(b, wl["DIR"].astype(np.float32), 0.0, 0.0, 12.0, 12.0, synth.NODATA).astype(np.int32)
    de = cu.basin_map2basin(b, wl["DEM"], 0.0, 0.0, 12.0, 12.0, synth.NODATA)
    acum, lng, pend, _ = cu.basin_basics(b, de, dv)
    inp = synth.make_shia_inputs(b, acum, lng, pend, dxp=12.0, dt=300.0, n_reg=T, seed=0, thresholds=(200, 2000),
                                 speed_type=(2, 2, 1), n_control=16, sediments=True, slides=True, rain_peak_mm=40.0)
    m = ModelsModule()
    synth.apply_to_models(m, inp)
    m.set_rain(inp["rain"], inp["pos_evento"])
    n_cont = int(np.count_nonzero(inp["control"])) + 1
    ret = m.shia_v1(None, None, inp["calib"], n_cont, 1, T, None, None, n)
    ref = oracle.shia_v1(inp, fast=True)
    assert ref["qsed"].max() > 0 and ref["voldepo"].max() > 0
    worst = {"q": _excess(ret[0], ref["q"]), "stoout": _excess(ret[9], ref["stoout"]), "qsed": _excess(ret[1], ref["qsed"])}
    for k in ("vs", "vd", "vsc", "vdc", "volero", "voldepo"):
        worst[k] = _excess(getattr(m, k), ref[k])
    assert max(worst.values()) <= 1.0, worst
    np.testing.assert_array_equal(np.asarray(m.sl_riskvector).reshape(-1), ref["sl_riskvector"])
    np.testing.assert_array_equal(np.asarray(m.sl_slideocurrence).reshape(-1), ref["sl_slideocurrence"])
    np.testing.assert_array_equal(np.asarray(m.sl_slidencelltime).reshape(-1), ref["sl_slidencelltime"])
