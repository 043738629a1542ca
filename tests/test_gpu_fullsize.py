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
