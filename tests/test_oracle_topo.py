"""CPU: the oracle's Path-A restatement against the hand-derived KAT and structural invariants."""
import json
import os

import numpy as np
import pytest

from conftest import basin_vectors, make_basin

HERE = os.path.dirname(os.path.abspath(__file__))


def test_kat1_golden():
    import oracle
    k = json.load(open(os.path.join(HERE, "golden", "kat1.json")))
    cu = oracle.Cu()
    cu.ncols, cu.nrows, cu.xll, cu.yll, cu.dx, cu.dy, cu.dxp, cu.nodata = (
        k["ncols"], k["nrows"], k["xll"], k["yll"], k["dx"], k["dx"], k["dxp"], k["nodata"])
    D = np.array(k["dir_rows_top_to_bottom"], np.int32).T          # (ncols,nrows) like wmf.py hands it over
    assert cu.coord2fil_col(*k["outlet_xy"]) == tuple(k["outlet_colfil"])
    n = cu.basin_find(*k["outlet_xy"], D)
    assert n == k["nceldas"]
    b = cu.basin_cut(n)
    assert b.tolist() == k["basin_f"]
    assert cu.basin_acum(b).tolist() == k["acum"]
    par = n - b[0]
    lev = np.zeros(n + 1, int)
    for i in range(n - 2, -1, -1):
        lev[i] = lev[par[i]] + 1
    assert lev[:n].tolist() == k["hop_level"]


@pytest.mark.parametrize("kind,nc,nr,seed", [("G1", 40, 30, 0), ("G2", 64, 48, 1), ("G2", 200, 150, 2), ("G1", 257, 129, 3)])
def test_structure_invariants(kind, nc, nr, seed):
    bs = make_basin(kind, nc, nr, seed)
    n, b = bs["n"], bs["basin"]
    assert n == (nc - 2) * (nr - 2)                         # generators: basin == all interior cells
    d = b[0]
    assert d[-1] == 0 and np.count_nonzero(d == 0) == 1     # single outlet, last
    par = n - d[:-1]                                         # 0-based parent index of cells 0..n-2
    assert np.all(par > np.arange(n - 1))                   # drenaid(i) > i
    lev = np.zeros(n, np.int64)
    for i in range(n - 2, -1, -1):
        lev[i] = lev[par[i]] + 1
    assert np.all(np.diff(lev) <= 0)                        # hop level non-increasing in i
    # children of a cell are contiguous and in 3x3 scan order
    starts = np.flatnonzero(np.diff(par) != 0) + 1
    runs = np.split(par, starts)
    assert len({r[0] for r in runs}) == len(runs)
    # every cell's (col,row) moves to its parent's (col,row) along its DIR code
    col, row = b[1].astype(int), b[2].astype(int)
    code = bs["DIR"][col - 1, row - 1]
    mx, my = (code - 1) % 3 - 1, 1 - (code - 1) // 3
    assert np.array_equal(col[:-1] + mx[:-1], col[par]) and np.array_equal(row[:-1] + my[:-1], row[par])
    acum = bs["cu"].basin_acum(b)
    assert acum[-1] == n and acum.min() == 1
    # acum == subtree sizes computed independently with numpy
    sizes = np.ones(n, np.int64)
    for i in range(n - 1):
        sizes[par[i]] += sizes[i]
    assert np.array_equal(acum, sizes)


def test_basics_hand_and_friends():
    bs = make_basin("G2", 90, 70, 4)
    o, b, n = bs["cu"], bs["basin"], bs["n"]
    v = basin_vectors(bs)
    assert np.all(v["pend"] > 0) and set(np.unique(v["lng"])) <= {np.float32(30.0), np.float32(30.0) * np.sqrt(np.float32(2.0))}
    assert o.area == pytest.approx(n * 900 / 1e6, rel=1e-6)
    par = n - b[0][:-1]
    np.testing.assert_array_equal(v["pend"][:-1], np.where(
        (s := np.abs(v["demv"][:-1] - v["demv"][par]) / v["lng"][:-1]) == 0, np.float32(0.001), s))
    assert v["pend"][-1] == v["pend"][-2]                   # outlet copies the previous cell (cuencas.f90:640)
    cauce = (v["acum"] >= 50).astype(np.int32)
    hand, hdnd, aq = o.geo_hand(b, v["elev"], v["lng"], cauce)
    assert np.all(hand[cauce == 1] == 0) and np.all(hdnd[cauce == 1] == 0)
    hs = np.flatnonzero(cauce == 0)
    assert np.all(aq[hs] > 0) and np.all(cauce[aq[hs] - 1] == 1) and np.all(hand[hs] > 0)
    assert np.all(hdnd[hs] >= v["lng"][hs])
    types = o.basin_stream_type(b, v["acum"], [30, 500])
    assert set(np.unique(types)) == {1, 2, 3}
    t = o.basin_time_to_out(b, v["lng"], np.ones(n, np.float32))
    assert t.argmax() < n // 2 and t[-1] == v["lng"][-1]
    prop = o.basin_propagate(b, np.ones(n, np.float32))
    np.testing.assert_array_equal(prop, v["acum"].astype(np.float32))
    av = o.basin_acum_var(b[0], np.ones(n, np.float32))
    np.testing.assert_array_equal(av[0], v["acum"].astype(np.float32))
    mc, mr = o.basin_2map_find(b)
    assert (mc, mr) == (bs["nc"] - 2, bs["nr"] - 2)
    m, xll, yll = o.basin_2map(b, v["acum"].astype(np.float32), mc, mr)
    assert m.max() == n and (xll, yll) == (30.0, 30.0)


def test_map2basin_identity_and_nodata_mean():
    bs = make_basin("G1", 50, 40, 5)
    o, b = bs["cu"], bs["basin"]
    got = o.basin_map2basin(b, bs["DEM"], 0, 0, 30, 30, o.nodata)
    np.testing.assert_array_equal(got, bs["DEM"][b[1] - 1, b[2] - 1])
    m = np.array(bs["DEM"], order="F", copy=True)
    hole = (b[1][7] - 1, b[2][7] - 1)
    m[hole] = o.nodata
    got = o.basin_map2basin(b, m, 0, 0, 30, 30, o.nodata)
    assert got[7] == pytest.approx(m[m != o.nodata].mean(), rel=1e-4)


@pytest.mark.parametrize("nc,nr,seed", [(24, 20, 0), (31, 17, 1), (40, 40, 2), (18, 33, 3)])
def test_two_independent_restatements_agree_on_partial_basins(nc, nr, seed):
    """oracle/topo_py.py (pure Python, written from cuencas.f90) vs the C oracle on random rasters with pits, nodata holes
    and a basin that covers only part of the map: identical basin tables and accumulations."""
    import oracle
    from oracle import topo_py
    from wmf_b200 import synth
    DIR, (oc, orow) = synth.make_partial_raster(nc, nr, seed)
    dx = 30.0
    x, y = dx * (oc - 0.5), dx * (nr - orow + 0.5)
    ocu = oracle.Cu()
    ocu.ncols, ocu.nrows, ocu.xll, ocu.yll, ocu.dx, ocu.dy, ocu.dxp, ocu.nodata = nc, nr, 0.0, 0.0, dx, dx, dx, synth.NODATA
    n = ocu.basin_find(x, y, DIR)
    b = ocu.basin_cut(n)
    bp = topo_py.basin_find_cut(x, y, DIR, 0.0, 0.0, dx)
    assert 10 < n < (nc * nr) and bp.shape == (3, n)          # a partial basin
    np.testing.assert_array_equal(bp, b)
    np.testing.assert_array_equal(topo_py.basin_acum(bp), ocu.basin_acum(b))
    assert (b[1, -1], b[2, -1]) == (oc, orow)
