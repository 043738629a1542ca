"""GPU parity, Path A (module cu): CUDA through the C ABI vs the CPU oracle, bit-exact for integer work."""
import numpy as np
import pytest

from conftest import basin_vectors, geo_to, make_basin

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cu():
    from wmf_b200 import cu
    return cu


def test_kat1(cu):
    """Hand-derived KAT-1 (cuencas.f90:525-591, 659-680)."""
    cu.ncols, cu.nrows, cu.xll, cu.yll, cu.dx, cu.dy, cu.dxp, cu.nodata = 5, 5, 0, 0, 1, 1, 1, -9999
    D = np.full((5, 5), -9999, np.int32)
    D[1, 1:4] = [2, 1, 2]
    D[2, 1:4] = [3, 2, 2]
    D[3, 1:4] = [6, 2, 4]
    n = cu.basin_find(2.5, 1.5, D.T, 5, 5)
    assert n == 9
    b = cu.basin_cut(n)
    assert b.tolist() == [[8, 5, 2, 2, 1, 1, 1, 1, 0], [4, 4, 3, 2, 4, 2, 3, 2, 3], [2, 3, 2, 2, 4, 4, 3, 3, 4]]
    assert cu.basin_acum(b, n).tolist() == [1, 2, 1, 1, 3, 1, 1, 3, 9]
    assert cu.coord2fil_col(2.5, 1.5) == (3, 4)


@pytest.mark.parametrize("kind,nc,nr,seed", [("G1", 64, 48, 0), ("G2", 96, 80, 1), ("G2", 512, 512, 0),
                                             ("G1", 512, 512, 2), ("G2", 700, 300, 2)])
def test_find_cut_acum_bit_exact(cu, kind, nc, nr, seed):
    bs = make_basin(kind, nc, nr, seed)
    geo_to(cu, bs)
    n = cu.basin_find(bs["x"], bs["y"], bs["DIR"], nc, nr)
    assert n == bs["n"] == (nc - 2) * (nr - 2)
    b = cu.basin_cut(n)
    assert b.shape == (3, n) and b.dtype == np.int32
    np.testing.assert_array_equal(b, bs["basin"])
    acum = cu.basin_acum(b, n)
    np.testing.assert_array_equal(acum, bs["cu"].basin_acum(bs["basin"]))
    assert acum[-1] == n
    assert float(cu.area) == pytest.approx(float(bs["cu"].area), rel=1e-7)


def test_find_geographic_coordinates(cu):
    """fp32 coordinate path with xll=-75.6, dx=1/3600 (SURVEY 8d adversarial case)."""
    bs = make_basin("G2", 200, 160, 0, dx=1.0 / 3600.0, xll=-75.6, yll=6.1)
    geo_to(cu, bs)
    assert cu.coord2fil_col(bs["x"], bs["y"]) == bs["cu"].coord2fil_col(bs["x"], bs["y"])
    n = cu.basin_find(bs["x"], bs["y"], bs["DIR"])
    assert n == bs["n"]
    np.testing.assert_array_equal(cu.basin_cut(n), bs["basin"])
    o = bs["cu"]
    for m in (bs["DEM"], bs["DIR"].astype(np.float32)):
        got = cu.basin_map2basin(bs["basin"], m, bs["xll"], bs["yll"], bs["dx"], bs["dx"], o.nodata)
        np.testing.assert_array_equal(got, o.basin_map2basin(bs["basin"], m, bs["xll"], bs["yll"], bs["dx"], bs["dx"], o.nodata))
    x, y = cu.basin_coordxy(bs["basin"])
    ox, oy = o.basin_coordxy(bs["basin"])
    np.testing.assert_array_equal(x, ox)
    np.testing.assert_array_equal(y, oy)


def test_outlet_outside_map(cu):
    bs = make_basin("G1", 32, 32, 0)
    geo_to(cu, bs)
    n = cu.basin_find(-1e6, bs["y"], bs["DIR"])
    assert n == 1 == bs["cu"].basin_find(-1e6, bs["y"], bs["DIR"])
    np.testing.assert_array_equal(cu.basin_cut(1), bs["cu"].basin_cut(1))


def test_cyclic_dir_is_an_error(cu):
    from wmf_b200 import WmfError
    bs = make_basin("G1", 32, 32, 0)
    geo_to(cu, bs)
    D = np.array(bs["DIR"], order="F", copy=True)
    oc, orow = (32 + 1) // 2, 31
    D[oc - 1, orow - 1] = 8          # outlet drains up ...
    D[oc - 1, orow - 2] = 2          # ... into a cell that drains back down
    with pytest.raises(WmfError):
        cu.basin_find(bs["x"], bs["y"], D)


@pytest.mark.parametrize("kind,nc,nr", [("G2", 96, 80), ("G1", 300, 200)])
def test_basics_map2basin_types(cu, kind, nc, nr):
    bs = make_basin(kind, nc, nr, 0)
    geo_to(cu, bs)
    o, b = bs["cu"], bs["basin"]
    v = basin_vectors(bs)
    dv = cu.basin_map2basin(b, bs["DIR"].astype(np.float32), 0, 0, 30, 30, o.nodata)
    np.testing.assert_array_equal(dv.astype(np.int32), v["dirv"])
    de = cu.basin_map2basin(b, bs["DEM"], 0, 0, 30, 30, o.nodata)
    np.testing.assert_array_equal(de, v["demv"])
    acum, lng, pend, elev = cu.basin_basics(b, de, dv.astype(np.int32))
    np.testing.assert_array_equal(acum, v["acum"])
    np.testing.assert_array_equal(lng, v["lng"])
    np.testing.assert_array_equal(pend, v["pend"])
    np.testing.assert_array_equal(elev, v["elev"])
    # scalars: area exact.  The two means are sequential fp32 sums in the reference (error ~ N*eps/2); the GPU
    # accumulates in fp64, so it is checked against the fp64 mean and the oracle against its own error bound.
    assert float(cu.area) == float(o.area)
    n = bs["n"]
    assert float(cu.pend_media) == pytest.approx(float(np.mean(v["pend"].astype(np.float64))), rel=1e-6)
    assert float(cu.elevacion) == pytest.approx(float(np.mean(de.astype(np.float64))), rel=1e-6)
    assert float(o.pend_media) == pytest.approx(float(cu.pend_media), rel=n * 6e-8)
    assert float(o.elevacion) == pytest.approx(float(cu.elevacion), rel=n * 6e-8)
    assert float(cu.centrox) == float(o.centrox) and float(cu.centroy) == float(o.centroy)
    for thr in ([30, 500], [5], []):
        np.testing.assert_array_equal(cu.basin_stream_type(b[0], acum, thr), o.basin_stream_type(b, acum, thr))
    cauce, ncau = cu.basin_stream_mask(acum, 40)
    oc, _, _, _, oncau = o.basin_stream_nod(b, acum, 40)
    np.testing.assert_array_equal(cauce, oc)
    assert ncau == oncau
    assert cu.basin_2map_find(b) == o.basin_2map_find(b)


def test_map2basin_nodata_fill_and_offset_map(cu):
    """A coarser, shifted map with nodata holes: gathered cells exact, mean-filled cells 1e-6 rel."""
    bs = make_basin("G2", 120, 90, 3)
    geo_to(cu, bs)
    o, b = bs["cu"], bs["basin"]
    rng = np.random.default_rng(0)
    m = rng.uniform(1, 5, (50, 40)).astype(np.float32)
    m[rng.random(m.shape) < 0.2] = o.nodata
    args = (-100.0, 200.0, 75.0, 75.0, o.nodata)
    got, ref = cu.basin_map2basin(b, m, *args), o.basin_map2basin(b, m, *args)
    np.testing.assert_allclose(got, ref, rtol=1e-6)
    vals, counts = np.unique(ref, return_counts=True)
    fill = vals[np.argmax(counts)]                 # the mean-fill value is by far the most frequent one
    np.testing.assert_array_equal(got[ref != fill], ref[ref != fill])   # gathered cells are bit-exact
    assert (got == o.nodata).any()                 # cells outside the shifted map keep nodata


def test_hand_propagate_acumvar_time(cu):
    bs = make_basin("G2", 160, 120, 4)
    geo_to(cu, bs)
    o, b = bs["cu"], bs["basin"]
    v = basin_vectors(bs)
    cauce = (v["acum"] >= 60).astype(np.int32)
    for got, ref in zip(cu.geo_hand(b, v["elev"], v["lng"], cauce), o.geo_hand(b, v["elev"], v["lng"], cauce)):
        np.testing.assert_array_equal(got, ref)
    rng = np.random.default_rng(1)
    var = rng.uniform(0, 3, bs["n"]).astype(np.float32)
    np.testing.assert_array_equal(cu.basin_acum_var(b[0], var), o.basin_acum_var(b[0], var))
    var2 = np.where(rng.random(bs["n"]) < 0.3, var, -var).astype(np.float32)
    np.testing.assert_array_equal(cu.basin_propagate(b, var2), o.basin_propagate(b, var2))
    speed = rng.uniform(0.1, 2, bs["n"]).astype(np.float32)
    np.testing.assert_array_equal(cu.basin_time_to_out(b, v["lng"], speed), o.basin_time_to_out(b, v["lng"], speed))
    # maps
    mc, mr = cu.basin_2map_find(b)
    gm, gx, gy = cu.basin_2map(b, var, mc, mr)
    om, ox, oy = o.basin_2map(b, var, mc, mr)
    np.testing.assert_array_equal(gm, om)
    assert (gx, gy) == (ox, oy)
    np.testing.assert_array_equal(cu.basin_float_var2map(b, var, bs["nc"], bs["nr"]),
                                  o.basin_float_var2map(b, var, bs["nc"], bs["nr"]))


def test_hand_global(cu):
    bs = make_basin("G2", 140, 100, 5)
    geo_to(cu, bs)
    o = bs["cu"]
    v = basin_vectors(bs)
    acum_map = o.basin_float_var2map(bs["basin"], v["acum"].astype(np.float32), bs["nc"], bs["nr"])
    red = np.where(acum_map == o.nodata, int(o.nodata), (acum_map > 50).astype(np.int32)).astype(np.int32)
    got, ref = cu.geo_hand_global(bs["DEM"], bs["DIR"], red), o.geo_hand_global(bs["DEM"], bs["DIR"], red)
    np.testing.assert_array_equal(got, ref)
    np.testing.assert_array_equal(cu.geo_acum_to_cauce(np.maximum(acum_map, 0).astype(np.int32), 50),
                                  o.geo_acum_to_cauce(np.maximum(acum_map, 0).astype(np.int32), 50))


@pytest.mark.parametrize("shape", [(300, 260, 3), (150, 210, 5)])
def test_hand_global_with_holes_vs_oracle_and_walk(cu, shape, monkeypatch):
    """geo_hand_global (cuencas.f90:2175-2219) on a raster several tiles wide with nodata holes in DIR, invalid direction
    codes, nodata cells in the network mask and a sparse random network: the tile pass + pointer jumping against the
    oracle's walk, and against the one-walk-per-cell kernel it replaced (WMF_HAND_ALGO=walk), bit for bit."""
    import oracle
    from wmf_b200 import synth
    nc, nr, seed = shape
    DIR, DEM = synth.make_raster("G2", nc, nr, seed)
    rng = np.random.default_rng(seed)
    DIR = np.array(DIR, copy=True)
    DIR[rng.random(DIR.shape) < 0.004] = int(synth.NODATA)
    DIR[rng.random(DIR.shape) < 0.001] = 5
    red = (rng.random(DIR.shape) < 0.01).astype(np.int32)
    red[rng.random(DIR.shape) < 0.002] = int(synth.NODATA)
    o = oracle.Cu()
    for c in (cu, o):
        c.ncols, c.nrows, c.xll, c.yll, c.dx, c.dy, c.dxp, c.nodata = nc, nr, 0.0, 0.0, 30.0, 30.0, 30.0, synth.NODATA
    DIR, red = np.asfortranarray(DIR), np.asfortranarray(red)
    want = o.geo_hand_global(DEM, DIR, red)
    got = cu.geo_hand_global(DEM, DIR, red)
    np.testing.assert_array_equal(got, want)
    assert (want != synth.NODATA).sum() > 0.3 * nc * nr and (want == synth.NODATA).sum() > 100
    for algo in ("walk", "jump_global"):
        monkeypatch.setenv("WMF_HAND_ALGO", algo)
        np.testing.assert_array_equal(cu.geo_hand_global(DEM, DIR, red), want)


def test_resident_pipeline(cu):
    """Device-resident tensors: same answers, no host copies of the big arrays."""
    import torch
    bs = make_basin("G2", 256, 200, 6)
    geo_to(cu, bs)
    d_dir = torch.from_numpy(np.ascontiguousarray(bs["DIR"].T)).cuda()   # (nrows,ncols)
    n = cu.basin_find(bs["x"], bs["y"], d_dir)
    assert n == bs["n"]
    b = cu.basin_cut(n, resident=True)
    assert b.is_cuda and tuple(b.shape) == (n, 3)
    np.testing.assert_array_equal(b.cpu().numpy().T, bs["basin"])
    acum = cu.basin_acum(b)
    np.testing.assert_array_equal(acum.cpu().numpy(), bs["cu"].basin_acum(bs["basin"]))


@pytest.mark.parametrize("nc,nr,seed,noise", [(64, 48, 0, 1.8), (200, 150, 1, 1.6), (333, 257, 2, 1.8), (512, 512, 3, 1.5),
                                               (90, 1200, 4, 1.4)])
def test_partial_basins_with_pits_and_holes(cu, nc, nr, seed, noise):
    """Random rasters where the outlet catches only part of the map (other pits, nodata holes, no nodata frame, up to 8
    cells draining into one): basin table, accumulation, basics and HAND equal the oracle's, bit for bit."""
    import oracle
    from wmf_b200 import synth
    DIR, (oc, orow) = synth.make_partial_raster(nc, nr, seed, noise=noise)
    dx = 30.0
    x, y = dx * (oc - 0.5), dx * (nr - orow + 0.5)
    ocu = oracle.Cu()
    for c in (cu, ocu):
        c.ncols, c.nrows, c.xll, c.yll, c.dx, c.dy, c.dxp, c.nodata = nc, nr, 0.0, 0.0, dx, dx, dx, synth.NODATA
    n = cu.basin_find(x, y, DIR, nc, nr)
    assert n == ocu.basin_find(x, y, DIR) and 10 < n < nc * nr
    b = cu.basin_cut(n)
    np.testing.assert_array_equal(b, ocu.basin_cut(n))
    acum = cu.basin_acum(b)
    np.testing.assert_array_equal(acum, ocu.basin_acum(b))
    rng = np.random.default_rng(seed)
    dem = (rng.random((nc, nr)) * 50 + 100).astype(np.float32)
    dv = cu.basin_map2basin(b, DIR.astype(np.float32), 0, 0, dx, dx, synth.NODATA).astype(np.int32)
    de = cu.basin_map2basin(b, dem, 0, 0, dx, dx, synth.NODATA)
    got = cu.basin_basics(b, de, dv)
    want = ocu.basin_basics(b, de, dv)
    for g, w in zip(got, want):
        np.testing.assert_array_equal(g, w)
    cauce = (acum >= 12).astype(np.int32)
    if cauce[-1]:
        for g, w in zip(cu.geo_hand(b, got[3], got[1], cauce), ocu.geo_hand(b, got[3], got[1], cauce)):
            np.testing.assert_array_equal(g, w)


@pytest.mark.parametrize("kind,nc,nr,umbral", [("G2", 200, 160, 30), ("G1", 150, 220, 12)])
def test_stream_nod(cu, kind, nc, nr, umbral):
    """basin_stream_nod (cuencas.f90:1340-1379): channel mask, hydrological nodes and tracing points, bit-exact."""
    bs = make_basin(kind, nc, nr, 3)
    geo_to(cu, bs)
    v = basin_vectors(bs)
    got = cu.basin_stream_nod(bs["basin"], v["acum"], umbral)
    want = bs["cu"].basin_stream_nod(bs["basin"], v["acum"], umbral)
    for g, w in zip(got[:3], want[:3]):
        np.testing.assert_array_equal(g, w)
    assert got[3:] == want[3:] and got[3] > 2 and got[4] > 10
