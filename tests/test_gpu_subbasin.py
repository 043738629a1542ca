"""GPU parity, sub-basin (hills) topology: basin_subbasin_nod / cut / find / horton / long / stream_prop / map2subbasin and
basin_stream_slope (cuencas.f90:1380-1438, 1835-2118) through the C ABI vs the CPU oracle, which is bit-identical to the
reference's compiled Fortran on them (tests/test_reference_pin.py).  Everything is bit-exact: the integer tables, and the
fp32 sums because the kernels add in the reference's order."""
import numpy as np
import pytest

from conftest import basin_vectors, geo_to, make_basin

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,nc,nr,seed,thr", [("G2", 96, 80, 1, 30), ("G1", 70, 60, 2, 12), ("G2", 300, 220, 3, 60),
                                                 ("G2", 120, 90, 4, 5), ("G2", 400, 300, 5, 4000)])
def test_subbasin_chain_as_simubasin_drives_it(kind, nc, nr, seed, thr):
    import oracle
    from wmf_b200 import CuModule
    bs = make_basin(kind, nc, nr, seed)
    b, n = bs["basin"], bs["n"]
    v = basin_vectors(bs)
    acum, lng, pend, elev = v["acum"], v["lng"], v["pend"], v["elev"]
    oc = bs["cu"]
    cu = CuModule()
    geo_to(cu, bs)
    # SimuBasin.__init__ (wmf.py:3019-3023)
    o1 = oc.basin_subbasin_nod(b, acum, thr)
    g1 = cu.basin_subbasin_nod(b, acum, thr, n)
    assert g1[2] == o1[2]
    nn = o1[2]
    np.testing.assert_array_equal(g1[0], o1[0])
    np.testing.assert_array_equal(g1[1], o1[1])
    of = oc.basin_subbasin_find(b, o1[1], nn)
    gf = cu.basin_subbasin_find(b, g1[1], nn, n)
    np.testing.assert_array_equal(gf[0], of[0])
    osb, gsb = oc.basin_subbasin_cut(nn), cu.basin_subbasin_cut(nn)
    np.testing.assert_array_equal(gsb, osb)
    with pytest.raises(Exception):
        cu.basin_subbasin_cut(nn)                      # sub_basins_temp is deallocated by the cut (cuencas.f90:1929)
    # set_Geomorphology (wmf.py:3664-3676)
    cauce, nodos, _, _, n_cauce = oc.basin_stream_nod(b, acum, thr)
    os_, gs = oc.basin_stream_slope(b, elev, lng, nodos, n_cauce), cu.basin_stream_slope(b, elev, lng, nodos, n_cauce, n)
    np.testing.assert_array_equal(gs[0], os_[0])
    np.testing.assert_array_equal(gs[1], os_[1])
    oh, gh = oc.basin_subbasin_horton(osb, n), cu.basin_subbasin_horton(gsb, n, nn)
    np.testing.assert_array_equal(gh[0], oh[0])
    np.testing.assert_array_equal(gh[1], oh[1])
    ol = oc.basin_subbasin_long(of[0], o1[0], lng, osb, oh[0])
    gl = cu.basin_subbasin_long(gf[0], g1[0], lng, gsb, gh[0], nn, n)
    np.testing.assert_array_equal(gl[0], ol[0])
    assert gl[1:] == ol[1:]
    op = oc.basin_subbasin_stream_prop(of[0], o1[0], lng, pend, nn)
    gp = cu.basin_subbasin_stream_prop(gf[0], g1[0], lng, pend, nn, n)
    np.testing.assert_array_equal(gp[0], op[0])
    np.testing.assert_array_equal(gp[1], op[1])
    for sm in (0, 1, 2):
        np.testing.assert_array_equal(cu.basin_subbasin_map2subbasin(gf[0], elev, nn, g1[0], sm, n),
                                      oc.basin_subbasin_map2subbasin(of[0], elev, nn, o1[0], sm))


def test_subbasin_against_the_reference_itself():
    from oracle import ref
    from wmf_b200 import CuModule
    if not ref.available():
        pytest.skip("oracle/_ref did not travel")
    bs = make_basin("G2", 260, 200, 8)
    b, n = bs["basin"], bs["n"]
    v = basin_vectors(bs)
    rc, cu = ref.RefCu(), CuModule()
    geo_to(rc, bs)
    geo_to(cu, bs)
    r1, g1 = rc.basin_subbasin_nod(b, v["acum"], 40), cu.basin_subbasin_nod(b, v["acum"], 40)
    assert r1[2] == g1[2]
    np.testing.assert_array_equal(g1[1], r1[1])
    rf, gf = rc.basin_subbasin_find(b, r1[1], r1[2]), cu.basin_subbasin_find(b, g1[1], g1[2])
    np.testing.assert_array_equal(gf[0], rf[0])
    np.testing.assert_array_equal(cu.basin_subbasin_cut(g1[2]), rc.basin_subbasin_cut(r1[2]))


def test_shia_on_the_hills_topology():
    """modelType = 'hills' (wmf.py:3696-3708): the model elements are the sub-basins, every one a channel element
    (unit_type = 3), drained through sub_basins(2,:).  The hills table built on the GPU feeds the same SHIA kernels."""
    import oracle
    from wmf_b200 import CuModule, ModelsModule, synth
    bs = make_basin("G2", 240, 200, 9)
    b, n = bs["basin"], bs["n"]
    v = basin_vectors(bs)
    cu = CuModule()
    geo_to(cu, bs)
    cauce, nodos, nn = cu.basin_subbasin_nod(b, v["acum"], 25, n)
    hills_own, _ = cu.basin_subbasin_find(b, nodos, nn, n)
    hills = cu.basin_subbasin_cut(nn)
    sub_horton, _ = cu.basin_subbasin_horton(hills, n, nn)
    sub_long, max_long, _ = cu.basin_subbasin_long(hills_own, cauce, v["lng"], hills, sub_horton, nn, n)
    s_slope, s_long = cu.basin_subbasin_stream_prop(hills_own, cauce, v["lng"], v["pend"], nn, n)
    assert nn > 100 and max_long > 0
    structure = np.zeros((3, nn), np.int32, order="F")
    structure[0] = hills[1]                                               # models.drena = hills[1] (wmf.py:3698)
    cells = np.bincount(hills_own, minlength=nn + 1)[1:][::-1]            # cells of the hill in column i (hill id nn-i)
    length = np.maximum(sub_long, 30.0).astype(np.float32)
    slope = np.nan_to_num(s_slope[::-1], nan=0.01).astype(np.float32)
    inp = synth.make_shia_inputs(structure, np.full(nn, 10 ** 6), length, np.maximum(slope, 1e-3), dxp=30.0, n_reg=80, seed=9,
                                 n_control=8)
    inp["elem_area"] = (cells * 30.0 ** 2).astype(np.float32)[None, :]
    assert (np.asarray(inp["unit_type"]) == 3).all()
    ref = oracle.shia_v1(inp)
    m = ModelsModule()
    synth.apply_to_models(m, inp)
    m.set_rain(inp["rain"], inp["pos_evento"])
    n_cont = int(np.count_nonzero(inp["control"])) + 1
    ret = m.shia_v1(None, None, inp["calib"], n_cont, 1, 80, None, None, nn)
    for got, want, nm in ((ret[0], ref["q"], "Q"), (ret[9], ref["stoout"], "StoOut")):
        got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
        assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max(), nm
    assert ref["q"][0].max() > 0


def test_stream_find_to_corr():
    """cuencas.f90:372-417: the outlet moved down the flow directions onto the channel map (SimuBasin(useCauceMap=...))."""
    from wmf_b200 import CuModule
    bs = make_basin("G2", 120, 90, 4)
    v = basin_vectors(bs)
    oc = bs["cu"]
    cu = CuModule()
    geo_to(cu, bs)
    amap = oc.basin_float_var2map(bs["basin"], v["acum"].astype(np.float32), 120, 90)
    cauce = np.asfortranarray((amap > 40).astype(np.int32))
    rng = np.random.default_rng(0)
    moved = 0
    for _ in range(40):
        c, f = int(rng.integers(2, 119)), int(rng.integers(2, 89))
        x, y = 30.0 * (c - 0.5), 30.0 * (90 - f + 0.5)
        want = oc.stream_find_to_corr(x, y, bs["DEM"], bs["DIR"], cauce)
        assert cu.stream_find_to_corr(x, y, bs["DEM"], bs["DIR"], cauce, 120, 90) == want
        moved += want != (np.float32(x), np.float32(y))
    assert moved > 20


def test_control_point_placement():
    """basin_point2var / basin_stream_point2stream (cuencas.f90:1230-1261, 1524-1577) as set_Control uses them."""
    from wmf_b200 import CuModule
    bs = make_basin("G2", 120, 90, 4)
    v = basin_vectors(bs)
    oc = bs["cu"]
    cu = CuModule()
    geo_to(cu, bs)
    cauce = (v["acum"] >= 40).astype(np.int32)
    rng = np.random.default_rng(1)
    xy = np.asfortranarray(np.vstack([rng.uniform(-100, 3700, 60), rng.uniform(-100, 2800, 60)]), dtype=np.float32)
    xy[:, 7] = xy[:, 3]                                   # two points on one cell: the later id wins
    ids = np.arange(101, 161, dtype=np.int32)
    want = oc.basin_stream_point2stream(bs["basin"], cauce, ids, xy)
    got = cu.basin_stream_point2stream(bs["basin"], cauce, ids, xy, 60, bs["n"])
    for g, w in zip(got, want):
        np.testing.assert_array_equal(g, w)
    assert np.bincount(want[0], minlength=3)[2] > 30 and np.bincount(want[0], minlength=3)[0] > 0
    for g, w in zip(cu.basin_point2var(bs["basin"], ids, xy, 60, bs["n"]), oc.basin_point2var(bs["basin"], ids, xy)):
        np.testing.assert_array_equal(g, w)


def test_stream_sections():
    """basin_stream_sections (cuencas.f90:1464-1523): the HydroFlash cross sections, bit-exact."""
    from wmf_b200 import CuModule
    bs = make_basin("G2", 90, 70, 6)
    v = basin_vectors(bs)
    cu = CuModule()
    geo_to(cu, bs)
    cauce = (v["acum"] >= 30).astype(np.int32)
    want = bs["cu"].basin_stream_sections(bs["basin"], cauce, v["dirv"], bs["DEM"], 4)
    got = cu.basin_stream_sections(bs["basin"], cauce, v["dirv"], bs["DEM"], 4, bs["n"], 90, 70)
    np.testing.assert_array_equal(got[0], want[0])
    np.testing.assert_array_equal(got[1], want[1])
