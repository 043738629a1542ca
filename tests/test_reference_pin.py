"""CPU: the oracle pinned against THE REFERENCE ITSELF.

``oracle/_ref/libwmf_ref.so`` holds the reference's own compiled Fortran (the COFF objects of its last f2py build, loaded
by ``oracle/ref_loader.c``; see ``oracle/ref.py``).  Every routine of SURVEY.md 8(a) is run through the reference's machine
code and through the C restatement in ``oracle/`` on the same inputs, and the results must be BIT-IDENTICAL: both are
IEEE fp32 in the same operation order with glibc's libm underneath.  That is what lets the GPU parity tests use the
restatement (which also offers fp64 diagnostics and runs at any size) as the checker.
"""
import json
import os

import numpy as np
import pytest

from conftest import basin_vectors, make_basin

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def ref():
    from oracle import ref as r
    if not r.available() and not r.build():
        pytest.skip("oracle/_ref/libwmf_ref.so is not built and /root/reference is absent")
    r.lib()
    return r


def same(a, b, name):
    a, b = np.asarray(a), np.asarray(b)
    assert a.reshape(-1).shape == b.reshape(-1).shape, (name, a.shape, b.shape)
    np.testing.assert_array_equal(a.reshape(-1), b.reshape(-1), err_msg=name)


def pair(ref, bs):
    import oracle
    rc, oc = ref.RefCu(), oracle.Cu()
    for c in (rc, oc):
        o = bs["cu"]
        c.ncols, c.nrows, c.xll, c.yll, c.dx, c.dy, c.dxp, c.nodata = o.ncols, o.nrows, o.xll, o.yll, o.dx, o.dy, o.dxp, o.nodata
    return rc, oc


def test_reference_objects_match_the_sources(ref):
    """The objects carry gfortran's run-time messages "At line N of file wmf/<src>.f90" for their allocate / deallocate
    statements: the same lines of the sources in the tree must hold such statements (checked when the tree is here)."""
    src_dir = "/root/reference/wmf"
    if not os.path.isdir(src_dir):
        pytest.skip("reference tree not present on this box")
    import re
    for obj, src in (("cuencas.o", "cuencas.f90"), ("modelosv2.o", "modelosv2.f90")):
        blob = open(os.path.join(ref.REF_OBJ_DIR, obj), "rb").read()
        lines = open(os.path.join(src_dir, src), encoding="latin-1").read().splitlines()
        hits = {int(m.group(1)) for m in re.finditer(rb"At line (\d+) of file wmf/" + src.encode(), blob)}
        assert hits
        for ln in hits:
            assert "allocate" in lines[ln - 1].lower(), (src, ln, lines[ln - 1])


def test_kat1_through_the_reference(ref):
    k = json.load(open(os.path.join(HERE, "golden", "kat1.json")))
    cu = ref.RefCu()
    cu.ncols, cu.nrows, cu.xll, cu.yll, cu.dx, cu.dy, cu.dxp, cu.nodata = (
        k["ncols"], k["nrows"], k["xll"], k["yll"], k["dx"], k["dx"], k["dxp"], k["nodata"])
    D = np.array(k["dir_rows_top_to_bottom"], np.int32).T
    assert cu.coord2fil_col(*k["outlet_xy"]) == tuple(k["outlet_colfil"])
    n = cu.basin_find(*k["outlet_xy"], D)
    assert n == k["nceldas"]
    b = cu.basin_cut(n)
    assert b.tolist() == k["basin_f"]
    assert cu.basin_acum(b).tolist() == k["acum"]


@pytest.mark.parametrize("kind,nc,nr,seed,dx,xll,yll", [
    ("G1", 40, 30, 0, 30.0, 0.0, 0.0), ("G2", 96, 80, 1, 12.0, 0.0, 0.0), ("G2", 200, 150, 2, 150.0, 0.0, 0.0),
    ("G1", 129, 65, 3, 1.0 / 3600.0, -75.6, 6.1),           # geographic fp32 coordinates (SURVEY 8d adversarial case)
])
def test_path_a_oracle_is_bit_identical_to_the_reference(ref, kind, nc, nr, seed, dx, xll, yll):
    from wmf_b200 import synth
    bs = make_basin(kind, nc, nr, seed, dx=dx, xll=xll, yll=yll)
    rc, oc = pair(ref, bs)
    assert rc.coord2fil_col(bs["x"], bs["y"]) == oc.coord2fil_col(bs["x"], bs["y"])
    n = rc.basin_find(bs["x"], bs["y"], bs["DIR"])
    assert n == oc.basin_find(bs["x"], bs["y"], bs["DIR"]) == bs["n"]
    b = rc.basin_cut(n)
    same(b, oc.basin_cut(n), "basin_cut")
    same(rc.basin_acum(b), oc.basin_acum(b), "basin_acum")
    assert rc.area == oc.area
    args = (bs["xll"], bs["yll"], bs["dx"], bs["dx"], synth.NODATA)
    dv = rc.basin_map2basin(b, bs["DIR"].astype(np.float32), *args)
    same(dv, oc.basin_map2basin(b, bs["DIR"].astype(np.float32), *args), "map2basin DIR")
    de = rc.basin_map2basin(b, bs["DEM"], *args)
    same(de, oc.basin_map2basin(b, bs["DEM"], *args), "map2basin DEM")
    hole = np.array(bs["DEM"], copy=True)
    hole[nc // 3: nc // 2, nr // 4: nr // 2] = synth.NODATA          # nodata inside the basin -> map mean (:1159-1180)
    same(rc.basin_map2basin(b, hole, *args), oc.basin_map2basin(b, hole, *args), "map2basin nodata")
    dvi = dv.astype(np.int32)
    r1, r2 = rc.basin_basics(b, de, dvi), oc.basin_basics(b, de, dvi)
    for a, c, nm in zip(r1, r2, ("acum", "long", "pend", "elev")):
        same(a, c, "basin_basics " + nm)
    for k in ("area", "pend_media", "elevacion", "centrox", "centroy"):
        assert getattr(rc, k) == getattr(oc, k), k
    acum, lng, pend, elev = r1
    for a, c, nm in zip(rc.basin_coordxy(b), oc.basin_coordxy(b), "XY"):
        same(a, c, "basin_coordXY " + nm)
    same(rc.basin_acum_var(b[0], lng), oc.basin_acum_var(b[0], lng), "basin_acum_var")
    for thr in (5, 30, 500):
        for a, c, nm in zip(rc.basin_stream_nod(b, acum, thr), oc.basin_stream_nod(b, acum, thr),
                            ("cauce", "nodos", "trazado", "n_nodos", "n_cauce")):
            same(a, c, f"basin_stream_nod({thr}) {nm}")
    same(rc.basin_stream_type(b, acum, [30, 500]), oc.basin_stream_type(b, acum, [30, 500]), "basin_stream_type")
    cauce = rc.basin_stream_nod(b, acum, 30)[0]
    for a, c, nm in zip(rc.geo_hand(b, elev, lng, cauce), oc.geo_hand(b, elev, lng, cauce), ("hand", "hdnd", "a_quien")):
        same(a, c, "geo_hand " + nm)
    speed = (0.1 + 0.01 * np.arange(n) % 3).astype(np.float32)
    same(rc.basin_time_to_out(b, lng, speed), oc.basin_time_to_out(b, lng, speed), "basin_time_to_out")
    same(rc.basin_propagate(b, elev), oc.basin_propagate(b, elev), "basin_propagate")
    mc, mr = rc.basin_2map_find(b)
    assert (mc, mr) == oc.basin_2map_find(b)
    m1, m2 = rc.basin_2map(b, elev, mc, mr), oc.basin_2map(b, elev, mc, mr)
    same(m1[0], m2[0], "basin_2map")
    assert m1[1:] == m2[1:]
    amap = rc.basin_float_var2map(b, acum.astype(np.float32), nc, nr)
    same(amap, oc.basin_float_var2map(b, acum.astype(np.float32), nc, nr), "basin_float_var2map")
    red = rc.geo_acum_to_cauce(amap.astype(np.int32), 30)
    same(red, oc.geo_acum_to_cauce(amap.astype(np.int32), 30), "geo_acum_to_cauce")
    same(rc.geo_hand_global(bs["DEM"], bs["DIR"], red), oc.geo_hand_global(bs["DEM"], bs["DIR"], red), "geo_hand_global")


def test_path_a_on_a_random_partial_basin(ref):
    """Irregular tree (pits, holes, up to 8 cells draining into one): basin_find's probe order and queue are the reference's."""
    import oracle
    from wmf_b200 import synth
    nc, nr, dx = 150, 110, 30.0
    DIR, (oc_, orow) = synth.make_partial_raster(nc, nr, 7, noise=1.5)
    x, y = dx * (oc_ - 0.5), dx * (nr - orow + 0.5)
    rc, oc = ref.RefCu(), oracle.Cu()
    for c in (rc, oc):
        c.ncols, c.nrows, c.xll, c.yll, c.dx, c.dy, c.dxp, c.nodata = nc, nr, 0.0, 0.0, dx, dx, dx, synth.NODATA
    n = rc.basin_find(x, y, DIR)
    assert n == oc.basin_find(x, y, DIR) and 1000 < n < (nc - 2) * (nr - 2)
    b = rc.basin_cut(n)
    same(b, oc.basin_cut(n), "basin_cut")
    same(rc.basin_acum(b), oc.basin_acum(b), "basin_acum")


CORE = ("q", "stoout", "hum", "st1", "st3", "balance", "speed", "areacontrol", "mean_rain", "acum_rain")
SED = ("qsed", "vs", "vd", "vsc", "vdc", "volero", "voldepo", "erot", "dept")
SLIDES = ("sl_zmin", "sl_zcrit", "sl_zmax", "sl_bo", "sl_riskvector", "sl_slideocurrence", "sl_slideacumulate",
          "sl_slidencelltime")
SHIA_CASES = {
    # name: (generator kwargs, module flags, extra outputs that must match)
    "linear": (dict(n_control_h=4), dict(show_storage=1, show_speed=1, show_mean_speed=1, show_area=1),
               ("mean_storage", "mean_speed")),
    "kinematic_returns": (dict(speed_type=(2, 2, 1), retorno_gr=1, retorno_aq=1, n_control_h=2),
                          dict(show_storage=1, show_mean_speed=1, show_mean_retorno=1, show_speed=1),
                          ("mean_storage", "mean_speed", "mean_retorno")),
    "all_kinematic": (dict(speed_type=(2, 2, 2)), dict(show_speed=1), ()),
    "sediments": (dict(speed_type=(2, 1, 1), sediments=True, rain_peak_mm=60.0), {}, SED),
    "sediments_slides": (dict(speed_type=(2, 2, 1), sediments=True, slides=True, retorno_gr=1, rain_peak_mm=60.0), {},
                         SED + SLIDES + ("retorned",)),
    "slides_gullies": (dict(slides=True, retorno_gr=1, rain_peak_mm=80.0), dict(sl_gullienogullie=1), SLIDES),
    "separate_fluxes": (dict(rain_peak_mm=50.0), dict(separate_fluxes=1), ("qseparated",)),
    "dump_switches": (dict(retorno_gr=1, rain_peak_mm=40.0), dict(save_vfluxes=1, save_rc=1, save_retorno=1),
                      ("mean_vfluxes", "rc_coef")),
    "separate_rain": (dict(rain_peak_mm=50.0, retorno_gr=1), dict(separate_rain=1),
                      ("qsep_byrain", "storage_conv", "storage_stra")),
    "separate_rain_kinematic": (dict(rain_peak_mm=50.0, speed_type=(2, 2, 1)), dict(separate_rain=1),
                                ("qsep_byrain", "storage_conv", "storage_stra")),
    # capillary tanks that start empty: (Storage_conv - Evp*Storage_conv/StoOut) is 0/0 and the compiled MAX(0., NaN) is 0
    "separate_rain_empty_capillary": (dict(rain_peak_mm=50.0), dict(separate_rain=1, _empty_capillary=1),
                                      ("qsep_byrain", "storage_conv", "storage_stra")),
}


@pytest.mark.parametrize("name", sorted(SHIA_CASES))
def test_shia_oracle_is_bit_identical_to_the_reference(ref, name, tmp_path):
    import oracle
    from wmf_b200 import synth
    kw, flags, extra = SHIA_CASES[name]
    seed = 5 + sorted(SHIA_CASES).index(name)
    bs = make_basin("G2" if seed % 2 else "G1", 70, 60, seed)
    v = basin_vectors(bs)
    T = 60
    inp = synth.make_shia_inputs(bs["basin"], v["acum"], v["lng"], v["pend"], dxp=30.0, n_reg=T, seed=seed, **kw)
    inp.update(flags)
    dumps = None
    if kw.get("retorno_gr"):
        inp["calib"][9] = 0.02                       # small gravitational capacity: tank 3 does spill back into tank 2
    if flags.get("separate_rain"):
        inp.update(synth.make_rain_types(inp, seed))
    if inp.pop("_empty_capillary", 0):
        st = np.array(inp["storage"], copy=True)
        st[0, ::7] = 0.0
        inp["storage"], inp["storage_constant"] = st, 0.0
    if name == "dump_switches":
        g = np.zeros(T, np.int32)
        g[[3, 20, T - 1]] = [1, 2, 3]
        inp["guarda_vfluxes"] = g
        dumps = {k: str(tmp_path / (k + ".bin")) for k in ("ruta_vfluxes", "ruta_retorno", "ruta_rc")}
    R = ref.shia_v1(inp, dumps=dumps)
    O = oracle.shia_v1(inp)
    assert np.abs(O["q"]).max() > 0 and np.isfinite(R["q"]).all()
    for k in CORE + tuple(extra):
        assert R.get(k) is not None, k
        same(R[k], O[k], f"{name}: {k}")
    for k in extra:
        assert np.abs(np.asarray(O[k], np.float64)).max() > 0, f"{name}: {k} is identically zero, the case does not exercise it"
    if dumps:
        N = bs["n"]
        same(np.fromfile(dumps["ruta_rc"], np.float32), np.asarray(O["rc_coef"]).reshape(-1, order="F"), "rc file")
        vf = np.fromfile(dumps["ruta_vfluxes"], np.float32).reshape(3, N, 4)
        same(vf[2].T, O["vfluxes_last"], "last vfluxes record")
        rf = np.fromfile(dumps["ruta_retorno"], np.float32).reshape(T, N)
        same(rf[T - 1], O["retorned_last"], "last retorno record")
        assert rf.max() > 0


def test_shia_continuation_from_stoin(ref):
    """StoIn (modelosv2.f90:271-277): a run continued from another run's final storages."""
    import oracle
    from wmf_b200 import synth
    bs = make_basin("G2", 64, 56, 21)
    v = basin_vectors(bs)
    inp = synth.make_shia_inputs(bs["basin"], v["acum"], v["lng"], v["pend"], dxp=30.0, n_reg=40, seed=21, speed_type=(2, 1, 1))
    first = oracle.shia_v1(inp)
    cont = dict(inp, stoin=first["stoout"])
    R, O = ref.shia_v1(cont), oracle.shia_v1(cont)
    for k in CORE:
        same(R[k], O[k], "continued run: " + k)
    assert not np.array_equal(O["q"], first["q"])


def test_hspeedin_leaves_vspeed_undefined_in_the_reference(ref):
    """With HspeedIn given the reference never sets vspeed (it is assigned only in the else branch, modelosv2.f90:285-299)
    and runs on whatever the heap holds.  The restatement and the CUDA path compute vspeed in both branches (DESIGN.md,
    deviations).  This loader hands the Fortran zeroed blocks, so here the reference must equal the restatement run with
    v_coef = 0 -- which shows the quirk is exactly that and nothing else differs on the HspeedIn path."""
    import oracle
    from wmf_b200 import synth
    bs = make_basin("G2", 64, 56, 22)
    v = basin_vectors(bs)
    inp = synth.make_shia_inputs(bs["basin"], v["acum"], v["lng"], v["pend"], dxp=30.0, n_reg=40, seed=22, speed_type=(2, 1, 1))
    first = oracle.shia_v1(inp)
    cont = dict(inp, stoin=first["stoout"], hspeedin=first["hspeed_out"])
    R = ref.shia_v1(cont)
    O0 = oracle.shia_v1(dict(cont, v_coef=np.zeros_like(np.asarray(inp["v_coef"]))))
    for k in CORE:
        same(R[k], O0[k], "HspeedIn run with vspeed = 0: " + k)
    assert not np.array_equal(oracle.shia_v1(cont)["q"], O0["q"])


def test_calc_speed(ref):
    import oracle
    rng = np.random.default_rng(3)
    for _ in range(200):
        sm, coef, expo, L, v0 = (float(np.float32(x)) for x in (rng.uniform(0, 5e4), rng.uniform(0.05, 3), rng.uniform(0.2, 2.0),
                                                               rng.uniform(10, 200), rng.uniform(0.01, 2)))
        a_ref, v_ref = ref.calc_speed(sm, coef, expo, L, v0, 300.0, 5)
        a_or, v_or = oracle.calc_speed(sm, coef, expo, L, v0, 300.0, 5)
        assert (np.float32(a_ref), np.float32(v_ref)) == (np.float32(a_or), np.float32(v_or))


def test_rain_interpolation_oracle_is_bit_identical_to_the_reference(ref, tmp_path):
    """rain_idw (modelosv2.f90:1195-1271), rain_mit (:1104-1194), rain_pre_mit (:1068-1101): record files, posIds and the
    sequential fp32 meanRain, with missing readings (-999), an all-missing step (0/0) and a drizzle threshold."""
    import oracle
    from scipy.spatial import Delaunay
    from wmf_b200 import synth
    bs = make_basin("G2", 70, 60, 9)
    x, y = bs["cu"].basin_coordxy(bs["basin"])
    xy = np.asfortranarray(np.vstack([x, y]), dtype=np.float32)
    coord, rain = synth.make_stations(70, 60, bs["dx"], 9, 40, 4)
    for pp, umbral in ((1.0, 0.0), (2.0, 0.0), (1.7, 0.0), (2.0, 2.5)):
        R, O = ref.rain_idw(xy, coord, rain, pp, umbral), oracle.rain_idw(xy, coord, rain, pp, umbral)
        for k in ("table", "pos_ids", "mean_rain"):
            same(R[k], O[k], f"rain_idw(pp={pp}, umbral={umbral}) {k}")
        assert O["table"].shape[0] > 5 and (O["pos_ids"] == 1).any()
    # the hills model of rain_idw: records of nhills sums
    v = basin_vectors(bs)
    cauce, nodos, nn = bs["cu"].basin_subbasin_nod(bs["basin"], v["acum"], 20)
    hills_own, _ = bs["cu"].basin_subbasin_find(bs["basin"], nodos, nn)
    R, O = ref.rain_idw(xy, coord, rain, 2.0, 0.0, mask=hills_own, nhills=nn), oracle.rain_idw(xy, coord, rain, 2.0, 0.0, mask=hills_own, nhills=nn)
    for k in ("table", "pos_ids", "mean_rain"):
        same(R[k], O[k], "rain_idw by hills " + k)
    assert O["table"].shape[1] == nn
    tin = np.asfortranarray((Delaunay(coord.T.astype(np.float64)).simplices.T + 1).astype(np.int32))
    perte = ref.rain_pre_mit(xy, tin, coord)
    assert perte.min() >= 1
    same(perte, oracle.rain_pre_mit(xy, tin, coord), "rain_pre_mit")
    R, O = ref.rain_mit(xy, coord, rain, tin, perte, 0.0), oracle.rain_mit(xy, coord, rain, tin, perte, 0.0)
    for k in ("table", "pos_ids", "mean_rain"):
        same(R[k], O[k], "rain_mit " + k)


@pytest.mark.parametrize("kind,nc,nr,seed,thr", [("G2", 96, 80, 1, 30), ("G1", 70, 60, 2, 12), ("G2", 200, 150, 3, 60),
                                                 ("G2", 120, 90, 4, 5)])
def test_subbasin_topology_oracle_is_bit_identical_to_the_reference(ref, kind, nc, nr, seed, thr):
    """The hills / sub-basin routines SimuBasin.__init__ and set_Geomorphology call even for the cells model
    (wmf.py:3019-3023, 3664-3676): basin_subbasin_nod / find / cut / horton / long / stream_prop / map2subbasin
    (cuencas.f90:1835-2118) and basin_stream_slope (:1380-1438), with their quirks -- the INTEGER `long` dummy, the
    implicitly INTEGER L_tramo, celdas_tramo(2) left over from an earlier reach, at most 10 tributaries in the Horton pass."""
    bs = make_basin(kind, nc, nr, seed)
    b, n = bs["basin"], bs["n"]
    v = basin_vectors(bs)
    rc, oc = pair(ref, bs)
    acum, lng, pend, elev = v["acum"], v["lng"], v["pend"], v["elev"]
    r1, o1 = rc.basin_subbasin_nod(b, acum, thr), oc.basin_subbasin_nod(b, acum, thr)
    assert r1[2] == o1[2] and r1[2] > 10
    nn = r1[2]
    same(r1[0], o1[0], "cauce")
    same(r1[1], o1[1], "nodos_fin")
    rf, of = rc.basin_subbasin_find(b, r1[1], nn), oc.basin_subbasin_find(b, o1[1], nn)
    same(rf[0], of[0], "sub_pert")
    rsb, osb = rc.basin_subbasin_cut(nn), oc.basin_subbasin_cut(nn)
    same(rsb, osb, "sub_basins")
    rh, oh = rc.basin_subbasin_horton(rsb, n), oc.basin_subbasin_horton(osb, n)
    same(rh[0], oh[0], "sub_horton")
    same(rh[1][rsb[0] - 1], oh[1][osb[0] - 1], "nod_horton at the node cells")
    assert oh[0].max() >= 3
    rl, ol = rc.basin_subbasin_long(rf[0], r1[0], lng, rsb, rh[0]), oc.basin_subbasin_long(of[0], o1[0], lng, osb, oh[0])
    same(rl[0], ol[0], "sub_basin_long")
    assert rl[1:] == ol[1:] and ol[1] > 0
    rp, op = rc.basin_subbasin_stream_prop(rf[0], r1[0], lng, pend, nn), oc.basin_subbasin_stream_prop(of[0], o1[0], lng, pend, nn)
    for a, c, nm in zip(rp, op, ("stream_slope", "stream_long")):
        np.testing.assert_array_equal(a, c, err_msg=nm)             # NaN == NaN for hills without channel cells
    for sm in (0, 1, 2):
        same(rc.basin_subbasin_map2subbasin(rf[0], elev, nn, r1[0], sm), oc.basin_subbasin_map2subbasin(of[0], elev, nn, o1[0], sm),
             f"map2subbasin({sm})")
    cauce, nodos, _, _, n_cauce = oc.basin_stream_nod(b, acum, thr)
    rs, os_ = rc.basin_stream_slope(b, elev, lng, nodos, n_cauce), oc.basin_stream_slope(b, elev, lng, nodos, n_cauce)
    same(rs[0], os_[0], "stream_s")
    same(rs[1], os_[1], "stream_l")
    assert (os_[1][cauce == 1] == np.floor(os_[1][cauce == 1])).all()         # the implicitly INTEGER reach length


def test_stream_find_to_corr_oracle_is_bit_identical_to_the_reference(ref):
    import oracle
    bs = make_basin("G2", 120, 90, 4)
    rc, oc = pair(ref, bs)
    v = basin_vectors(bs)
    amap = oc.basin_float_var2map(bs["basin"], v["acum"].astype(np.float32), 120, 90)
    cauce = (amap > 40).astype(np.int32)
    rng = np.random.default_rng(0)
    for _ in range(100):
        c, f = int(rng.integers(2, 119)), int(rng.integers(2, 89))
        x, y = 30.0 * (c - 0.5), 30.0 * (90 - f + 0.5)
        assert rc.stream_find_to_corr(x, y, bs["DEM"], bs["DIR"], cauce) == oc.stream_find_to_corr(x, y, bs["DEM"], bs["DIR"], cauce)


@pytest.mark.parametrize("seed,kw", [(12, dict(rain_peak_mm=60.0)), (13, dict(rain_peak_mm=90.0, speed_type=(2, 1, 1)))])
def test_hydroflash_oracle_is_bit_identical_to_the_reference(ref, seed, kw):
    """sim_floods = 1 (modelosv2.f90:728-743, flood_params :1730-1749, flood_debris_flow2 :1786-1822): the flooded-cell map,
    the hydraulic maps of the evaluated channel cells, the flooded cross sections and flood_rdf / flood_diff of the last
    evaluation (flood_area is never assigned by the Fortran)."""
    import oracle
    from wmf_b200 import synth
    bs = make_basin("G2", 90, 70, seed)
    v = basin_vectors(bs)
    inp = synth.make_shia_inputs(bs["basin"], v["acum"], v["lng"], v["pend"], dxp=30.0, n_reg=60, seed=seed, thresholds=(20, 150), **kw)
    inp.update(synth.make_flood_inputs(inp, v["elev"], v["pend"], 4, seed))
    R, O = ref.shia_v1(inp), oracle.shia_v1(inp)
    for k in CORE + ("flood_flood", "flood_h", "flood_ufr", "flood_cr", "flood_q", "flood_qsed", "flood_speed", "flood_profundidad",
                     "flood_scalars"):
        same(R[k], O[k], "floods: " + k)
    counts = np.bincount(O["flood_flood"], minlength=3)
    assert counts[1] > 100 and counts[2] > 100 and np.isfinite(O["flood_profundidad"]).all()


@pytest.mark.parametrize("case", range(16))
def test_fuzz_shia_against_the_reference(ref, case, tmp_path):
    """Random partial basins (irregular trees, up to 8 cells draining into one) x random switches: the restatement must stay
    bit-identical to the reference's Fortran on whatever combination comes up (conftest.fuzz_case)."""
    import oracle
    from conftest import fuzz_case
    fc = fuzz_case(case)
    if fc is None:
        pytest.skip("the random outlet caught a tiny basin")
    inp, mode, extra, flags = fc["inp"], fc["mode"], fc["extra"], fc["flags"]
    rc = ref.RefCu()
    rc.ncols, rc.nrows, rc.xll, rc.yll, rc.dx, rc.dy, rc.dxp, rc.nodata = fc["geo"]
    assert rc.basin_find(fc["x"], fc["y"], fc["DIR"]) == fc["n"]
    same(rc.basin_cut(fc["n"]), fc["basin"], "basin_cut")
    for a, c, nm in zip(rc.basin_basics(fc["basin"], fc["demv"], fc["dirv"]), fc["basics"], ("acum", "long", "pend", "elev")):
        same(a, c, "basin_basics " + nm)
    dumps = {k: str(tmp_path / (k + ".bin")) for k in ("ruta_vfluxes", "ruta_rc")} if mode == "dumps" else None
    R, O = ref.shia_v1(inp, dumps=dumps), oracle.shia_v1(inp)
    for k in CORE + tuple(extra):
        np.testing.assert_array_equal(np.asarray(R[k]).reshape(-1), np.asarray(O[k]).reshape(-1), err_msg=f"case {case} ({mode}): {k}")
    for k, fl in (("mean_storage", "show_storage"), ("mean_speed", "show_mean_speed"), ("mean_retorno", "show_mean_retorno")):
        if flags.get(fl) and R.get(k) is not None:
            np.testing.assert_array_equal(np.asarray(R[k]).reshape(-1), np.asarray(O[k]).reshape(-1), err_msg=f"case {case}: {k}")


def test_control_point_placement_oracle_is_bit_identical_to_the_reference(ref):
    """basin_point2var (cuencas.f90:1230-1261) and basin_stream_point2stream (:1524-1577): points inside / outside the basin,
    on and off the stream network."""
    bs = make_basin("G2", 120, 90, 4)
    rc, oc = pair(ref, bs)
    v = basin_vectors(bs)
    cauce = (v["acum"] >= 40).astype(np.int32)
    rng = np.random.default_rng(1)
    xy = np.asfortranarray(np.vstack([rng.uniform(-100, 3700, 60), rng.uniform(-100, 2800, 60)]), dtype=np.float32)
    ids = np.arange(101, 161, dtype=np.int32)
    for a, c, nm in zip(rc.basin_stream_point2stream(bs["basin"], cauce, ids, xy), oc.basin_stream_point2stream(bs["basin"], cauce, ids, xy),
                        ("res_coord", "basin_pts", "xy_new")):
        same(a, c, "basin_stream_point2stream " + nm)
    for a, c, nm in zip(rc.basin_point2var(bs["basin"], ids, xy), oc.basin_point2var(bs["basin"], ids, xy), ("res_coord", "basin_pts")):
        same(a, c, "basin_point2var " + nm)


def test_stream_sections_oracle_is_bit_identical_to_the_reference(ref):
    """basin_stream_sections (cuencas.f90:1464-1523), the cross sections HydroFlash reads."""
    bs = make_basin("G2", 90, 70, 6)
    rc, oc = pair(ref, bs)
    v = basin_vectors(bs)
    cauce = (v["acum"] >= 30).astype(np.int32)
    for num in (2, 4):
        r, o = rc.basin_stream_sections(bs["basin"], cauce, v["dirv"], bs["DEM"], num), oc.basin_stream_sections(bs["basin"], cauce, v["dirv"], bs["DEM"], num)
        same(r[0], o[0], "secciones")
        same(r[1], o[1], "secciones_cel")
        assert (o[1] == -9999).any() and (o[1] > 0).any()
