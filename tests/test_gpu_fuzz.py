"""GPU parity on random partial basins x random switch combinations (conftest.fuzz_case): the cases on which the oracle is
checked bit for bit against the reference's Fortran (tests/test_reference_pin.py::test_fuzz_shia_against_the_reference) are
run through the CUDA path and compared with the oracle at the north-star tolerance."""
import numpy as np
import pytest

from conftest import fuzz_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", range(16))
def test_fuzz_shia_gpu_vs_oracle(case, tmp_path):
    import oracle
    from test_gpu_shia import close, compare_core
    from wmf_b200 import ModelsModule, synth
    fc = fuzz_case(case)
    if fc is None:
        pytest.skip("the random outlet caught a tiny basin")
    inp, mode, extra, n, T = dict(fc["inp"]), fc["mode"], fc["extra"], fc["n"], fc["T"]
    # (separate_* together with the per-step means: two launch sequences behind one shia_v1 call, _models._switch_passes)
    ref = oracle.shia_v1(inp)
    m = ModelsModule()
    synth.apply_to_models(m, inp)
    m.set_rain(inp["rain"], inp["pos_evento"])
    if mode == "sepr":
        m.set_rain_types(inp["conv"], inp["stra"], inp["pos_conv"], inp["pos_stra"])
    n_cont = int(np.count_nonzero(inp["control"])) + 1
    n_conth = max(1, int(np.count_nonzero(inp["control_h"])))
    paths = [None] * 9
    if mode == "dumps":
        paths[2], paths[8] = str(tmp_path / "vf.bin"), str(tmp_path / "rc.bin")        # ruta_vfluxes, ruta_rc
    ret = m.shia_v1(None, None, inp["calib"], n_cont, n_conth, T, None, None, n, *paths)
    compare_core(ret, m, ref, n)
    if inp.get("show_speed"):
        close(ret[7], ref["speed"], name="Speed")
    if inp.get("show_storage"):
        close(m.mean_storage, ref["mean_storage64"], name="mean_storage")
    if inp.get("show_mean_speed"):
        close(m.mean_speed, ref["mean_speed64"], name="mean_speed")
    if mode == "sepf":
        close(ret[2], ref["qseparated"], name="Qseparated")
    if mode == "sepr":
        close(ret[10], ref["qsep_byrain"], name="Qsep_byrain")
        close(m.storage_conv, ref["storage_conv"], name="Storage_conv")
    if mode == "dumps":
        close(m.mean_vfluxes, ref["mean_vfluxes64"], name="mean_vfluxes")
        close(m.rc_coef, ref["rc_coef"], name="rc_coef")
    if mode in ("sed", "sed_slides"):
        close(ret[1], ref["qsed"], name="Qsed")
        for k in ("vs", "vd", "vsc", "vdc"):
            close(getattr(m, k), ref[k], name=k)
        close(m.voldepo, ref["voldepo"], name="VolDEPo")
    if mode in ("slides", "sed_slides"):
        np.testing.assert_array_equal(np.asarray(m.sl_riskvector).reshape(-1), ref["sl_riskvector"])
        np.testing.assert_array_equal(np.asarray(m.sl_slideocurrence).reshape(-1), ref["sl_slideocurrence"])
        np.testing.assert_array_equal(np.asarray(m.sl_slidencelltime).reshape(-1), ref["sl_slidencelltime"])


@pytest.mark.parametrize("case", range(12))
def test_fuzz_topology_gpu_vs_oracle(case):
    """Path A and the sub-basin chain on the random partial basins: bit-exact against the oracle."""
    import oracle
    from wmf_b200 import CuModule, synth
    fc = fuzz_case(case)
    if fc is None:
        pytest.skip("the random outlet caught a tiny basin")
    nc, nr = fc["geo"][0], fc["geo"][1]
    oc, cu = oracle.Cu(), CuModule()
    for c in (oc, cu):
        c.ncols, c.nrows, c.xll, c.yll, c.dx, c.dy, c.dxp, c.nodata = fc["geo"]
    n = cu.basin_find(fc["x"], fc["y"], fc["DIR"], nc, nr)
    assert n == fc["n"]
    b = cu.basin_cut(n)
    np.testing.assert_array_equal(b, fc["basin"])
    got = cu.basin_basics(b, fc["demv"], fc["dirv"])
    for g, w, nm in zip(got, fc["basics"], ("acum", "long", "pend", "elev")):
        np.testing.assert_array_equal(np.asarray(g).reshape(-1), w, err_msg=nm)
    acum, lng, pend, elev = fc["basics"]
    thr = 4 + case
    o1, g1 = oc.basin_subbasin_nod(b, acum, thr), cu.basin_subbasin_nod(b, acum, thr, n)
    assert g1[2] == o1[2]
    np.testing.assert_array_equal(g1[1], o1[1])
    nn = o1[2]
    of, gf = oc.basin_subbasin_find(b, o1[1], nn), cu.basin_subbasin_find(b, g1[1], nn, n)
    np.testing.assert_array_equal(gf[0], of[0])
    osb, gsb = oc.basin_subbasin_cut(nn), cu.basin_subbasin_cut(nn)
    np.testing.assert_array_equal(gsb, osb)
    oh, gh = oc.basin_subbasin_horton(osb, n), cu.basin_subbasin_horton(gsb, n, nn)
    np.testing.assert_array_equal(gh[0], oh[0])
    ol, gl = oc.basin_subbasin_long(of[0], o1[0], lng, osb, oh[0]), cu.basin_subbasin_long(gf[0], g1[0], lng, gsb, gh[0], nn, n)
    np.testing.assert_array_equal(gl[0], ol[0])
    assert gl[1:] == ol[1:]
    op, gp = oc.basin_subbasin_stream_prop(of[0], o1[0], lng, pend, nn), cu.basin_subbasin_stream_prop(gf[0], g1[0], lng, pend, nn, n)
    np.testing.assert_array_equal(gp[0], op[0])
    np.testing.assert_array_equal(gp[1], op[1])
    cauce, nodos, _, _, ncau = oc.basin_stream_nod(b, acum, thr)
    os_, gs = oc.basin_stream_slope(b, elev, lng, nodos, ncau), cu.basin_stream_slope(b, elev, lng, nodos, ncau, n)
    np.testing.assert_array_equal(gs[0], os_[0])
    np.testing.assert_array_equal(gs[1], os_[1])
    for g, w, nm in zip(cu.geo_hand(b, elev, lng, cauce), oc.geo_hand(b, elev, lng, cauce), ("hand", "hdnd", "a_quien")):
        np.testing.assert_array_equal(np.asarray(g).reshape(-1), w, err_msg="geo_hand " + nm)
