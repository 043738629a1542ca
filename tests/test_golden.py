"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py FROM THE REFERENCE'S OWN COMPILED
FORTRAN, oracle/_ref; see that script).

CPU leg: the C restatement in oracle/ reproduces the reference's vectors (topology bit-exact; floats are bit-identical
on the day of generation and held to 1e-6 here to survive a libm update).
GPU leg: the CUDA path through the C ABI matches the same vectors at the north-star tolerance without touching oracle/.
"""
import importlib.util
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
mg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mg)

NAMES = sorted(mg.CASES)


def load(name):
    return dict(np.load(os.path.join(HERE, "golden", name + ".npz")))


def close(got, ref, rtol, name):
    got, ref = np.asarray(got, np.float64).reshape(-1), np.asarray(ref, np.float64).reshape(-1)
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    scale = max(float(np.abs(ref).max()), 1e-30)
    bad = np.abs(got - ref) > rtol * np.abs(ref) + 1e-7 * scale
    assert not bad.any(), f"{name}: {bad.sum()} of {bad.size} values off"


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_golden(name):
    g = load(name)
    bs, v, inp, flags = mg.build_case(name)
    np.testing.assert_array_equal(bs["basin"], g["basin_f"])
    np.testing.assert_array_equal(v["acum"], g["acum"])
    np.testing.assert_array_equal(v["lng"], g["long"])
    np.testing.assert_array_equal(v["pend"], g["pend"])
    assert int(inp["rain"].astype(np.int64).sum()) == int(g["rain_checksum"][0])
    ref = mg.run_oracle(inp, flags)
    for k in mg.KEEP:
        if "o_" + k in g:
            close(ref[k], g["o_" + k], 1e-6, f"{name}:{k}")


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_matches_golden(name):
    from conftest import geo_to
    from wmf_b200 import CuModule, ModelsModule, synth
    g = load(name)
    bs, v, inp, flags = mg.build_case(name)          # inputs only (generators); no oracle model run below
    cu = CuModule()
    geo_to(cu, bs)
    n = cu.basin_find(bs["x"], bs["y"], bs["DIR"], bs["nc"], bs["nr"])
    b = cu.basin_cut(n)
    np.testing.assert_array_equal(b, g["basin_f"])
    np.testing.assert_array_equal(cu.basin_acum(b), g["acum"])
    m = ModelsModule()
    synth.apply_to_models(m, inp)
    for k, val in flags.items():
        setattr(m, k, val)
    m.set_rain(inp["rain"], inp["pos_evento"])
    n_cont = int(np.count_nonzero(inp["control"])) + 1
    n_conth = max(1, int(np.count_nonzero(inp["control_h"])))
    ret = m.shia_v1(None, None, inp["calib"], n_cont, n_conth, inp["n_reg"], None, None, n)
    close(ret[0], g["o_q"], 1e-5, "Q")
    close(ret[9], g["o_stoout"], 1e-5, "StoOut")
    close(ret[3], g["o_hum"], 1e-5, "Hum")
    tot = float(np.abs(g["o_stoout"]).sum())
    assert np.abs(ret[6] - g["o_balance64"]).max() <= 2e-7 * tot
    np.testing.assert_array_equal(m.acum_rain.reshape(-1), g["o_acum_rain"])
    if flags.get("separate_fluxes"):
        assert g["o_qseparated"].max() > 0
        close(ret[2], g["o_qseparated"], 1e-5, "Qseparated")
    if "o_vs" in g:
        close(ret[1], g["o_qsed"], 1e-5, "Qsed")
        close(m.vd, g["o_vd"], 1e-5, "Vd")
        np.testing.assert_array_equal(m.sl_riskvector.reshape(-1), g["o_sl_riskvector"])


TOPO_KEYS = ("t_cauce", "t_nodos_fin", "t_sub_pert", "t_sub_basins", "t_sub_horton", "t_nod_horton_nodes", "t_sub_basin_long",
             "t_max_long", "t_hill_slope", "t_hill_long", "t_stream_s", "t_stream_l", "r_table", "r_pos", "r_mean", "r_table_hills")


def test_oracle_reproduces_topo_rain_golden():
    """Sub-basin chain + rain_idw (cells and hills): the reference's vectors against today's restatement."""
    import oracle
    g = load("topo_rain")
    bs, v, xy, coord, rain = mg.build_topo_case()
    O = mg.run_topo(bs["cu"], oracle, bs, v, xy, coord, rain, int(g["thr"][0]))
    for k in TOPO_KEYS:
        np.testing.assert_array_equal(np.asarray(O[k]), g[k], err_msg=k)


@pytest.mark.gpu
def test_cuda_matches_topo_rain_golden():
    from conftest import geo_to
    from wmf_b200 import CuModule, ModelsModule
    g = load("topo_rain")
    bs, v, xy, coord, rain = mg.build_topo_case()
    b, n, thr = bs["basin"], bs["n"], int(g["thr"][0])
    cu, m = CuModule(), ModelsModule()
    geo_to(cu, bs)
    cauce, nodos, nn = cu.basin_subbasin_nod(b, v["acum"], thr, n)
    np.testing.assert_array_equal(nodos, g["t_nodos_fin"])
    sub_pert, _ = cu.basin_subbasin_find(b, nodos, nn, n)
    np.testing.assert_array_equal(sub_pert, g["t_sub_pert"])
    sb = cu.basin_subbasin_cut(nn)
    np.testing.assert_array_equal(sb, g["t_sub_basins"])
    sh, nh = cu.basin_subbasin_horton(sb, n, nn)
    np.testing.assert_array_equal(sh, g["t_sub_horton"])
    sl, mx, nmx = cu.basin_subbasin_long(sub_pert, cauce, v["lng"], sb, sh, nn, n)
    np.testing.assert_array_equal(sl, g["t_sub_basin_long"])
    assert [mx, nmx] == g["t_max_long"].tolist()
    ps, pl = cu.basin_subbasin_stream_prop(sub_pert, cauce, v["lng"], v["pend"], nn, n)
    np.testing.assert_array_equal(ps, g["t_hill_slope"])
    np.testing.assert_array_equal(pl, g["t_hill_long"])
    c2, nod2, _, _, ncau = cu.basin_stream_nod(b, v["acum"], thr, n)
    ss, sll = cu.basin_stream_slope(b, v["elev"], v["lng"], nod2, ncau, n)
    np.testing.assert_array_equal(ss, g["t_stream_s"])
    np.testing.assert_array_equal(sll, g["t_stream_l"])
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        p1, p2 = os.path.join(d, "c.bin"), os.path.join(d, "h.bin")
        mean, pos = m.rain_idw(xy, coord, rain, 2.0, 0, p1, 0.0, np.ones(n))
        m.rain_idw(xy, coord, rain, 2.0, nn, p2, 0.0, sub_pert)
        np.testing.assert_array_equal(np.fromfile(p1, np.int32).reshape(-1, n), g["r_table"])
        np.testing.assert_array_equal(np.fromfile(p2, np.int32).reshape(-1, nn), g["r_table_hills"])
    np.testing.assert_array_equal(pos, g["r_pos"])
    # the reference's meanRain is a sequential fp32 sum over the cells (error up to ~N*eps/2); the GPU sums in fp64
    np.testing.assert_allclose(mean, g["r_mean"], rtol=n * 6e-8, atol=1e-9)
