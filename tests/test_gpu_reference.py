"""GPU parity against THE REFERENCE ITSELF: the CUDA path through the C ABI vs the reference's own compiled Fortran
(oracle/_ref/libwmf_ref.so, built in the container from the objects in the reference tree; the .so travels to the GPU
box).  Path A must be bit-exact, Path B within the north-star tolerance (1e-5 relative, fp32)."""
import numpy as np
import pytest

from conftest import basin_vectors, geo_to, make_basin

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    from oracle import ref as r
    if not r.available() and not r.build():
        pytest.skip("oracle/_ref/libwmf_ref.so did not travel to this box")
    r.lib()
    return r


def close(got, want, name, rtol=1e-5, afloor=1e-7):
    got, want = np.asarray(got, np.float64).reshape(-1), np.asarray(want, np.float64).reshape(-1)
    assert got.shape == want.shape, (name, got.shape, want.shape)
    scale = max(float(np.abs(want).max()), 1e-30)
    bad = np.abs(got - want) > rtol * np.abs(want) + afloor * scale
    assert not bad.any(), f"{name}: {bad.sum()} of {bad.size} values off"


@pytest.mark.parametrize("kind,nc,nr,seed", [("G2", 300, 220, 4), ("G1", 257, 129, 5)])
def test_path_a_bit_exact_against_the_reference(ref, kind, nc, nr, seed):
    from wmf_b200 import CuModule, synth
    bs = make_basin(kind, nc, nr, seed)
    rc, cu = ref.RefCu(), CuModule()
    geo_to(rc, bs)
    geo_to(cu, bs)
    n = cu.basin_find(bs["x"], bs["y"], bs["DIR"], nc, nr)
    assert n == rc.basin_find(bs["x"], bs["y"], bs["DIR"])
    b = cu.basin_cut(n)
    np.testing.assert_array_equal(b, rc.basin_cut(n))
    np.testing.assert_array_equal(np.asarray(cu.basin_acum(b)).reshape(-1), rc.basin_acum(b))
    args = (bs["xll"], bs["yll"], bs["dx"], bs["dx"], synth.NODATA)
    dv = rc.basin_map2basin(b, bs["DIR"].astype(np.float32), *args).astype(np.int32)
    de = rc.basin_map2basin(b, bs["DEM"], *args)
    np.testing.assert_array_equal(np.asarray(cu.basin_map2basin(b, bs["DEM"], *args)).reshape(-1), de)
    got, want = cu.basin_basics(b, de, dv), rc.basin_basics(b, de, dv)
    for g, w, nm in zip(got, want, ("acum", "long", "pend", "elev")):
        np.testing.assert_array_equal(np.asarray(g).reshape(-1), w, err_msg="basin_basics " + nm)
    acum, lng, pend, elev = want
    cauce = rc.basin_stream_nod(b, acum, 30)[0]
    for g, w, nm in zip(cu.geo_hand(b, elev, lng, cauce), rc.geo_hand(b, elev, lng, cauce), ("hand", "hdnd", "a_quien")):
        np.testing.assert_array_equal(np.asarray(g).reshape(-1), w, err_msg="geo_hand " + nm)
    np.testing.assert_array_equal(np.asarray(cu.basin_stream_type(b[0], acum, np.array([30, 500], np.int32))).reshape(-1),
                                  rc.basin_stream_type(b, acum, [30, 500]))


@pytest.mark.parametrize("case", ["linear", "kinematic_returns", "sed_slides", "separate_fluxes"])
def test_shia_against_the_reference(ref, case):
    from wmf_b200 import ModelsModule, synth
    kw, flags = {
        "linear": (dict(n_control_h=3), dict(show_speed=1)),
        "kinematic_returns": (dict(speed_type=(2, 2, 1), retorno_gr=1, retorno_aq=1), {}),
        "sed_slides": (dict(speed_type=(2, 2, 1), sediments=True, slides=True, retorno_gr=1, rain_peak_mm=60.0), {}),
        "separate_fluxes": (dict(rain_peak_mm=50.0), dict(separate_fluxes=1)),
    }[case]
    bs = make_basin("G2", 120, 90, 31)
    v = basin_vectors(bs)
    inp = synth.make_shia_inputs(bs["basin"], v["acum"], v["lng"], v["pend"], dxp=30.0, n_reg=80, seed=31, **kw)
    R = ref.shia_v1(dict(inp, **flags))
    m = ModelsModule()
    synth.apply_to_models(m, inp)
    for k, val in flags.items():
        setattr(m, k, val)
    m.set_rain(inp["rain"], inp["pos_evento"])
    n_cont = int(np.count_nonzero(inp["control"])) + 1
    n_conth = max(1, int(np.count_nonzero(inp["control_h"])))
    ret = m.shia_v1(None, None, inp["calib"], n_cont, n_conth, inp["n_reg"], None, None, bs["n"])
    close(ret[0], R["q"], "Q")
    close(ret[9], R["stoout"], "StoOut")
    close(ret[3], R["hum"], "Hum")
    close(ret[7], R["speed"], "Speed")
    close(m.mean_rain, R["mean_rain"], "Mean_Rain")
    np.testing.assert_array_equal(np.asarray(m.acum_rain).reshape(-1), R["acum_rain"])
    if case == "separate_fluxes":
        close(ret[2], R["qseparated"], "Qseparated")
    if case == "sed_slides":
        close(ret[1], R["qsed"], "Qsed")
        close(m.vd, R["vd"], "Vd")
        np.testing.assert_array_equal(np.asarray(m.sl_riskvector).reshape(-1), R["sl_riskvector"])
        np.testing.assert_array_equal(np.asarray(m.sl_slideocurrence).reshape(-1), R["sl_slideocurrence"])


def test_dump_files_against_the_reference(ref, tmp_path):
    """save_storage / save_speed / save_vfluxes / save_retorno / save_rc: the direct-access files the reference's own Fortran
    writes (its write_float_basin calls, modelosv2.f90:799-823, 863-867) against the files the drop-in writes: same sizes, same
    record numbering, values within the north-star tolerance."""
    import os
    from wmf_b200 import ModelsModule, synth
    bs = make_basin("G2", 100, 80, 35)
    v = basin_vectors(bs)
    T = 50
    inp = synth.make_shia_inputs(bs["basin"], v["acum"], v["lng"], v["pend"], dxp=30.0, n_reg=T, seed=35, speed_type=(2, 1, 1),
                                 retorno_gr=1, rain_peak_mm=40.0)
    inp["calib"][9] = 0.02
    guarda = np.zeros(T, np.int32)
    guarda[[5, 21, 49]] = [1, 2, 3]
    flags = dict(save_storage=1, save_speed=1, save_vfluxes=1, save_retorno=1, save_rc=1)
    names = ("ruta_storage", "ruta_speed", "ruta_vfluxes", "ruta_retorno", "ruta_rc")
    rdir, gdir = tmp_path / "ref", tmp_path / "gpu"
    os.makedirs(rdir), os.makedirs(gdir)
    rpaths = {k: str(rdir / (k + ".bin")) for k in names}
    gpaths = {k: str(gdir / (k + ".bin")) for k in names}
    ref.shia_v1(dict(inp, guarda_cond=guarda, guarda_vfluxes=guarda, **flags), dumps=rpaths)
    m = ModelsModule()
    synth.apply_to_models(m, inp)
    for k, val in flags.items():
        setattr(m, k, val)
    m.guarda_cond, m.guarda_vfluxes = guarda, guarda
    m.set_rain(inp["rain"], inp["pos_evento"])
    n_cont = int(np.count_nonzero(inp["control"])) + 1
    m.shia_v1(None, None, inp["calib"], n_cont, 1, T, None, None, bs["n"], gpaths["ruta_storage"], gpaths["ruta_speed"],
              gpaths["ruta_vfluxes"], None, None, None, None, gpaths["ruta_retorno"], gpaths["ruta_rc"])
    for k in names:
        assert os.path.getsize(gpaths[k]) == os.path.getsize(rpaths[k]), k
        g, r = np.fromfile(gpaths[k], np.float32), np.fromfile(rpaths[k], np.float32)
        assert np.abs(r).max() > 0, k
        close(g, r, k)
