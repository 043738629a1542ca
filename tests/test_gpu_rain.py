"""GPU parity, rainfall interpolation: rain_idw (modelosv2.f90:1195-1271) and rain_mit (:1104-1194) through the C ABI vs
the CPU oracle -- which is bit-identical to the reference's compiled Fortran on these routines
(tests/test_reference_pin.py::test_rain_interpolation_oracle_is_bit_identical_to_the_reference).  The record table is
integer (mm*1000 truncated): it must be bit-exact; meanRain is a sum over the cells: 1e-6 relative against the fp64 twin."""
import os

import numpy as np
import pytest

from conftest import make_basin

pytestmark = pytest.mark.gpu


def setup_case(nc=120, nr=90, seed=3, ns=11, T=50):
    from wmf_b200 import synth
    bs = make_basin("G2", nc, nr, seed)
    x, y = bs["cu"].basin_coordxy(bs["basin"])
    xy = np.asfortranarray(np.vstack([x, y]), dtype=np.float32)
    coord, rain = synth.make_stations(nc, nr, bs["dx"], ns, T, seed)
    return bs, xy, coord, rain


def check(got_mean, got_pos, table, want):
    np.testing.assert_array_equal(got_pos, want["pos_ids"])
    assert table.shape == want["table"].shape
    np.testing.assert_array_equal(table, want["table"])
    np.testing.assert_allclose(got_mean, want["mean_rain64"], rtol=1e-6, atol=1e-9)
    assert want["table"].max() > 0 and (want["pos_ids"] == 1).any() and want["pos_ids"].max() == table.shape[0]


@pytest.mark.parametrize("pp", [1.0, 2.0, 1.7])
def test_rain_idw(pp, tmp_path):
    import oracle
    from wmf_b200 import ModelsModule
    bs, xy, coord, rain = setup_case()
    want = oracle.rain_idw(xy, coord, rain, pp, 0.0)
    m = ModelsModule()
    path = str(tmp_path / "idw.bin")
    mean, pos = m.rain_idw(xy, coord, rain, pp, 0, path, 0.0, np.ones(bs["n"]), bs["n"], coord.shape[1], rain.shape[1])
    check(mean, pos, np.fromfile(path, np.int32).reshape(-1, bs["n"]), want)


def test_rain_idw_threshold_and_reference_file(tmp_path):
    """umbral > 0 drops the drizzle fields (:1244); the file equals the one the reference's own Fortran writes."""
    import oracle
    from oracle import ref
    from wmf_b200 import ModelsModule
    bs, xy, coord, rain = setup_case(seed=4)
    all_wet = oracle.rain_idw(xy, coord, rain, 2.0, 0.0)
    umbral = float(np.median(all_wet["table"][1:].max(axis=1))) / 1000.0     # half of the wet fields never exceed it
    want = oracle.rain_idw(xy, coord, rain, 2.0, umbral)
    assert (want["pos_ids"] == 1).sum() > (all_wet["pos_ids"] == 1).sum()
    m = ModelsModule()
    path = str(tmp_path / "idw.bin")
    mean, pos = m.rain_idw(xy, coord, rain, 2.0, 0, path, umbral, np.ones(bs["n"]))
    check(mean, pos, np.fromfile(path, np.int32).reshape(-1, bs["n"]), want)
    if ref.available():
        os.makedirs(str(tmp_path / "ref"))
        r = ref.rain_idw(xy, coord, rain, 2.0, umbral, workdir=str(tmp_path / "ref"))
        assert open(path, "rb").read() == open(str(tmp_path / "ref" / "idw.bin"), "rb").read()
        np.testing.assert_array_equal(r["pos_ids"], pos)


def test_rain_mit(tmp_path):
    import oracle
    from scipy.spatial import Delaunay
    from wmf_b200 import ModelsModule
    bs, xy, coord, rain = setup_case(seed=5)
    tin = np.asfortranarray((Delaunay(coord.T.astype(np.float64)).simplices.T + 1).astype(np.int32))
    # triangle of every cell: barycentric test as rain_pre_mit (modelosv2.f90:1068-1101), last hit wins
    perte = np.zeros(bs["n"], np.int32)
    for k in range(tin.shape[1]):
        p0, p1, p2 = (coord[:, tin[j, k] - 1].astype(np.float64) for j in range(3))
        area = abs(0.5 * (-p1[1] * p2[0] + p0[1] * (-p1[0] + p2[0]) + p0[0] * (p1[1] - p2[1]) + p1[0] * p2[1]))
        s = 1 / (2 * area) * (p0[1] * p2[0] - p0[0] * p2[1] + (p2[1] - p0[1]) * xy[0] + (p0[0] - p2[0]) * xy[1])
        t = 1 / (2 * area) * (p0[0] * p1[1] - p0[1] * p1[0] + (p0[1] - p1[1]) * xy[0] + (p1[0] - p0[0]) * xy[1])
        perte[(s > 0) & (t > 0) & (s + t < 1)] = k + 1
    assert perte.min() >= 1
    want = oracle.rain_mit(xy, coord, rain, tin, perte, 0.0)
    m = ModelsModule()
    # rain_pre_mit on the GPU: bit-exact against the oracle's (its looser test -- only s > 0 and t > 0 -- picks other triangles)
    np.testing.assert_array_equal(m.rain_pre_mit(xy, tin, coord).reshape(-1), oracle.rain_pre_mit(xy, tin, coord))
    path = str(tmp_path / "mit.bin")
    mean, pos = m.rain_mit(xy, coord, rain, tin, perte[None, :], 0, path, 0.0, np.ones(bs["n"]))
    check(mean, pos, np.fromfile(path, np.int32).reshape(-1, bs["n"]), want)
    with pytest.raises(ValueError):
        m.rain_mit(xy, coord, rain, tin, np.zeros(bs["n"]), 0, path, 0.0, np.ones(bs["n"]))
    with pytest.raises(NotImplementedError):
        m.rain_mit(xy, coord, rain, tin, perte[None, :], 5, path, 0.0, np.full(bs["n"], 2))     # hills branch of rain_mit: unreachable


def test_resident_table_feeds_shia_without_leaving_the_gpu():
    """Stations -> rain_idw(resident=True) -> shia_v1(None, None, ...): the same hydrograph as with the host table."""
    import oracle
    from conftest import basin_vectors
    from wmf_b200 import ModelsModule, synth
    bs, xy, coord, rain = setup_case(nc=100, nr=80, seed=6, T=60)
    v = basin_vectors(bs)
    inp = synth.make_shia_inputs(bs["basin"], v["acum"], v["lng"], v["pend"], dxp=30.0, n_reg=60, seed=6)
    want = oracle.rain_idw(xy, coord, rain, 2.0, 0.0)
    inp["rain"], inp["pos_evento"] = want["table"], want["pos_ids"]
    ref_run = oracle.shia_v1(inp)
    m = ModelsModule()
    synth.apply_to_models(m, inp)
    mean, pos = m.rain_idw(xy, coord, rain, 2.0, 0, None, 0.0, np.ones(bs["n"]), resident=True)
    np.testing.assert_array_equal(pos, want["pos_ids"])
    n_cont = int(np.count_nonzero(inp["control"])) + 1
    ret = m.shia_v1(None, None, inp["calib"], n_cont, 1, 60, None, None, bs["n"])
    q, qr = np.asarray(ret[0], np.float64), np.asarray(ref_run["q"], np.float64)
    assert np.abs(q - qr).max() <= 1e-5 * np.abs(qr).max() and qr.max() > 0
    np.testing.assert_array_equal(m.acum_rain.reshape(-1), ref_run["acum_rain"])


def test_rain_idw_by_hills(tmp_path):
    """The hills model of rain_idw (modelosv2.f90:1210-1216, 1257-1259): records of nhills values, each the SUM of the hill's
    cells in ascending cell order (module models' basin_subbasin_map2subbasin with sum_mean = 2), column i = hill nhills-i+1."""
    import oracle
    from conftest import basin_vectors
    from wmf_b200 import CuModule, ModelsModule
    bs, xy, coord, rain = setup_case(seed=7)
    v = basin_vectors(bs)
    oc = bs["cu"]
    cauce, nodos, nn = oc.basin_subbasin_nod(bs["basin"], v["acum"], 25)
    hills_own, _ = oc.basin_subbasin_find(bs["basin"], nodos, nn)
    want = oracle.rain_idw(xy, coord, rain, 2.0, 0.0, mask=hills_own, nhills=nn)
    m = ModelsModule()
    path = str(tmp_path / "idw_hills.bin")
    mean, pos = m.rain_idw(xy, coord, rain, 2.0, nn, path, 0.0, hills_own, bs["n"], coord.shape[1], rain.shape[1])
    table = np.fromfile(path, np.int32).reshape(-1, nn)
    np.testing.assert_array_equal(pos, want["pos_ids"])
    np.testing.assert_array_equal(table, want["table"])
    np.testing.assert_allclose(mean, want["mean_rain64"], rtol=1e-6, atol=1e-9)
    assert nn > 50 and table.shape[1] == nn and table.max() > 0


def test_rain_idw_gauge_on_a_cell_centre(tmp_path):
    """A gauge exactly on a cell centre gives that cell an infinite weight: the masked sums must then really skip the
    stations without data (inf * 0), so the library falls back from its mask-free fast path; same records as the oracle."""
    import oracle
    from wmf_b200 import ModelsModule
    bs, xy, coord, rain = setup_case(seed=8)
    coord = np.array(coord, copy=True, order="F")
    coord[:, 2] = xy[:, 1234]
    coord[:, 5] = xy[:, 77]
    want = oracle.rain_idw(xy, coord, rain, 2.0, 0.0)
    m = ModelsModule()
    path = str(tmp_path / "idw.bin")
    mean, pos = m.rain_idw(xy, coord, rain, 2.0, 0, path, 0.0, np.ones(bs["n"]))
    np.testing.assert_array_equal(pos, want["pos_ids"])
    np.testing.assert_array_equal(np.fromfile(path, np.int32).reshape(-1, bs["n"]), want["table"])
