"""The tile decomposition behind csrc/pathfind.cu, restated on the CPU (oracle/tile_model.py), against the oracle's FIFO
search (= the reference's basin_find / basin_cut / basin_acum, cuencas.f90:525-607, 659-680): identical tables for tiny
tiles (many boundary crossings), partial basins with up to 8 tributaries per cell, and the bench generators."""
import numpy as np
import pytest

from wmf_b200 import synth


def _check(DIR, oc_, orow, nc, nr, ts, dx=30.0):
    import oracle
    from oracle import tile_model as tm
    oc = oracle.Cu()
    oc.ncols, oc.nrows, oc.xll, oc.yll, oc.dx, oc.dy, oc.dxp, oc.nodata = nc, nr, 0.0, 0.0, dx, dx, dx, synth.NODATA
    n = oc.basin_find(dx * (oc_ - 0.5), dx * (nr - orow + 0.5), DIR)
    b = oc.basin_cut(n)
    a = np.asarray(oc.basin_acum(b)).reshape(-1)
    ql, qp, qa, _ = tm.tile_find(DIR, oc_ - 1, orow - 1, ts)
    np.testing.assert_array_equal(tm.basin_table(ql, qp, nc), b)
    np.testing.assert_array_equal(qa[::-1], a)
    return n


@pytest.mark.parametrize("ts", [4, 8, 16])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_partial_basins(ts, seed):
    nc, nr = 37 + 3 * seed, 29 + 2 * seed
    DIR, (oc_, orow) = synth.make_partial_raster(nc, nr, 100 + seed, noise=1.5)
    assert _check(DIR, oc_, orow, nc, nr, ts) > 100


@pytest.mark.parametrize("kind", ["G1", "G2"])
def test_bench_generators(kind):
    nc, nr = 41, 33
    DIR, _ = synth.make_raster(kind, nc, nr, 1)
    oc_, orow = synth.outlet_colrow(nc, nr)
    assert _check(DIR, oc_, orow, nc, nr, 8) == (nc - 2) * (nr - 2)


def test_torch_raster_generator_matches_numpy():
    for kind, nc, nr, seed in (("G2", 97, 83, 0), ("G2", 300, 200, 3), ("G1", 120, 90, 1)):
        D, _ = synth.make_raster(kind, nc, nr, seed)
        t = synth.make_dir_torch(kind, nc, nr, seed, device="cpu", chunk_rows=64)
        np.testing.assert_array_equal(np.ascontiguousarray(D.T), t.numpy())
