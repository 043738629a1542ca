"""GPU parity, Path B (module models): the CUDA wavefront through the C ABI vs the CPU oracle.

Tolerance (BASELINE.json north_star): simulated discharge and tank storages within 1e-5 relative in fp32,
with an absolute floor scaled by the series maximum for near-zero flows.  `balance` is a
catastrophic-cancellation quantity (modelosv2.f90:846): the GPU accumulates it in fp64 and is compared with
the oracle's fp64 restatement of the same sum, scaled by the total storage.
"""
import os

import numpy as np
import pytest

from conftest import basin_vectors, make_basin

pytestmark = pytest.mark.gpu

RTOL = 1e-5


@pytest.fixture(autouse=True, params=["auto", "4"])
def time_block(request, monkeypatch):
    """Every test runs twice: with the library's own choice of steps per time block (1 for basins this small) and with
    the K = 4 blocks the C2-sized runs use."""
    if request.param != "auto":
        monkeypatch.setenv("WMF_SHIA_K", request.param)
    return request.param


def close(got, ref, rtol=RTOL, afloor=1e-7, name=""):
    """|got-ref| <= rtol*|ref| + afloor*max|ref|.  The absolute floor is one fp32 epsilon of the largest value of
    the field: residues of the cancellations the model itself performs (e.g. Qskr - qsSUStot) sit below it."""
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    scale = max(float(np.abs(ref).max()), 1e-30)
    err = np.abs(got - ref)
    tol = rtol * np.abs(ref) + afloor * scale
    bad = err > tol
    assert not bad.any(), f"{name}: {bad.sum()} of {bad.size} off; worst excess {np.max(err / tol):.3e} x tol"


def build(kind="G2", nc=96, nr=80, seed=0, n_reg=120, **kw):
    import oracle
    from wmf_b200 import synth
    bs = make_basin(kind, nc, nr, seed)
    v = basin_vectors(bs)
    inp = synth.make_shia_inputs(bs["basin"], v["acum"], v["lng"], v["pend"], dxp=30.0, n_reg=n_reg, seed=seed, **kw)
    return bs, inp


def run_gpu(inp, flags=None, stoin=None, hspeedin=None):
    from wmf_b200 import ModelsModule, synth
    m = ModelsModule()
    synth.apply_to_models(m, inp)
    for k, v in (flags or {}).items():
        setattr(m, k, v)
    m.set_rain(inp["rain"], inp["pos_evento"])
    n_cont = int(np.count_nonzero(inp["control"])) + 1
    n_conth = max(1, int(np.count_nonzero(inp["control_h"])))
    ret = m.shia_v1(None, None, inp["calib"], n_cont, n_conth, inp["n_reg"], stoin, hspeedin, inp["nceldas"])
    return m, ret


def run_oracle(inp, flags=None, stoin=None, hspeedin=None):
    import oracle
    d = dict(inp)
    d.update(flags or {})
    d["stoin"], d["hspeedin"] = stoin, hspeedin
    return oracle.shia_v1(d)


def compare_core(ret, m, ref, N):
    q, qsed, qsep, hum, st1, st3, balance, speed, area, stoout, qrain = ret
    close(q, ref["q"], name="Q")
    close(stoout, ref["stoout"], name="StoOut")
    close(hum, ref["hum"], name="Hum")
    close(st1, ref["st1"], name="St1")
    close(st3, ref["st3"], name="St3")
    close(m.mean_rain.reshape(-1), ref["mean_rain"], name="Mean_Rain")
    np.testing.assert_array_equal(m.acum_rain.reshape(-1), ref["acum_rain"])
    # balance: fp64 on both sides; noise floor = fp32 rounding of the state (5N values) ~ 1e-7 * total / sqrt(N)
    tot = float(np.abs(ref["stoout"]).sum())
    assert np.abs(balance - ref["balance64"]).max() <= 2e-7 * tot, (np.abs(balance - ref["balance64"]).max(), tot)
    assert np.abs(ref["balance64"]).max() <= 1e-6 * tot        # the model conserves mass
    return q


def test_baseline_linear():
    bs, inp = build(n_control_h=4)
    m, ret = run_gpu(inp, dict(show_storage=1, show_speed=1, show_mean_speed=1))
    ref = run_oracle(inp, dict(show_storage=1, show_speed=1, show_mean_speed=1))
    q = compare_core(ret, m, ref, bs["n"])
    assert q.shape == (17, inp["n_reg"]) and q[0].max() > 0
    close(ret[7], ref["speed"], name="Speed")
    close(ret[8], ref["areacontrol"], name="AreaControl")
    # means: the GPU accumulates in fp64; the reference's sequential fp32 sum is only good to ~N*eps/2
    close(m.mean_storage, ref["mean_storage64"], name="mean_storage")
    close(m.mean_speed, ref["mean_speed64"], name="mean_speed")
    close(ref["mean_storage"], ref["mean_storage64"], rtol=bs["n"] * 6e-8, name="oracle fp32 vs fp64 mean_storage")
    assert ret[9].min() >= 0.0


def test_linear_is_bit_exact_where_no_powf():
    """Without evaporation (Calib(1)=0 => Evp_loss = 0*powf = 0) every remaining operation is +,-,*,/ and min/max in the
    oracle's order, so the hillslope-only part of the model must agree to the last bit."""
    bs, inp = build("G1", 64, 48, 1, n_reg=60, thresholds=(10 ** 9, 10 ** 9 + 1))   # no channel cells
    inp["calib"][0] = 0.0
    m, ret = run_gpu(inp)
    ref = run_oracle(inp)
    np.testing.assert_array_equal(ret[9], ref["stoout"])
    np.testing.assert_array_equal(ret[0], ref["q"])


@pytest.mark.parametrize("speed_type", [(2, 2, 1), (2, 2, 2), (1, 2, 1)])
def test_kinematic_hillslope_and_returns(speed_type):
    bs, inp = build("G2", 120, 90, 2, n_reg=150, speed_type=speed_type, retorno_gr=1, retorno_aq=1)
    flags = dict(show_mean_retorno=0)
    m, ret = run_gpu(inp, flags)
    ref = run_oracle(inp, flags)
    compare_core(ret, m, ref, bs["n"])
    close(m.retorned.reshape(-1), ref["retorned"], name="Retorned")


def test_mean_retorno_mode():
    bs, inp = build("G2", 100, 70, 3, n_reg=100, retorno_gr=1, rain_peak_mm=80.0)
    inp["max_gravita"][:] = 4.0                      # tiny gravitational store => returns happen
    flags = dict(show_mean_retorno=1)
    m, ret = run_gpu(inp, flags)
    ref = run_oracle(inp, flags)
    compare_core(ret, m, ref, bs["n"])
    assert ref["mean_retorno"].max() > 0
    close(m.mean_retorno, ref["mean_retorno64"], name="mean_retorno")
    assert np.all(m.retorned == 0)


def test_resume_from_stoout_and_hspeed():
    """Checkpoint/resume: running T1 then T2 steps from (StoOut, hspeed) equals running T1+T2."""
    from wmf_b200 import ModelsModule, synth
    bs, inp = build("G2", 90, 90, 4, n_reg=80, speed_type=(2, 2, 1))
    m, full = run_gpu(inp)
    ref_full = run_oracle(inp)
    T1 = 37
    a = dict(inp); a["n_reg"] = T1; a["pos_evento"] = inp["pos_evento"][:T1]; a["evpserie"] = inp["evpserie"][:T1]
    m1, r1 = run_gpu(a)
    hs1 = m1._run(np.asarray(a["calib"], np.float32)[None, :], 17, 1, T1, a["rain"], a["pos_evento"])["hspeed_out"]
    b = dict(inp); b["n_reg"] = inp["n_reg"] - T1; b["pos_evento"] = inp["pos_evento"][T1:]; b["evpserie"] = inp["evpserie"][T1:]
    m2, r2 = run_gpu(b, stoin=r1[9], hspeedin=hs1)
    close(np.concatenate([r1[0], r2[0]], axis=1), full[0], name="Q resumed")
    close(r2[9], full[9], name="StoOut resumed")
    ro = run_oracle(b, stoin=r1[9], hspeedin=hs1)
    close(r2[0], ro["q"], name="Q resumed vs oracle")


def test_sediments():
    bs, inp = build("G2", 110, 90, 5, n_reg=120, speed_type=(2, 2, 1), sediments=True)
    m, ret = run_gpu(inp)
    ref = run_oracle(inp)
    compare_core(ret, m, ref, bs["n"])
    assert ref["qsed"].max() > 0
    close(ret[1], ref["qsed"], name="Qsed")
    for k in ("vs", "vd", "vsc", "vdc"):
        close(getattr(m, k), ref[k], name=k)
    close(m.volero, ref["volero"], name="VolERO")
    close(m.voldepo, ref["voldepo"], name="VolDEPo")
    close(m.erot, ref["erot64"], name="EROt")     # global running sums: fp64 on the GPU, fp64 twin in the oracle
    close(m.dept, ref["dept64"], name="DEPt")
    close(ref["erot"], ref["erot64"], rtol=2e-3, name="oracle fp32 vs fp64 EROt")


@pytest.mark.parametrize("gullie", [0, 1])
def test_slides(gullie):
    bs, inp = build("G2", 110, 90, 6, n_reg=120, slides=True, retorno_gr=1, rain_peak_mm=60.0)
    inp["sl_gullienogullie"] = gullie
    m, ret = run_gpu(inp)
    ref = run_oracle(inp)
    compare_core(ret, m, ref, bs["n"])
    # The soil thresholds (slide_allocate, modelosv2.f90:1630-1639) divide by differences of tangents that vanish
    # for some angle pairs (tan rs = tan fa; gw tan fa = gs (tan fa - tan rs)): next to those poles a 1-ulp
    # difference between libm's and the GPU's tangent is amplified without bound.  So: 1e-5 on >= 99.5 % of the
    # cells, median error at fp32 rounding level, and the risk classification derived from them identical.
    for k in ("sl_zmin", "sl_zcrit", "sl_zmax", "sl_bo"):
        got, want = getattr(m, k).reshape(-1).astype(np.float64), ref[k].astype(np.float64)
        rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-30)
        assert np.mean(rel <= 1e-5) >= 0.995 and np.median(rel) <= 2e-7, (k, np.mean(rel <= 1e-5), np.median(rel))
    np.testing.assert_array_equal(m.sl_riskvector.reshape(-1), ref["sl_riskvector"])
    assert ref["sl_slidencelltime"].sum() > 0 and len(np.unique(ref["sl_slideocurrence"])) >= 2
    # failures are threshold events: allow a handful of cells to flip when Zw sits within 1e-5 of Zcrit
    occ_diff = (m.sl_slideocurrence.reshape(-1) != ref["sl_slideocurrence"]).sum()
    assert occ_diff <= max(2, bs["n"] // 2000), occ_diff
    if occ_diff == 0:
        np.testing.assert_array_equal(m.sl_slideacumulate.reshape(-1), ref["sl_slideacumulate"])
        np.testing.assert_array_equal(m.sl_slidencelltime, ref["sl_slidencelltime"])


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_slide_marks_with_scrambled_unit_types(seed):
    """slide_hill2gullie (modelosv2.f90:1683-1707) marks downstream "while the unit type does not grow".  wmf.py derives the
    types from the accumulated area, so they never decrease downstream; here a fifth of them is re-drawn at random (hillslope
    cells below channel cells, gullies below rivers), which exercises every branch of the rule the wavefront ring implements:
    the landslide maps, counters and per-step cell counts must still be the oracle's."""
    bs, inp = build("G2", 100, 84, 30 + seed, n_reg=90, slides=True, retorno_gr=1, rain_peak_mm=70.0)
    rng = np.random.default_rng(seed)
    ut = np.array(inp["unit_type"], copy=True)
    pick = rng.random(ut.shape[1]) < 0.2
    pick[-1] = False                                           # the outlet stays a channel cell
    ut[0, pick] = rng.integers(1, 4, int(pick.sum()))
    inp["unit_type"] = ut
    inp["sl_gullienogullie"] = 1
    m, ret = run_gpu(inp)
    ref = run_oracle(inp)
    close(ret[0], ref["q"], name="Q")
    np.testing.assert_array_equal(m.sl_riskvector.reshape(-1), ref["sl_riskvector"])
    assert ref["sl_slidencelltime"].sum() > 0 and (ref["sl_slideocurrence"] == 3.0).any()
    occ_diff = (m.sl_slideocurrence.reshape(-1) != ref["sl_slideocurrence"]).sum()
    assert occ_diff <= max(2, bs["n"] // 2000), occ_diff
    if occ_diff == 0:
        np.testing.assert_array_equal(m.sl_slideacumulate.reshape(-1), ref["sl_slideacumulate"])
        np.testing.assert_array_equal(m.sl_slidencelltime, ref["sl_slidencelltime"])


def test_sediments_and_slides_together():
    bs, inp = build("G1", 90, 120, 7, n_reg=90, speed_type=(2, 2, 1), sediments=True, slides=True)
    m, ret = run_gpu(inp)
    ref = run_oracle(inp)
    compare_core(ret, m, ref, bs["n"])
    close(ret[1], ref["qsed"], name="Qsed")


def test_rain_files_roundtrip(tmp_path):
    """shia_v1 fed from a .bin/.hdr pair in the reference's on-disk format equals the in-memory run."""
    from wmf_b200 import ModelsModule, synth
    from wmf_b200._models import write_rain_files
    bs, inp = build("G2", 80, 60, 8, n_reg=50)
    write_rain_files(str(tmp_path / "rain.bin"), str(tmp_path / "rain.hdr"), inp["rain"], inp["pos_evento"])
    m = ModelsModule()
    synth.apply_to_models(m, inp)
    ret = m.shia_v1(str(tmp_path / "rain.bin"), str(tmp_path / "rain.hdr"), inp["calib"], 17, 1, 50, None, None, bs["n"])
    _, mem = run_gpu(inp)
    np.testing.assert_array_equal(ret[0], mem[0])
    np.testing.assert_array_equal(ret[9], mem[9])


def test_bad_topology_is_rejected():
    from wmf_b200 import WmfError
    bs, inp = build("G2", 64, 48, 9, n_reg=10)
    d = np.array(inp["drena"], order="F", copy=True)
    d[0, 3] = bs["n"] - 1            # parent index 1 < cell index 3: not upstream-first
    inp2 = dict(inp); inp2["drena"] = d
    with pytest.raises(WmfError):
        run_gpu(inp2)


def test_ensemble_matches_single_runs():
    """M members in one launch (member-minor layout) == M independent single runs, bit for bit."""
    from wmf_b200 import ModelsModule, synth
    bs, inp = build("G2", 80, 64, 10, n_reg=60, speed_type=(2, 1, 1))
    calibs = synth.make_ensemble_calibs(40, seed=1)
    m = ModelsModule()
    synth.apply_to_models(m, inp)
    ens = m._run(calibs, 17, 1, 60, inp["rain"], inp["pos_evento"])
    assert ens["q"].shape == (17, 60, 40)
    for k in (0, 7, 39):
        one = dict(inp); one["calib"] = calibs[k]
        _, ret = run_gpu(one)
        np.testing.assert_array_equal(ens["q"][:, :, k], ret[0])
        np.testing.assert_array_equal(ens["stoout"][:, :, k], ret[9])
    # scores (wmf.py:749-754, 775-781) against member 0's hydrograph
    import ctypes as C
    c = m.ctx
    qobs = np.ascontiguousarray(ens["q"][0, :, 0]).astype(np.float32)
    qobs[5] = np.nan
    scores = np.empty((40, 2), np.float32)
    c.check(c.L.wmf_eval_scores(c.h, ens["q"].ctypes.data, 40, 17, 60, 0, qobs.ctypes.data, scores.ctypes.data))
    for k in (0, 3, 39):
        s_o, s_s = qobs.astype(np.float64), ens["q"][0, :, k].astype(np.float64)
        med = np.nansum(s_o) / np.sum(np.isfinite(s_o))
        nash = 1 - np.nansum((s_o - s_s) ** 2) / np.nansum((s_o - med) ** 2)
        mo, ms = np.nanargmax(s_o), np.nanargmax(s_s)
        qp = (s_o[mo] - s_s[ms]) / s_o[mo] * 100
        assert scores[k, 0] == pytest.approx(nash, rel=1e-5, abs=1e-6)
        assert scores[k, 1] == pytest.approx(qp, rel=1e-5, abs=1e-4)


def test_streamed_rain_and_pdl_are_bit_identical(monkeypatch):
    """The pinned-memory rain streaming (tiles copied next to the running wavefront) and programmatic dependent launch
    change scheduling only: results must equal the plain path bit for bit."""
    import torch
    from wmf_b200 import ModelsModule, synth
    bs, inp = build("G2", 100, 80, 21, n_reg=90)
    _, plain = run_gpu(inp)
    pin = torch.from_numpy(np.ascontiguousarray(inp["rain"])).pin_memory()
    monkeypatch.setenv("WMF_RAIN_STREAM_MIN_MB", "0")
    monkeypatch.setenv("WMF_RAIN_TILES", "5,4")
    m = ModelsModule()
    synth.apply_to_models(m, inp)
    m.set_rain(pin.numpy(), inp["pos_evento"])
    n_cont = int(np.count_nonzero(inp["control"])) + 1
    ret = m.shia_v1(None, None, inp["calib"], n_cont, 1, inp["n_reg"], None, None, inp["nceldas"])
    for k in (0, 6, 9):
        np.testing.assert_array_equal(ret[k], plain[k])
    monkeypatch.setenv("WMF_SHIA_PDL", "0")
    ret2 = m.shia_v1(None, None, inp["calib"], n_cont, 1, inp["n_reg"], None, None, inp["nceldas"])
    for k in (0, 6, 9):
        np.testing.assert_array_equal(ret2[k], plain[k])
    monkeypatch.setenv("WMF_SHIA_K", "0")          # the one-step-per-launch kernel streams too
    ret3 = m.shia_v1(None, None, inp["calib"], n_cont, 1, inp["n_reg"], None, None, inp["nceldas"])
    close(ret3[0], plain[0], name="Q one-step kernel")


def test_evaluate_population_single_gpu():
    """wmf_b200.ensemble on one GPU (no process group): batched members + device scores == per-member numpy scores."""
    from wmf_b200 import ModelsModule, synth
    from wmf_b200.ensemble import evaluate_population
    bs, inp = build("G2", 72, 60, 22, n_reg=50)
    calibs = synth.make_ensemble_calibs(10, seed=3)
    m = ModelsModule()
    synth.apply_to_models(m, inp)
    n_cont = int(np.count_nonzero(inp["control"])) + 1
    one = dict(inp); one["calib"] = calibs[4]
    _, ret = run_gpu(one)
    qobs = ret[0][0].copy()
    sc = evaluate_population(m, calibs, qobs, n_cont, inp["n_reg"], inp["rain"], inp["pos_evento"], row=0,
                             members_per_launch=4).cpu().numpy()
    assert sc.shape == (10, 2)
    assert sc[4, 0] == pytest.approx(1.0, abs=1e-6) and abs(sc[4, 1]) < 1e-4      # the member that made qobs
    for k in (0, 9):
        o = dict(inp); o["calib"] = calibs[k]
        _, r = run_gpu(o)
        s_o, s_s = qobs.astype(np.float64), r[0][0].astype(np.float64)
        nash = 1 - ((s_o - s_s) ** 2).sum() / ((s_o - s_o.mean()) ** 2).sum()
        assert sc[k, 0] == pytest.approx(nash, rel=1e-5, abs=1e-6)


def test_sediment_slide_kernels_agree(monkeypatch):
    """SED + slides through the time-blocked kernel (K=4) and through the one-step-per-launch kernel: same per-cell fp32
    operations in the same order, so every per-cell result is identical; only the global fp64 totals (atomics) may differ
    in the last bits."""
    bs, inp = build("G2", 130, 100, 23, n_reg=75, speed_type=(2, 2, 1), sediments=True, slides=True, retorno_gr=1,
                    rain_peak_mm=60.0)
    inp["sl_gullienogullie"] = 1
    m1, r1 = run_gpu(inp)
    monkeypatch.setenv("WMF_SHIA_K", "0")
    m0, r0 = run_gpu(inp)
    for k in (0, 1, 3, 9):
        np.testing.assert_array_equal(r1[k], r0[k])
    for k in ("vs", "vd", "vsc", "vdc", "volero", "voldepo", "sl_slideocurrence", "sl_slideacumulate", "sl_slidencelltime",
              "retorned"):
        np.testing.assert_array_equal(getattr(m1, k), getattr(m0, k))
    close(m1.erot, m0.erot, rtol=1e-6, name="EROt")
    close(m1.dept, m0.dept, rtol=1e-6, name="DEPt")
    assert r1[1].max() > 0 and m1.sl_slidencelltime.sum() > 0


def test_storage_and_speed_dumps(tmp_path):
    """save_storage / save_speed (modelosv2.f90:799-803, 815-818): record guarda_cond(t) of the .StObin file holds StoOut
    after step t and record t of the speed file holds hspeed -- i.e. exactly what a run truncated after step t returns."""
    from wmf_b200 import ModelsModule, synth
    from wmf_b200._models import read_float_basin_ncol
    bs, inp = build("G2", 90, 70, 24, n_reg=40, speed_type=(2, 1, 1))
    N, T = bs["n"], 40
    guarda = np.zeros(T, np.int32)
    guarda[[4, 17, 39]] = [1, 2, 3]                  # record numbers (record 1 replaces the file, modelosv2.f90:937-941)
    m = ModelsModule()
    synth.apply_to_models(m, inp)
    m.set_rain(inp["rain"], inp["pos_evento"])
    m.save_storage, m.save_speed = 1, 1
    m.guarda_cond = guarda
    n_cont = int(np.count_nonzero(inp["control"])) + 1
    sto_path, spd_path = str(tmp_path / "s.StObin"), str(tmp_path / "v.bin")
    full = m.shia_v1(None, None, inp["calib"], n_cont, 1, T, None, None, N, sto_path, spd_path)
    hs_full = m._run(np.asarray(inp["calib"], np.float32)[None, :], n_cont, 1, T, inp["rain"], inp["pos_evento"])
    for t, rec in ((4, 1), (17, 2), (39, 3)):
        a = dict(inp); a["n_reg"] = t + 1; a["pos_evento"] = inp["pos_evento"][:t + 1]; a["evpserie"] = inp["evpserie"][:t + 1]
        mt = ModelsModule()
        synth.apply_to_models(mt, a)
        part = mt._run(np.asarray(a["calib"], np.float32)[None, :], n_cont, 1, t + 1, a["rain"], a["pos_evento"])
        got, res = read_float_basin_ncol(sto_path, rec, N, 5)
        assert res == 0
        np.testing.assert_array_equal(got, part["stoout"])
        spd, res = read_float_basin_ncol(spd_path, t + 1, N, 4)
        assert res == 0
        np.testing.assert_array_equal(spd, part["hspeed_out"])
    np.testing.assert_array_equal(read_float_basin_ncol(sto_path, 3, N, 5)[0], full[9])
    assert os.path.getsize(spd_path) == 4 * 4 * N * T and os.path.getsize(sto_path) == 4 * 5 * N * 3


@pytest.mark.parametrize("show_means", [0, 1])
def test_vfluxes_rc_retorno_dumps(tmp_path, show_means):
    """save_vfluxes / save_rc / save_retorno (modelosv2.f90:515-524, 806-823, 863-867): record guarda_vfluxes(t) of the
    vfluxes file holds the four vertical fluxes of step t (vflux 2 and 3 carrying the return flow, :510-511), record t of
    the retorno file the return flow of that step, the rc file the two runoff-coefficient accumulators after the final
    clamp; mean_vfluxes is the per-step mean.  ``show_means`` = 1 sends the run through the one-step kernel."""
    from wmf_b200 import ModelsModule, synth
    from wmf_b200._models import read_float_basin_ncol
    bs, inp = build("G2", 90, 70, 27, n_reg=40, retorno_gr=1, rain_peak_mm=40.0)
    inp["calib"][9] = 0.02                           # small gravitational capacity: tank 3 does spill back into tank 2
    N, T = bs["n"], 40
    steps = (3, 20, 39)
    guarda = np.zeros(T, np.int32)
    guarda[list(steps)] = [1, 2, 3]
    flags = dict(save_vfluxes=1, save_rc=1, save_retorno=1, show_storage=show_means)
    m = ModelsModule()
    synth.apply_to_models(m, inp)
    m.set_rain(inp["rain"], inp["pos_evento"])
    for k, v in flags.items():
        setattr(m, k, v)
    m.guarda_vfluxes = guarda
    n_cont = int(np.count_nonzero(inp["control"])) + 1
    pv, pr, pc = str(tmp_path / "vf.bin"), str(tmp_path / "ret.bin"), str(tmp_path / "rc.bin")
    ret = m.shia_v1(None, None, inp["calib"], n_cont, 1, T, None, None, N, None, None, pv, None, None, None, None, pr, pc)
    ref = run_oracle(inp, flags)
    compare_core(ret, m, ref, N)
    close(m.mean_vfluxes, ref["mean_vfluxes64"], name="mean_vfluxes")
    close(ref["mean_vfluxes"], ref["mean_vfluxes64"], rtol=N * 6e-8, name="oracle fp32 vs fp64 mean_vfluxes")
    close(m.rc_coef, ref["rc_coef"], name="rc_coef")
    rc_file, res = read_float_basin_ncol(pc, 1, N, 2)
    assert res == 0
    np.testing.assert_array_equal(rc_file, m.rc_coef)
    assert (ref["rc_coef"][0] <= ref["rc_coef"][1]).all() and ref["rc_coef"][0].max() > 0
    assert not m.retorned.any()                      # Retorned is reset after every step in this mode (:841-843)
    seen_return = False
    for t, rec in zip(steps, (1, 2, 3)):
        a = dict(inp); a["n_reg"] = t + 1; a["pos_evento"] = inp["pos_evento"][:t + 1]; a["evpserie"] = inp["evpserie"][:t + 1]
        part = run_oracle(a, flags)
        got, res = read_float_basin_ncol(pv, rec, N, 4)
        assert res == 0
        close(got, part["vfluxes_last"], name=f"vfluxes record {rec}")
        r, res = read_float_basin_ncol(pr, t + 1, N, 1)
        assert res == 0
        close(r.reshape(-1), part["retorned_last"], name=f"retorno record {t + 1}")
        seen_return |= bool(part["retorned_last"].max() > 0)
    assert seen_return
    assert os.path.getsize(pv) == 4 * 4 * N * 3 and os.path.getsize(pr) == 4 * N * T and os.path.getsize(pc) == 4 * 2 * N
    # the dumps do not perturb the run
    m0, ret0 = run_gpu(inp, dict(show_storage=show_means))
    np.testing.assert_array_equal(ret0[0], ret[0])
    np.testing.assert_array_equal(ret0[9], ret[9])


def test_shia_on_random_partial_basin():
    """SHIA on the irregular tree of a random partial basin (pits, holes, up to 6-8 cells draining into one, several
    channel heads): same tolerance as the regular generators."""
    import oracle
    from wmf_b200 import synth
    nc, nr, dx = 150, 110, 30.0
    DIR, (oc, orow) = synth.make_partial_raster(nc, nr, 7, noise=1.5)
    x, y = dx * (oc - 0.5), dx * (nr - orow + 0.5)
    ocu = oracle.Cu()
    ocu.ncols, ocu.nrows, ocu.xll, ocu.yll, ocu.dx, ocu.dy, ocu.dxp, ocu.nodata = nc, nr, 0.0, 0.0, dx, dx, dx, synth.NODATA
    n = ocu.basin_find(x, y, DIR)
    b = ocu.basin_cut(n)
    rng = np.random.default_rng(7)
    hop = np.zeros(n + 1)
    par = n - b[0]
    for i in range(n - 2, -1, -1):
        hop[i] = hop[par[i]] + 1
    demv = (10 + 0.5 * hop[:n] + 0.2 * rng.random(n)).astype(np.float32)
    dv = ocu.basin_map2basin(b, DIR.astype(np.float32), 0, 0, dx, dx, synth.NODATA).astype(np.int32)
    acum, lng, pend, _ = ocu.basin_basics(b, demv, dv)
    inp = synth.make_shia_inputs(b, acum, lng, pend, dxp=dx, n_reg=90, seed=7, thresholds=(20, 300), speed_type=(2, 1, 1),
                                 retorno_gr=1, n_control_h=3)
    nch = np.bincount(par[par < n], minlength=n)
    assert nch.max() >= 5 and (np.asarray(inp["unit_type"]) > 1).sum() > 50
    m, ret = run_gpu(inp)
    ref = run_oracle(inp)
    compare_core(ret, m, ref, n)


def test_separate_fluxes():
    """separate_fluxes = 1 (modelosv2.f90:658-680, 697-704, 752-759): the runoff / subsurface / base-flow shares of the
    discharge at the outlet and the control points, including the reference's "an emptied channel zeroes the receiver's
    shares" rule, which makes the gather order-dependent."""
    bs, inp = build("G2", 110, 84, 25, n_reg=100, rain_peak_mm=50.0)
    inp["separate_fluxes"] = 1
    m, ret = run_gpu(inp, dict(separate_fluxes=1))
    ref = run_oracle(inp, dict(separate_fluxes=1))
    compare_core(ret, m, ref, bs["n"])
    assert ref["qseparated"].max() > 0 and ret[2].shape == ref["qseparated"].shape
    close(ret[2], ref["qseparated"], name="Qseparated")
    # without the switch the third return value is the zero array f2py would hand back
    m0, ret0 = run_gpu(inp, dict(separate_fluxes=0))
    assert not ret0[2].any()
    np.testing.assert_array_equal(ret0[0], ret[0])


def test_switch_combinations_in_one_call(tmp_path):
    """The reference runs any combination of its switches in ONE shia_v1 call (wmf.py:4634-4637 returns QsepDict and
    QsediDict of the same run; SimuBasin(SeparateFluxes='si', SaveStorage='si', ShowStorage='si') is an ordinary
    configuration).  separate_fluxes + separate_rain + sediments + slides + a storage dump + the per-step means together:
    every output against the oracle's single run, which is bit-identical to the reference's Fortran for each of them."""
    from wmf_b200 import ModelsModule, synth
    from wmf_b200._models import read_float_basin_ncol
    bs, inp = build("G2", 100, 80, 31, n_reg=72, rain_peak_mm=60.0, speed_type=(2, 2, 1), sediments=True, slides=True)
    types = synth.make_rain_types(inp, 31)
    T, N = inp["n_reg"], bs["n"]
    guarda = np.zeros(T, np.int32)
    guarda[[10, 40, T - 1]] = [1, 2, 3]
    flags = dict(separate_fluxes=1, separate_rain=1, sim_sediments=1, sim_slides=1, save_storage=1, show_storage=1,
                 show_mean_speed=1, show_speed=1)
    ref = run_oracle(dict(inp, **types), {k: v for k, v in flags.items() if k != "save_storage"})   # (the dump is checked below)
    m = ModelsModule()
    synth.apply_to_models(m, inp)
    for k, v in flags.items():
        setattr(m, k, v)
    m.guarda_cond = guarda
    m.set_rain(inp["rain"], inp["pos_evento"])
    m.set_rain_types(types["conv"], types["stra"], types["pos_conv"], types["pos_stra"])
    n_cont = int(np.count_nonzero(inp["control"])) + 1
    n_conth = max(1, int(np.count_nonzero(inp["control_h"])))
    sto = str(tmp_path / "sto.StObin")
    ret = m.shia_v1(None, None, inp["calib"], n_cont, n_conth, T, None, None, N, sto)
    compare_core(ret, m, ref, N)
    assert ref["qseparated"].max() > 0 and ref["qsep_byrain"].max() > 0 and ref["qsed"].max() > 0
    close(ret[2], ref["qseparated"], name="Qseparated")
    close(ret[10], ref["qsep_byrain"], name="Qsep_byrain")
    close(m.storage_conv, ref["storage_conv"], name="Storage_conv")
    close(m.storage_stra, ref["storage_stra"], name="Storage_stra")
    close(ret[1], ref["qsed"], name="Qsed")
    for k in ("vs", "vd", "vsc", "vdc"):
        close(getattr(m, k), ref[k], name=k)
    close(m.voldepo, ref["voldepo"], name="VolDEPo")
    np.testing.assert_array_equal(np.asarray(m.sl_slideocurrence).reshape(-1), ref["sl_slideocurrence"])
    close(ret[7], ref["speed"], name="Speed")
    close(m.mean_storage, ref["mean_storage64"], name="mean_storage")
    close(m.mean_speed, ref["mean_speed64"], name="mean_speed")
    rec, _ = read_float_basin_ncol(sto, 3, N, 5)               # the dump of the last step is the final storage
    close(rec, ref["stoout"], name="storage dump, record 3")
    # the switches are as the caller left them
    for k, v in flags.items():
        assert int(getattr(m, k)) == v


@pytest.mark.parametrize("variant", ["linear_return", "kinematic", "empty_capillary"])
def test_separate_rain(variant, tmp_path):
    """separate_rain = 1 (modelosv2.f90:346-359, 452-456, 471-474, 485-488, 530-547, 580-592, 605-608, 623-628, 645-655,
    683-686, 713-716, 767-770): convective / stratiform shadows of the five tanks and Qsep_byrain, including the stale rain
    types of dry steps, the integer division by 1000 and the 0/0 of an empty capillary tank.  The oracle is bit-identical
    to the reference's compiled Fortran on these cases (tests/test_reference_pin.py)."""
    from wmf_b200 import ModelsModule, synth
    from wmf_b200._models import write_rain_files
    kw = dict(rain_peak_mm=50.0)
    if variant == "linear_return":
        kw.update(retorno_gr=1)
    if variant == "kinematic":
        kw.update(speed_type=(2, 2, 1))
    bs, inp = build("G2", 100, 80, 28, n_reg=90, **kw)
    if variant == "linear_return":
        inp["calib"][9] = 0.02
    if variant == "empty_capillary":
        st = np.array(inp["storage"], copy=True)
        st[0, ::7] = 0.0
        inp["storage"], inp["storage_constant"] = st, 0.0
    types = synth.make_rain_types(inp, 28)
    ref = run_oracle(dict(inp, **types))
    m = ModelsModule()
    synth.apply_to_models(m, inp)
    m.set_rain(inp["rain"], inp["pos_evento"])
    m.separate_rain = 1
    n_cont = int(np.count_nonzero(inp["control"])) + 1
    T, N = inp["n_reg"], bs["n"]
    if variant == "kinematic":       # through the files, as wmf.py drives it (ruta_binConv, ruta_binStra, ruta_hdrConv, ruta_hdrStra)
        paths = {}
        for k in ("conv", "stra"):
            paths[k] = (str(tmp_path / (k + ".bin")), str(tmp_path / (k + ".hdr")))
            write_rain_files(paths[k][0], paths[k][1], types[k], np.concatenate([types["pos_" + k][:T], [1]]))
        ret = m.shia_v1(None, None, inp["calib"], n_cont, 1, T, None, None, N, None, None, None,
                        paths["conv"][0], paths["stra"][0], paths["conv"][1], paths["stra"][1])
    else:
        m.set_rain_types(types["conv"], types["stra"], types["pos_conv"], types["pos_stra"])
        ret = m.shia_v1(None, None, inp["calib"], n_cont, 1, T, None, None, N)
    compare_core(ret, m, ref, N)
    assert ref["qsep_byrain"].max() > 0 and ret[10].shape == ref["qsep_byrain"].shape
    close(ret[10], ref["qsep_byrain"], name="Qsep_byrain")
    close(m.storage_conv, ref["storage_conv"], name="Storage_conv")
    close(m.storage_stra, ref["storage_stra"], name="Storage_stra")
    assert np.isfinite(m.storage_conv).all()


def test_channel_cell_draining_into_hillslope_cell():
    """Unit types are user input: a channel cell may drain into a hillslope cell, whose tank 5 then fills and never
    drains (modelosv2.f90:617-621, 672): no kernel may assume that a hillslope cell's tank 5 is a constant of the run."""
    bs, inp = build("G2", 80, 70, 26, n_reg=60)
    ut = np.array(inp["unit_type"], copy=True).reshape(-1)
    d = np.asarray(inp["drena"]).reshape(3, -1, order="F")[0] if np.asarray(inp["drena"]).ndim == 2 else np.asarray(inp["drena"])
    n = ut.size
    ch = np.nonzero((ut > 1) & (d > 0))[0]
    pick = ch[len(ch) // 2]
    par = n - d[pick]
    ut[par] = 1                                   # a hillslope cell in the middle of the channel network
    inp["unit_type"] = ut[None, :]
    m, ret = run_gpu(inp)
    ref = run_oracle(inp)
    compare_core(ret, m, ref, bs["n"])
    assert ret[9][4, par] > inp["storage"][4, par]          # its channel tank did receive water


def test_run_scenarios_concurrent_contexts(time_block):
    """wmf_b200.ensemble.run_scenarios: storm scenarios on several contexts of one GPU give exactly the results of running
    them one after the other."""
    from wmf_b200 import ModelsModule, synth
    from wmf_b200.ensemble import run_scenarios
    bs, inp = build("G2", 90, 70, 33, n_reg=60, speed_type=(2, 1, 1))
    n_cont = int(np.count_nonzero(inp["control"])) + 1
    rains = [np.ascontiguousarray(np.roll(inp["rain"], 7 * k, axis=1)) for k in range(5)]

    def setup(m):
        synth.apply_to_models(m, inp)

    def run_one(m, rain):
        m.set_rain(rain, inp["pos_evento"])
        return m.shia_v1(None, None, inp["calib"], n_cont, 1, 60, None, None, bs["n"])[0].copy()

    par = run_scenarios(setup, rains, run_one, concurrency=3)
    m = ModelsModule()
    setup(m)
    for k, rain in enumerate(rains):
        np.testing.assert_array_equal(par[k], run_one(m, rain))
    assert not np.array_equal(par[0], par[1])


@pytest.mark.parametrize("kin", [False, True])
def test_hydroflash_floods(kin):
    """sim_floods = 1 (modelosv2.f90:728-743, flood_params :1730-1749, flood_debris_flow2 :1786-1822): the flooded-cell map
    ("last writer in (step, cell) order wins"), the hydraulic maps and cross sections of the evaluated channel cells and the
    module scalars of the last evaluation.  The oracle is bit-identical to the reference's Fortran on this sub-model."""
    from conftest import basin_vectors, make_basin
    from wmf_b200 import synth
    bs = make_basin("G2", 110, 84, 41)
    v = basin_vectors(bs)
    kw = dict(speed_type=(2, 1, 1)) if kin else {}
    inp = synth.make_shia_inputs(bs["basin"], v["acum"], v["lng"], v["pend"], dxp=30.0, n_reg=80, seed=41, thresholds=(20, 150),
                                 rain_peak_mm=70.0, **kw)
    inp.update(synth.make_flood_inputs(inp, v["elev"], v["pend"], 4, 41))
    m, ret = run_gpu(inp)
    ref = run_oracle(inp)
    compare_core(ret, m, ref, bs["n"])
    counts = np.bincount(ref["flood_flood"], minlength=3)
    assert counts[1] > 100 and counts[2] > 100
    got = np.asarray(m.flood_flood).reshape(-1)
    assert np.count_nonzero(got != ref["flood_flood"]) <= 2, np.count_nonzero(got != ref["flood_flood"])
    ev = ref["flood_flood"] == 2
    for k in ("flood_h", "flood_q", "flood_speed"):                      # the outflow and speed of the step, and their ratio
        close(np.asarray(getattr(m, k)).reshape(-1)[ev], ref[k][ev], name=k)
    # ufr = speed / (5.75 log(h/d50) + 6.25) loses digits where the denominator cancels, and Cr / Qsed follow it: these are
    # ill-conditioned functions of values that agree to 1e-5, so they are held to 1e-3 on all but a handful of cells
    for k in ("flood_ufr", "flood_cr", "flood_qsed"):
        g, r = np.asarray(getattr(m, k), np.float64).reshape(-1)[ev], np.asarray(ref[k], np.float64)[ev]
        ok = np.abs(g - r) <= 1e-3 * np.abs(r) + 1e-7 * np.abs(r).max()
        assert ok.mean() > 0.99, (k, ok.mean())
    prof_g, prof_r = np.asarray(m.flood_profundidad), ref["flood_profundidad"]
    same_level = np.abs(prof_g - prof_r).max(axis=0) < 1e-3            # a cell whose loop stopped at the same height
    assert same_level.mean() > 0.995
    assert abs(float(m.flood_diff) - float(ref["flood_scalars"][2])) < 1e-6
    close(np.float32(m.flood_rdf), ref["flood_scalars"][0], name="flood_rdf")


def test_hydroflash_together_with_other_switches():
    """sim_floods next to separate_fluxes and the per-step means in ONE shia_v1 call (three launch sequences behind it,
    _models._switch_passes): every output against the oracle's single run."""
    from conftest import basin_vectors, make_basin
    from wmf_b200 import synth
    bs = make_basin("G2", 110, 84, 41)
    v = basin_vectors(bs)
    inp = synth.make_shia_inputs(bs["basin"], v["acum"], v["lng"], v["pend"], dxp=30.0, n_reg=80, seed=41, thresholds=(20, 150),
                                 rain_peak_mm=70.0)
    inp.update(synth.make_flood_inputs(inp, v["elev"], v["pend"], 4, 41))
    flags = dict(separate_fluxes=1, show_storage=1, show_mean_speed=1)
    m, ret = run_gpu(inp, flags)
    ref = run_oracle(inp, flags)
    compare_core(ret, m, ref, bs["n"])
    close(ret[2], ref["qseparated"], name="Qseparated")
    close(m.mean_storage, ref["mean_storage64"], name="mean_storage")
    got = np.asarray(m.flood_flood).reshape(-1)
    assert np.count_nonzero(ref["flood_flood"] == 2) > 100
    assert np.count_nonzero(got != ref["flood_flood"]) <= 2
    ev = ref["flood_flood"] == 2
    close(np.asarray(m.flood_q).reshape(-1)[ev], ref["flood_q"][ev], name="flood_q")
    assert int(m.sim_floods) == 1 and int(m.separate_fluxes) == 1 and int(m.show_storage) == 1


def test_scenario_ensemble_with_sediments_and_slides():
    """Storm scenarios as a member axis (BASELINE configs[4]): three scenarios = three sequences of the same rain records
    run in ONE launch sequence with sediments + landslides; each is compared with a single oracle run of that sequence."""
    from wmf_b200 import ModelsModule, synth
    bs, inp = build("G2", 110, 90, 5, n_reg=60, speed_type=(2, 2, 1), sediments=True, slides=True, rain_peak_mm=60.0)
    N, T = bs["n"], 60
    pos = np.asarray(inp["pos_evento"], np.int32)
    scen = np.stack([pos, np.roll(pos, 7), pos[::-1].copy()]).astype(np.int32)
    m = ModelsModule()
    synth.apply_to_models(m, inp)
    n_cont = int(np.count_nonzero(inp["control"])) + 1
    n_conth = max(1, int(np.count_nonzero(inp["control_h"])))
    calibs = np.tile(np.asarray(inp["calib"], np.float32)[None, :], (3, 1))
    out = m._run(calibs, n_cont, n_conth, T, inp["rain"], scen)
    assert out["q"].shape == (n_cont, T, 3) and out["vs"].shape == (3, N, 3) and out["sl_slideocurrence"].shape == (N, 3)
    slid_any = 0
    for k in range(3):
        ref = run_oracle(dict(inp, pos_evento=scen[k]))
        close(out["q"][:, :, k], ref["q"], name=f"Q scenario {k}")
        close(out["stoout"][:, :, k], ref["stoout"], name=f"StoOut scenario {k}")
        close(out["qsed"][:, :, :, k], ref["qsed"], name=f"Qsed scenario {k}")
        for nm in ("vs", "vd", "vsc", "vdc"):
            close(out[nm][:, :, k], ref[nm], name=f"{nm} scenario {k}")
        close(out["volero"][:, k], ref["volero"], name=f"VolERO scenario {k}")
        close(out["voldepo"][:, k], ref["voldepo"], name=f"VolDEPo scenario {k}")
        close(out["erot"][:, k], ref["erot64"], name=f"EROt scenario {k}")
        close(out["dept"][:, k], ref["dept64"], name=f"DEPt scenario {k}")
        np.testing.assert_array_equal(out["sl_slideocurrence"][:, k], ref["sl_slideocurrence"])
        np.testing.assert_array_equal(out["sl_slideacumulate"][:, k], ref["sl_slideacumulate"])
        np.testing.assert_array_equal(out["sl_slidencelltime"][:, k], ref["sl_slidencelltime"])
        slid_any += int(np.count_nonzero(ref["sl_slideocurrence"]))
    assert slid_any > 0
    assert not np.array_equal(out["q"][:, :, 0], out["q"][:, :, 1])


def test_eval_scores_against_the_references_own_functions():
    """wmf_eval_scores vs `__eval_nash__` / `__eval_q_pico__` of wmf.py itself (wmf.py:749-754, 775-781): the golden table was
    produced by executing the reference's unmodified source (tests/golden/make_eval_scores.py), NaN gaps in the observations."""
    import torch
    from wmf_b200 import _lib
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "eval_scores.npz"))
    qobs, qsim, want = g["qobs"], g["qsim"], g["scores"]
    M, T = qsim.shape
    c = _lib.default_context()
    dev = f"cuda:{c.device}"
    d_q = torch.from_numpy(np.ascontiguousarray(qsim)).to(dev)         # (M, T) with n_cont = 1: member slowest
    d_o = torch.from_numpy(qobs).to(dev)
    sc = torch.empty((M, 2), dtype=torch.float32, device=dev)
    c.check(c.L.wmf_eval_scores(c.h, d_q.data_ptr(), M, 1, T, 0, d_o.data_ptr(), sc.data_ptr()))
    got = sc.cpu().numpy().astype(np.float64)
    np.testing.assert_allclose(got[:, 0], want[:, 0], rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(got[:, 1], want[:, 1], rtol=2e-6, atol=2e-5)
    assert want[0, 0] > 0.999


def test_nsga2_generation_loop_on_the_gpu():
    """calib_nsga2 (the loop of SimuBasin.Calib_NSGAII, wmf.py:4662-4733) with every generation evaluated in one ensemble
    launch sequence: the calibration that produced the "observations" must be matched better and better."""
    from wmf_b200 import ModelsModule, nsga2, synth
    bs, inp = build("G2", 72, 60, 22, n_reg=50)
    m = ModelsModule()
    synth.apply_to_models(m, inp)
    n_cont = int(np.count_nonzero(inp["control"])) + 1
    _, ret = run_gpu(inp)
    qobs = ret[0][0].copy()
    base = np.asarray(inp["calib"], np.float64)
    el = nsga2.NsgaElement(**{g: (0.5 * base[k], 1.5 * base[k]) if base[k] > 0 else (0.0, 0.0) for k, g in enumerate(nsga2.GENES)})
    ev = nsga2.make_gpu_evaluator(m, qobs, n_cont, inp["n_reg"], inp["rain"], inp["pos_evento"], row=0, members_per_launch=16)
    pop, fit, hist = nsga2.calib_nsga2(ev, el, pop_size=16, ngen=3, seed=0)
    assert pop.shape == (16, 11) and np.isfinite(fit).all()
    assert hist[-1][:, 0].max() >= hist[0][:, 0].max() and hist[-1][:, 0].max() > 0.5
