"""The reference's own `wmf.py`, run unmodified, against the drop-in boundary.

`tests/golden/make_simubasin_trace.py` imports `/root/reference/wmf/wmf.py` as it is, puts recording proxies where the f2py
modules `cu` and `models` stood (backed by the CPU oracle = the reference's Fortran, bit for bit) and drives the modelling
workflow: `SimuBasin(...)` (wmf.py:2989-3093), `set_Geomorphology` (:3625-3710), `set_PhysicVariables`, `set_Storage`,
`set_Control`, `rain_interpolate_idw` (:3369-3430), `run_shia` (:4337-4641).  The trace holds every module-variable
assignment (including the in-place edits of `models.h_coef[...]`), every call with the positional arguments and dims wmf.py
passed, and what came back.  The reference tree does not exist on the GPU box, so here the trace is REPLAYED against
`wmf_b200.cu` / `wmf_b200.models`: same calls, same argument objects (float64 maps where wmf.py builds float64 maps, `None`
for the eighteenth path, ...), results compared call by call -- Path A bit-exact, the hydrograph within 1e-5.
"""
import gzip
import os
import pickle

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
TRACE = os.path.join(HERE, "golden", "simubasin_trace.pkl.gz")
RESULT_VARS = {"mean_rain", "acum_rain"}          # module variables shia_v1 leaves behind (read back at wmf.py:4548-4601)


def _load():
    with gzip.open(TRACE, "rb") as f:
        return pickle.load(f)


def test_trace_fixture_is_complete():
    """CPU: the committed trace covers the workflow's calls."""
    d = _load()
    calls = {e[1] + "." + e[2] for e in d["trace"] if e[0] == "call"}
    for need in ("cu.basin_find", "cu.basin_cut", "cu.basin_acum", "cu.basin_basics", "cu.basin_subbasin_nod", "cu.basin_stream_type",
                 "cu.basin_stream_point2stream", "models.rain_idw", "models.shia_v1"):
        assert need in calls, need
    assert "rain.hdr" in d["files"] and "rain.bin" in d["files"]


@pytest.mark.skipif(not os.path.exists("/root/reference/wmf/wmf.py"), reason="reference tree not present")
def test_trace_fixture_is_current(tmp_path):
    """CPU, where the reference tree exists: running wmf.py again gives the committed trace."""
    import subprocess
    import sys
    out = tmp_path / "trace.pkl.gz"
    subprocess.run([sys.executable, "-W", "ignore", os.path.join(HERE, "golden", "make_simubasin_trace.py"), str(out)], check=True,
                   capture_output=True, timeout=600)
    with gzip.open(out, "rb") as f:
        new = pickle.load(f)
    old = _load()
    assert len(new["trace"]) == len(old["trace"])
    for a, b in zip(new["trace"], old["trace"]):
        assert a[:3] == b[:3]
        if a[0] == "call" and a[1] == "cu":
            for x, y in zip(_flat(a[4]), _flat(b[4])):
                np.testing.assert_array_equal(np.asarray(x), np.asarray(y))


def _flat(v):
    return list(v) if isinstance(v, tuple) else [v]


def _close(got, want, name, rtol=1e-5, afloor=1e-7):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    assert got.shape == want.shape, (name, got.shape, want.shape)
    tol = rtol * np.abs(want) + afloor * max(float(np.abs(want).max()), 1e-30)
    bad = np.abs(got - want) > tol
    assert not bad.any(), f"{name}: {bad.sum()} of {bad.size} off, worst {np.max(np.abs(got - want) / tol):.3g} x tol"


@pytest.mark.gpu
def test_replay_the_reference_workflow_on_the_shim(tmp_path):
    from wmf_b200 import CuModule, ModelsModule
    d = _load()
    mods = {"cu": CuModule(), "models": ModelsModule()}
    for name, blob in d["files"].items():
        if name.endswith(".hdr"):                         # written by wmf.py's Python, not by a Fortran routine
            (tmp_path / name).write_bytes(blob)
    rain_dir = d["rain_dir"]

    def arg(a):
        if isinstance(a, str) and a.startswith(rain_dir):
            return str(tmp_path / os.path.basename(a))
        return a

    n_calls = 0
    for ev in d["trace"]:
        if ev[0] == "set":
            _, mod, name, value = ev
            if mod == "models" and name in RESULT_VARS:
                _close(getattr(mods[mod], name), value, f"models.{name}", rtol=2e-3 if name == "mean_rain" else 1e-5)
                continue
            setattr(mods[mod], name, value)
            continue
        _, mod, name, args, want = ev
        got = getattr(mods[mod], name)(*[arg(a) for a in args])
        n_calls += 1
        label = f"{mod}.{name} (call {n_calls})"
        if mod == "cu":                                   # Path A and its neighbours: identical, NaN for NaN
            for k, (g, w) in enumerate(zip(_flat(got), _flat(want))):
                np.testing.assert_array_equal(np.asarray(g).reshape(-1), np.asarray(w).reshape(-1), err_msg=f"{label}[{k}]")
            assert len(_flat(got)) == len(_flat(want)), label
        elif name == "rain_idw":
            np.testing.assert_array_equal(np.asarray(got[1]).reshape(-1), np.asarray(want[1]).reshape(-1), err_msg=label + " posIds")
            _close(got[0], want[0], label + " meanRain", rtol=1e-5, afloor=1e-6)
            assert (tmp_path / "rain.bin").read_bytes() == d["files"]["rain.bin"], "rain_idw wrote a different record file"
        elif name == "shia_v1":
            names = ("Qsim", "Qsed", "Qseparated", "Hum", "St1", "St3", "Balance", "Speed", "Area", "StoOut", "Qsep_byrain")
            assert len(got) == len(want) == 11
            for nm, g, w in zip(names, got, want):
                if nm == "Balance":                      # a cancellation residue (modelosv2.f90:846): compared against the storage scale
                    assert np.abs(np.asarray(g) - np.asarray(w)).max() <= 1e-5 * float(np.abs(want[9]).sum())
                else:
                    _close(g, w, f"{label} {nm}")
            assert float(np.asarray(want[0])[0].max()) > 0.1
        else:
            raise AssertionError(f"unexpected call in the trace: {label}")
    assert n_calls >= 20
