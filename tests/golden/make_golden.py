#!/usr/bin/env python
"""Regenerate tests/golden/*.npz: outputs of the CPU oracle (oracle/, the C restatement of the reference's Fortran) on
small seeded synthetic cases.  The real reference cannot run in this image (no Fortran compiler), so these vectors do
NOT pin the oracle to the reference -- they freeze the oracle's output of the day they were made, so that later edits
of oracle/ or of the generators cannot drift silently, and give the GPU tests a fixture that does not need the oracle.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

CASES = {
    # name: (generator, ncols, nrows, seed, n_reg, synth kwargs, model flags)
    "shia_linear": ("G2", 40, 32, 11, 48, dict(n_control_h=2), dict(show_speed=1)),
    "shia_kinematic_returns": ("G1", 36, 40, 12, 40, dict(speed_type=(2, 2, 1), retorno_gr=1, retorno_aq=1), {}),
    "shia_sed_slides": ("G2", 44, 36, 13, 36, dict(speed_type=(2, 2, 1), sediments=True, slides=True, retorno_gr=1,
                                                    rain_peak_mm=60.0), {}),
}


def build_case(name):
    import oracle
    from conftest import basin_vectors, make_basin
    from wmf_b200 import synth
    kind, nc, nr, seed, n_reg, kw, flags = CASES[name]
    bs = make_basin(kind, nc, nr, seed)
    v = basin_vectors(bs)
    inp = synth.make_shia_inputs(bs["basin"], v["acum"], v["lng"], v["pend"], dxp=30.0, n_reg=n_reg, seed=seed, **kw)
    return bs, v, inp, flags


def run_oracle(inp, flags):
    import oracle
    d = dict(inp)
    d.update(flags)
    return oracle.shia_v1(d)


KEEP = ("q", "stoout", "hum", "st1", "st3", "balance64", "speed", "areacontrol", "mean_rain", "acum_rain", "qsed", "vs", "vd",
        "vsc", "vdc", "volero", "voldepo", "retorned", "sl_riskvector", "sl_slideocurrence", "sl_slidencelltime")


def main():
    for name in CASES:
        bs, v, inp, flags = build_case(name)
        ref = run_oracle(inp, flags)
        out = {"basin_f": bs["basin"], "acum": v["acum"], "long": v["lng"], "pend": v["pend"], "elev": v["elev"],
               "rain_checksum": np.array([int(inp["rain"].astype(np.int64).sum())]),
               "calib": np.asarray(inp["calib"], np.float32)}
        for k in KEEP:
            if k in ref and ref[k] is not None:
                out["o_" + k] = np.asarray(ref[k])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, {k: v.shape for k, v in out.items() if k.startswith("o_")})


if __name__ == "__main__":
    main()
