#!/usr/bin/env python
"""Regenerate tests/golden/*.npz: outputs of THE REFERENCE ITSELF on small seeded synthetic cases.

Every ``o_*`` array is produced by the reference's own compiled Fortran (``oracle/_ref/libwmf_ref.so`` = the COFF objects
of the reference's last f2py build, loaded by ``oracle/ref_loader.c``, driven by ``oracle/ref.py``) and, at generation
time, checked to be bit-identical to the C restatement in ``oracle/``.  The one exception is ``o_balance64``, the fp64
twin of the mass balance that only the restatement can produce (the reference sums in fp32).  The script needs
``/root/reference`` (or a prebuilt ``oracle/_ref``); the committed vectors let the tests run without either.

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
    "shia_separate_fluxes": ("G2", 44, 40, 14, 40, dict(rain_peak_mm=50.0), dict(separate_fluxes=1)),
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


def run_reference(inp, flags):
    from oracle import ref
    d = dict(inp)
    d.update(flags)
    return ref.shia_v1(d)


KEEP = ("q", "stoout", "hum", "st1", "st3", "balance", "balance64", "speed", "areacontrol", "mean_rain", "acum_rain", "qsed", "qseparated", "vs", "vd",
        "vsc", "vdc", "volero", "voldepo", "retorned", "sl_riskvector", "sl_slideocurrence", "sl_slidencelltime")


def build_topo_case():
    """Small basin + gauges for the sub-basin / rain-interpolation fixture (inputs only)."""
    from conftest import basin_vectors, make_basin
    from wmf_b200 import synth
    bs = make_basin("G2", 60, 48, 21)
    v = basin_vectors(bs)
    x, y = bs["cu"].basin_coordxy(bs["basin"])
    xy = np.asfortranarray(np.vstack([x, y]), dtype=np.float32)
    coord, rain = synth.make_stations(60, 48, bs["dx"], 7, 24, 21)
    return bs, v, xy, coord, rain


def run_topo(cu, mod, bs, v, xy, coord, rain, thr=15):
    """The sub-basin chain and rain_idw (cells and hills) through ``cu`` (RefCu / oracle.Cu) and ``mod`` (oracle.ref / oracle)."""
    b, n = bs["basin"], bs["n"]
    cauce, nodos, nn = cu.basin_subbasin_nod(b, v["acum"], thr)
    sub_pert, _ = cu.basin_subbasin_find(b, nodos, nn)
    sb = cu.basin_subbasin_cut(nn)
    sh, nh = cu.basin_subbasin_horton(sb, n)
    sl, mx, nmx = cu.basin_subbasin_long(sub_pert, cauce, v["lng"], sb, sh)
    pslope, plong = cu.basin_subbasin_stream_prop(sub_pert, cauce, v["lng"], v["pend"], nn)
    c2, nod2, _, _, ncau = cu.basin_stream_nod(b, v["acum"], thr)
    ss, sll = cu.basin_stream_slope(b, v["elev"], v["lng"], nod2, ncau)
    idw = mod.rain_idw(xy, coord, rain, 2.0, 0.0)
    idwh = mod.rain_idw(xy, coord, rain, 2.0, 0.0, mask=sub_pert, nhills=nn)
    return {"t_cauce": cauce, "t_nodos_fin": nodos, "t_sub_pert": sub_pert, "t_sub_basins": sb, "t_sub_horton": sh,
            "t_nod_horton_nodes": nh[sb[0] - 1], "t_sub_basin_long": sl, "t_max_long": np.array([mx, nmx], np.float64),
            "t_hill_slope": pslope, "t_hill_long": plong, "t_stream_s": ss, "t_stream_l": sll,
            "r_table": idw["table"], "r_pos": idw["pos_ids"], "r_mean": idw["mean_rain"],
            "r_table_hills": idwh["table"], "thr": np.array([thr])}


def main():
    import oracle
    from oracle import ref as refmod
    bs, v, xy, coord, rain = build_topo_case()
    rcu = refmod.RefCu()
    o = bs["cu"]
    rcu.ncols, rcu.nrows, rcu.xll, rcu.yll, rcu.dx, rcu.dy, rcu.dxp, rcu.nodata = o.ncols, o.nrows, o.xll, o.yll, o.dx, o.dy, o.dxp, o.nodata
    R, O = run_topo(rcu, refmod, bs, v, xy, coord, rain), run_topo(o, oracle, bs, v, xy, coord, rain)
    for k in R:
        assert np.array_equal(np.asarray(R[k]), np.asarray(O[k]), equal_nan=True), ("topo_rain", k)
    R["source"] = np.array(["reference Fortran objects via oracle/_ref"])
    np.savez_compressed(os.path.join(HERE, "topo_rain.npz"), **R)
    print("topo_rain", {k: np.asarray(a).shape for k, a in R.items()})
    for name in CASES:
        bs, v, inp, flags = build_case(name)
        orc = run_oracle(inp, flags)
        ref = run_reference(inp, flags)
        from oracle import ref as refmod
        rcu = refmod.RefCu()
        o = bs["cu"]
        rcu.ncols, rcu.nrows, rcu.xll, rcu.yll, rcu.dx, rcu.dy, rcu.dxp, rcu.nodata = o.ncols, o.nrows, o.xll, o.yll, o.dx, o.dy, o.dxp, o.nodata
        n = rcu.basin_find(bs["x"], bs["y"], bs["DIR"])
        rb = rcu.basin_cut(n)
        racum, rlng, rpend, relev = rcu.basin_basics(rb, v["demv"], v["dirv"])
        for a, b in ((rb, bs["basin"]), (racum, v["acum"]), (rlng, v["lng"]), (rpend, v["pend"]), (relev, v["elev"])):
            assert np.array_equal(a, b), "Path A: reference and restatement differ"
        for k in KEEP:
            if k in ref and ref[k] is not None and k in orc:
                assert np.array_equal(np.asarray(ref[k]).reshape(-1), np.asarray(orc[k]).reshape(-1)), (name, k)
        ref["balance64"] = orc["balance64"]
        out = {"basin_f": bs["basin"], "acum": v["acum"], "long": v["lng"], "pend": v["pend"], "elev": v["elev"],
               "rain_checksum": np.array([int(inp["rain"].astype(np.int64).sum())]),
               "calib": np.asarray(inp["calib"], np.float32),
               "source": np.array(["reference Fortran objects (gfortran 5.3.0 -O3 -funroll-loops) via oracle/_ref; balance64 from oracle/"])}
        for k in KEEP:
            if k in ref and ref[k] is not None:
                out["o_" + k] = np.asarray(ref[k]).reshape(np.asarray(orc[k]).shape) if k in orc else np.asarray(ref[k])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, {k: v.shape for k, v in out.items() if k.startswith("o_")})


if __name__ == "__main__":
    main()
