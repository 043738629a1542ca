#!/usr/bin/env python
"""Path A at scale: basin_find + basin_cut + basin_acum (+ basics / HAND) on synthetic G2 rasters, DIR resident in HBM.
Prints one JSON line per size: Mcell/s, levels, per-level time, achieved GB/s against the 28 B/cell algorithmic figure."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1024,4096,8192")
    ap.add_argument("--kind", default="G2")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--check", action="store_true", help="verify structural invariants on the host (slow for large N)")
    ap.add_argument("--extras", action="store_true", help="also time basin_map2basin, basin_basics, stream mask and geo_hand (resident)")
    ap.add_argument("--cpu-ref", action="store_true", help="run the reference's own compiled Fortran (oracle/_ref) on the same "
                    "raster: time find+cut+acum on one host core and require the GPU tables to be bit-identical to it")
    ap.add_argument("--cpu-ref-max-cells", type=int, default=70_000_000)
    a = ap.parse_args()
    import torch
    from wmf_b200 import Context, CuModule, synth
    ctx = Context(0)
    cu = CuModule(ctx)
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    for s in a.sizes.split(","):
        nc = nr = int(s)
        t0 = time.time()
        DIR, DEM = synth.make_raster(a.kind, nc, nr, seed=0)
        gen_s = time.time() - t0
        d_dem = torch.from_numpy(np.ascontiguousarray(DEM.T)).cuda() if a.extras else None
        del DEM
        cu.ncols, cu.nrows, cu.xll, cu.yll, cu.dx, cu.dy, cu.dxp, cu.nodata = nc, nr, 0.0, 0.0, 30.0, 30.0, 30.0, synth.NODATA
        x, y = synth.outlet_xy(nc, nr, 0.0, 0.0, 30.0)
        d_dir = torch.from_numpy(np.ascontiguousarray(DIR.T)).cuda()
        cpu = None
        if a.cpu_ref and nc * nr <= a.cpu_ref_max_cells:
            from oracle import ref
            rc = ref.RefCu()
            rc.ncols, rc.nrows, rc.xll, rc.yll, rc.dx, rc.dy, rc.dxp, rc.nodata = nc, nr, 0.0, 0.0, 30.0, 30.0, 30.0, synth.NODATA
            t0 = time.perf_counter()
            rn = rc.basin_find(x, y, DIR)
            t1 = time.perf_counter()
            rb = rc.basin_cut(rn)
            t2 = time.perf_counter()
            racum = rc.basin_acum(rb)
            t3 = time.perf_counter()
            cpu = dict(n=rn, b=rb, acum=racum, find_ms=1e3 * (t1 - t0), cut_ms=1e3 * (t2 - t1), acum_ms=1e3 * (t3 - t2))
        del DIR
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        best = None
        for _ in range(a.reps):
            torch.cuda.synchronize()
            ev[0].record()
            n = cu.basin_find(x, y, d_dir)
            ev[1].record()
            b = cu.basin_cut(n, resident=True)
            ev[2].record()
            acum = cu.basin_acum(b)
            ev[3].record()
            torch.cuda.synchronize()
            ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
            if best is None or sum(ts) < sum(best):
                best = ts
        levels = cu.basin_find_depth()
        total = sum(best)
        assert int(acum[-1].item()) == n, (int(acum[-1].item()), n)
        line = {"raster": f"{nc}x{nr} {a.kind}", "cells": int(n), "levels": int(levels), "find_ms": best[0], "cut_ms": best[1],
                "acum_ms": best[2], "total_ms": total, "mcell_per_s": n / total / 1e3,
                "us_per_level_find": 1e3 * best[0] / max(levels, 1),
                "achieved_gbs_28B": 28.0 * n / (total / 1e3) / 1e9, "frac_of_measured_hbm": 28.0 * n / (total / 1e3) / 1e9 / peak,
                "gen_s": gen_s}
        if cpu is not None:
            assert cpu["n"] == n
            same_b = bool((b.cpu().numpy().reshape(-1) == cpu["b"].reshape(-1, order="F")).all())
            same_a = bool((acum.cpu().numpy().reshape(-1) == cpu["acum"]).all())
            tot = cpu["find_ms"] + cpu["cut_ms"] + cpu["acum_ms"]
            line.update({"cpu_reference": {"kind": "reference", "cores": 1, "find_ms": cpu["find_ms"], "cut_ms": cpu["cut_ms"],
                                           "acum_ms": cpu["acum_ms"], "mcell_per_s": n / tot / 1e3,
                                           "what": "the reference's own compiled Fortran (oracle/_ref), same raster"},
                         "bit_exact_vs_reference": same_b and same_a, "speedup_vs_reference": tot / total})
            assert same_b and same_a, "GPU tables differ from the reference"
            cpu = None
        if a.extras:
            def timed(fn):
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                best_t, res = None, None
                for _ in range(2):
                    e0.record()
                    res = fn()
                    e1.record()
                    torch.cuda.synchronize()
                    best_t = e0.elapsed_time(e1) if best_t is None else min(best_t, e0.elapsed_time(e1))
                return best_t, res
            d_dirf = d_dir.float()
            t_m2b, demv = timed(lambda: cu.basin_map2basin(b, d_dem, 0.0, 0.0, 30.0, 30.0, synth.NODATA))
            _, dirv = timed(lambda: cu.basin_map2basin(b, d_dirf, 0.0, 0.0, 30.0, 30.0, synth.NODATA))
            dirv = dirv.int()
            t_bas, (ac2, lng, pend, elev) = timed(lambda: cu.basin_basics(b, demv, dirv))
            assert bool((ac2 == acum).all())
            t_typ, _ = timed(lambda: cu.basin_stream_type(b, acum, [30, 500]))
            cauce = (acum >= 500).int()
            t_hand, (hand, hdnd, aq) = timed(lambda: cu.geo_hand(b, elev, lng, cauce))
            assert float(hand.min()) >= 0.0
            # raster-wide HAND (geo_hand_global, cuencas.f90:2175-2219): accumulation back on the raster, network mask
            # acum > 1000 cells, every other cell walks down DIR to the network
            t_v2m, acum_map = timed(lambda: cu.basin_float_var2map(b, acum.float(), nc, nr))
            t_net, red = timed(lambda: cu.geo_acum_to_cauce(acum_map.int(), 1000))
            t_hg, hand_g = timed(lambda: cu.geo_hand_global(d_dem, d_dir, red))
            line.update({"var2map_ms": t_v2m, "acum_to_cauce_ms": t_net, "geo_hand_global_ms": t_hg,
                         "geo_hand_global_gbs_16B": 16.0 * nc * nr / (t_hg / 1e3) / 1e9})
            del acum_map, red, hand_g
            line.update({"map2basin_ms": t_m2b, "basics_ms": t_bas, "stream_type_ms": t_typ, "geo_hand_ms": t_hand,
                         "basics_gbs_40B": 40.0 * n / (t_bas / 1e3) / 1e9, "geo_hand_gbs_40B": 40.0 * n / (t_hand / 1e3) / 1e9})
            # sub-basin (hills) topology as SimuBasin.__init__ / set_Geomorphology drive it (wmf.py:3019-3023, 3664-3676)
            thr = 500
            t_nod, (cau2, nodos_fin, nn) = timed(lambda: cu.basin_subbasin_nod(b, acum, thr))
            t_find, (hills_own, _) = timed(lambda: cu.basin_subbasin_find(b, nodos_fin, nn))
            hills = cu.basin_subbasin_cut(nn)
            cu.basin_subbasin_nod(b, acum, thr)
            t_hort, (sub_h, _) = timed(lambda: cu.basin_subbasin_horton(hills, n, nn))
            ho, ca, lg, pe = (x.cpu().numpy() for x in (hills_own, cau2, lng, pend))
            t_long, _ = timed(lambda: cu.basin_subbasin_long(ho, ca, lg, hills, sub_h, nn, n))
            t_prop, _ = timed(lambda: cu.basin_subbasin_stream_prop(ho, ca, lg, pe, nn, n))
            ca3, nod3, _, _, ncau = cu.basin_stream_nod(b, acum, thr)
            t_ssl, _ = timed(lambda: cu.basin_stream_slope(b, elev, lng, nod3, ncau))
            line.update({"subbasin": {"threshold": thr, "n_nodos": int(nn), "nod_ms": t_nod, "find_ms": t_find, "horton_ms": t_hort,
                                      "long_ms_host_arrays": t_long, "stream_prop_ms_host_arrays": t_prop, "stream_slope_ms": t_ssl}})
            del d_dem, d_dirf, demv, dirv, ac2, lng, pend, elev, cauce, hand, hdnd, aq
        if a.check:
            bb = b.cpu().numpy().reshape(-1, 3)
            d = bb[:, 0]
            assert d[-1] == 0 and (d[:-1] > 0).all()
            par = n - d[:-1]
            assert (par > np.arange(n - 1)).all()
            line["checked"] = True
        print(json.dumps(line), flush=True)
        del d_dir, b, acum
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
