#!/usr/bin/env python
"""Stage-by-stage comparison of the CUDA tile path of basin_find (csrc/pathfind.cu, dumped with WMF_PF_DUMP) with the CPU
model of the same decomposition (oracle/tile_model.py, tile edge 64) and with the oracle's queue.  Development tool."""
import argparse
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", default="partial", choices=["partial", "G1", "G2"])
    ap.add_argument("--nc", type=int, default=150)
    ap.add_argument("--nr", type=int, default=130)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    import oracle
    from oracle import tile_model as tm
    from wmf_b200 import CuModule, synth
    nc, nr, dx = a.nc, a.nr, 30.0
    if a.kind == "partial":
        DIR, (oc_, orow) = synth.make_partial_raster(nc, nr, a.seed, noise=1.6)
    else:
        DIR, _ = synth.make_raster(a.kind, nc, nr, a.seed)
        oc_, orow = synth.outlet_colrow(nc, nr)
    x, y = dx * (oc_ - 0.5), dx * (nr - orow + 0.5)
    oc, cu = oracle.Cu(), CuModule()
    for c in (oc, cu):
        c.ncols, c.nrows, c.xll, c.yll, c.dx, c.dy, c.dxp, c.nodata = nc, nr, 0.0, 0.0, dx, dx, dx, synth.NODATA
    n_o = oc.basin_find(x, y, DIR)
    b_o = oc.basin_cut(n_o)
    a_o = np.asarray(oc.basin_acum(b_o)).reshape(-1)
    ql, qp, qa, st = tm.tile_find(DIR, oc_ - 1, orow - 1, 64)
    assert np.array_equal(tm.basin_table(ql, qp, nc), b_o), "the CPU model itself differs from the oracle"
    d = tempfile.mkdtemp(prefix="pf_")
    os.environ["WMF_PF_DUMP"] = d
    err = None
    try:
        n = cu.basin_find(x, y, DIR, nc, nr)
    except Exception as e:                                  # the dumps up to the failing stage are still there
        err, n = e, -1
    print(f"{a.kind} {nc}x{nr} seed {a.seed}: oracle n = {n_o}, gpu n = {n}", "" if err is None else f"(gpu error: {err})")

    def load(name, dt):
        p = os.path.join(d, f"pf_{name}.bin")
        return np.fromfile(p, dtype=dt) if os.path.exists(p) else None

    def cmp(name, got, want, mask=None):
        if got is None:
            print(f"  {name:8s} not dumped")
            return
        got, want = np.asarray(got).reshape(-1), np.asarray(want).reshape(-1)
        if mask is not None:
            m = np.asarray(mask).reshape(-1)
            bad = np.nonzero(m & (got != want))[0]
        else:
            bad = np.nonzero(got != want)[0]
        print(f"  {name:8s} {'ok' if bad.size == 0 else f'{bad.size} differ, first at {bad[:5]}: got {got[bad[:5]]} want {want[bad[:5]]}'}")

    nn = st["pterm"].size
    is_perim = np.zeros(nn, bool)
    ntx = (nc + 63) // 64
    for r in range(nr):
        for c in range(nc):
            if r % 64 in (0, 63) or c % 64 in (0, 63):
                is_perim[((r // 64) * ntx + c // 64) * 256 + tm.perim_index(r % 64, c % 64, 64)] = True
    pt = st["pterm"].copy()
    pt[pt >= 0] %= 256
    is_exit = st["bdown"] >= 0
    cmp("pterm", load("pterm", np.int32), pt, is_perim)
    cmp("bdown", load("bdown", np.int32), st["bdown"], is_perim)
    cmp("lsize", load("lsize", np.uint32), st["lsize"], is_exit)
    cmp("bnext", load("bnext", np.int32), st["bnext"], is_exit)
    w = load("w", np.uint64)
    live_exit = is_exit & st["alive"]
    cmp("G", None if w is None else (w & 0xFFFFFFFF), st["G"], live_exit)
    S = load("S", np.uint32)
    basin = np.zeros(nr * nc, bool)
    basin[ql] = True
    cmp("S", S, st["S"], basin)
    cmp("boff", load("boff", np.uint32), st["boff"], live_exit)
    eho = load("eho", np.uint32)
    if eho is not None:
        eho = eho.reshape(-1, 2)
        pb = np.zeros(nn, bool)
        for l in ql:
            r, c = divmod(int(l), nc)
            if r % 64 in (0, 63) or c % 64 in (0, 63):
                pb[((r // 64) * ntx + c // 64) * 256 + tm.perim_index(r % 64, c % 64, 64)] = True
        cmp("eh", eho[:, 0], st["eh"], pb)
        cmp("eo", eho[:, 1], st["eo"], pb)
    LR = load("LR", np.uint32)
    if LR is not None:
        LR = LR.reshape(-1, 2)
        code = np.full(nr * nc, 0xFFFF, np.int64)
        hop = np.zeros(nr * nc, np.int64)
        osum = np.zeros(nr * nc, np.int64)
        nchild = st["nchild"].reshape(-1)
        for (r, c), (cd, h, o) in st["LT"].items():
            code[r * nc + c] = 0xFFFE if cd == tm.ROOT_OUT else (0xFFFF if cd == tm.DEAD else cd % 256)
            hop[r * nc + c], osum[r * nc + c] = h, o
        cmp("LR.tc", LR[:, 0] & 0xFFFF, code, basin)
        cmp("LR.h", (LR[:, 0] >> 16) & 0xFFF, hop, basin)
        cmp("LR.nch", LR[:, 0] >> 28, nchild, basin)
        cmp("LR.o", LR[:, 1], osum, basin)
    nx, dp = load("nx", np.int32), load("dp", np.uint32)
    if nx is not None:
        cmp("alive", nx == -1, st["alive"], is_exit)
        dp = dp.reshape(-1, 2)
        cmp("Dn", dp[:, 0], st["Dn"], live_exit)
        cmp("Pn", dp[:, 1], st["Pn"], live_exit)
    it0 = load("items0", np.uint32)                 # interleaved (key, value) pairs at position preorder
    k0, v0 = (it0[0::2], it0[1::2]) if it0 is not None else (None, None)
    if k0 is not None and k0.size == n_o:
        # position = preorder: depth of the cell stored there, raster index
        dep = {int(l): int(dd) for dd, l in zip(_depths(qp), ql)}
        want_v = np.zeros(n_o, np.int64)
        want_k = np.zeros(n_o, np.int64)
        pre = _preorder(ql, qp, st["nchild"].reshape(-1))
        want_v[pre] = ql
        want_k[pre] = [dep[int(l)] for l in ql]
        cmp("vals0", v0, want_v)
        cmp("depth0", k0 & 0x0FFFFFFF, want_k)
    if err is None:
        b = cu.basin_cut(n)
        print("  basin_f", "ok" if np.array_equal(b, b_o) else "DIFFERS")
        ac = cu.basin_acum(b)
        print("  acum   ", "ok" if np.array_equal(np.asarray(ac).reshape(-1), a_o) else "DIFFERS")
        os.environ["WMF_ACUM_ALGO"] = "climb"
        print("  acum (climb)", "ok" if np.array_equal(np.asarray(cu.basin_acum(b)).reshape(-1), a_o) else "DIFFERS")
        del os.environ["WMF_ACUM_ALGO"]
    print("dumps in", d)


def _depths(q_par):
    d = np.zeros(q_par.size, np.int64)
    for j in range(1, q_par.size):
        d[j] = d[q_par[j] - 1] + 1
    return d


def _preorder(q_lin, q_par, nchild_flat):
    """preorder index of every queue position (children of a cell are consecutive in the queue, in probe order)."""
    n = q_lin.size
    first = np.ones(n, np.int64)
    cs = 1
    for j in range(n):
        first[j] = cs
        cs += int(nchild_flat[q_lin[j]])
    size = np.ones(n, np.int64)
    for j in range(n - 1, 0, -1):
        size[q_par[j] - 1] += size[j]
    pre = np.zeros(n, np.int64)
    for j in range(n):
        run = pre[j] + 1
        for c in range(first[j], first[j] + int(nchild_flat[q_lin[j]])):
            pre[c] = run
            run += size[c]
    return pre


if __name__ == "__main__":
    main()
