"""Seeded synthetic inputs for both hot paths.

Everything here is *input generation* shared by tests and bench.py: D8 direction /
elevation rasters that are valid by construction (acyclic, single outlet, 1-cell nodata
frame -- the frame avoids the reference's out-of-bounds read when a basin fills the
whole raster, cuencas.f90:584-586), and SHIA/TETIS parameter maps, rainfall and control
points in the ranges used by the reference notebooks.

Rasters are returned in the wmf.py convention: shape ``(ncols, nrows)`` (the transposed
GDAL array, wmf.py:341-348), Fortran-contiguous, so that element (col,row) sits at
``row*ncols + col`` -- exactly what f2py hands to the Fortran.
"""
from __future__ import annotations

import numpy as np

NODATA = -9999.0

# keypad D8 codes: 7 8 9 / 4 5 6 / 1 2 3 ; move (mx,my) -> 5 - 3*my + mx  (cuencas.f90:445,2684-2699)
_MOVES = [(-1, -1), (0, -1), (1, -1), (-1, 0), (1, 0), (-1, 1), (0, 1), (1, 1)]


def hash32(seed: int, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Vectorised 32-bit integer hash of (seed, a, b) -> uint32 (murmur3 finaliser mix)."""
    with np.errstate(over="ignore"):
        h = (np.asarray(a, dtype=np.uint64) * np.uint64(0x9E3779B1)
             + np.asarray(b, dtype=np.uint64) * np.uint64(0x85EBCA77)
             + np.uint64((int(seed) * 0xC2B2AE3D + 0x27D4EB2F) & 0xFFFFFFFF)) & np.uint64(0xFFFFFFFF)
        h ^= h >> np.uint64(16)
        h = (h * np.uint64(0x85EBCA6B)) & np.uint64(0xFFFFFFFF)
        h ^= h >> np.uint64(13)
        h = (h * np.uint64(0xC2B2AE35)) & np.uint64(0xFFFFFFFF)
        h ^= h >> np.uint64(16)
    return h.astype(np.uint32)


def u01(seed: int, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    return (hash32(seed, a, b).astype(np.float64) / 4294967296.0).astype(np.float32)


def outlet_colrow(nc: int, nr: int):
    """1-based (col,row) of the synthetic outlet: bottom-centre interior cell."""
    return (nc + 1) // 2, nr - 1


def outlet_xy(nc: int, nr: int, xll: float, yll: float, dx: float):
    """Coordinates of the outlet cell centre, to be fed to basin_find (coord2fil_col, cuencas.f90:77-90)."""
    oc, orow = outlet_colrow(nc, nr)
    return xll + dx * (oc - 0.5), yll + dx * (nr - orow + 0.5)


def make_raster(kind: str, nc: int, nr: int, seed: int = 0, chunk_rows: int = 2048):
    """Return (DIR int32 (nc,nr), DEM float32 (nc,nr)) for generator ``kind`` in {"G1","G2"}.

    G1 "curtain": every interior cell above the bottom interior row drains down-left / down /
    down-right (hash-chosen, clipped at the side walls); bottom-row cells drain sideways to the
    outlet column.  G2 "radial dendritic": every cell drains to a hash-chosen neighbour whose
    Chebyshev distance to the outlet is strictly smaller.  Both: basin == all interior cells.
    """
    assert nc >= 5 and nr >= 4
    oc, orow = outlet_colrow(nc, nr)
    dir_rc = np.full((nr, nc), int(NODATA), dtype=np.int32)   # [row, col]
    dem_rc = np.full((nr, nc), NODATA, dtype=np.float32)
    cols = np.arange(1, nc + 1, dtype=np.int64)[None, :]
    for r0 in range(1, nr + 1, chunk_rows):
        r1 = min(nr + 1, r0 + chunk_rows)
        rows = np.arange(r0, r1, dtype=np.int64)[:, None]
        interior = (cols >= 2) & (cols <= nc - 1) & (rows >= 2) & (rows <= nr - 1)
        h = hash32(seed, np.broadcast_to(cols, (r1 - r0, nc)), np.broadcast_to(rows, (r1 - r0, nc)))
        noise = (hash32(seed + 7919, np.broadcast_to(cols, (r1 - r0, nc)),
                        np.broadcast_to(rows, (r1 - r0, nc))).astype(np.float64) / 4294967296.0)
        if kind == "G1":
            choice = (h % np.uint32(3)).astype(np.int32) + 1            # 1,2,3
            choice = np.where((cols == 2) & (choice == 1), 2, choice)
            choice = np.where((cols == nc - 1) & (choice == 3), 2, choice)
            bottom = rows == nr - 1
            side = np.where(cols < oc, 6, np.where(cols > oc, 4, 2)).astype(np.int32)
            d = np.where(bottom, side, choice).astype(np.int32)
            pot = 2 * (nr - 1 - rows) + np.abs(cols - oc)
        elif kind == "G2":
            dc = cols - oc
            dr = rows - orow
            cheb = np.maximum(np.abs(dc), np.abs(dr))
            valid = []
            for mx, my in _MOVES:
                c2, r2 = cols + mx, rows + my
                inside = (c2 >= 2) & (c2 <= nc - 1) & (r2 >= 2) & (r2 <= nr - 1)
                ch2 = np.maximum(np.abs(dc + mx), np.abs(dr + my))
                valid.append(inside & (ch2 < cheb))
            cnt = np.zeros(cheb.shape, dtype=np.int32)
            for v in valid:
                cnt += v
            pick = (h % np.maximum(cnt, 1).astype(np.uint32)).astype(np.int32)
            d = np.full(cheb.shape, 2, dtype=np.int32)                   # outlet keeps 2
            run = np.zeros(cheb.shape, dtype=np.int32)
            for (mx, my), v in zip(_MOVES, valid):
                sel = v & (run == pick)
                d = np.where(sel, 5 - 3 * my + mx, d).astype(np.int32)
                run += v
            pot = cheb
        else:
            raise ValueError(kind)
        dir_rc[r0 - 1:r1 - 1] = np.where(interior, d, int(NODATA))
        dem = (10.0 + 0.5 * pot + 0.25 * noise).astype(np.float32)
        dem_rc[r0 - 1:r1 - 1] = np.where(interior, dem, np.float32(NODATA))
    return dir_rc.T, dem_rc.T   # (ncols, nrows), F-contiguous views


def make_dir_torch(kind: str, nc: int, nr: int, seed: int = 0, device="cuda", chunk_rows: int = 4096):
    """The DIR raster of ``make_raster`` built with torch ops on ``device`` (bit-identical; CPU-side generation of a
    20000 x 20000 raster takes minutes).  Returns an int32 tensor of shape (nrows, ncols) = the resident layout of a
    ``(ncols, nrows)`` f2py raster (element (col,row) at ``row*ncols + col``)."""
    import torch
    assert nc >= 5 and nr >= 4 and kind in ("G1", "G2")
    oc, orow = outlet_colrow(nc, nr)
    M32 = 0xFFFFFFFF

    def h32(sd, a, b):
        # int64 arithmetic wraps modulo 2^64, so the low 32 bits of every product are those of the uint64 version
        h = (a * 0x9E3779B1 + b * 0x85EBCA77 + ((int(sd) * 0xC2B2AE3D + 0x27D4EB2F) & M32)) & M32
        h = h ^ (h >> 16)
        h = (h * 0x85EBCA6B) & M32
        h = h ^ (h >> 13)
        h = (h * 0xC2B2AE35) & M32
        return h ^ (h >> 16)

    out = torch.empty((nr, nc), dtype=torch.int32, device=device)
    cols = torch.arange(1, nc + 1, dtype=torch.int64, device=device)[None, :]
    nod = int(NODATA)
    for r0 in range(1, nr + 1, chunk_rows):
        r1 = min(nr + 1, r0 + chunk_rows)
        rows = torch.arange(r0, r1, dtype=torch.int64, device=device)[:, None]
        interior = (cols >= 2) & (cols <= nc - 1) & (rows >= 2) & (rows <= nr - 1)
        h = h32(seed, cols.expand(r1 - r0, nc), rows.expand(r1 - r0, nc))
        if kind == "G1":
            choice = h % 3 + 1
            choice = torch.where((cols == 2) & (choice == 1), 2, choice)
            choice = torch.where((cols == nc - 1) & (choice == 3), 2, choice)
            side = torch.where(cols < oc, 6, torch.where(cols > oc, 4, 2)).expand(r1 - r0, nc)
            d = torch.where((rows == nr - 1).expand(r1 - r0, nc), side, choice)
        else:
            dc, dr = cols - oc, rows - orow
            cheb = torch.maximum(dc.abs(), dr.abs())
            valid = []
            for mx, my in _MOVES:
                c2, r2 = cols + mx, rows + my
                inside = (c2 >= 2) & (c2 <= nc - 1) & (r2 >= 2) & (r2 <= nr - 1)
                valid.append(inside & (torch.maximum((dc + mx).abs(), (dr + my).abs()) < cheb))
            cnt = torch.zeros_like(cheb)
            for v in valid:
                cnt = cnt + v
            pick = h % cnt.clamp(min=1)
            d = torch.full_like(cheb, 2)
            run = torch.zeros_like(cheb)
            for (mx, my), v in zip(_MOVES, valid):
                d = torch.where(v & (run == pick), 5 - 3 * my + mx, d)
                run = run + v
        out[r0 - 1:r1 - 1] = torch.where(interior, d, nod).to(torch.int32)
    return out


def make_partial_raster(nc: int, nr: int, seed: int = 0, noise: float = 1.8, p_hole: float = 0.03):
    """Random D8 raster whose outlet catches only PART of the map: (DIR int32 (nc,nr), outlet (col,row) 1-based).

    Every cell gets a pseudo-elevation = distance to the outlet + ``noise`` * uniform; it drains to a hash-chosen
    neighbour that is strictly lower (so the map is acyclic), local minima and ``p_hole`` of the cells are nodata.  Cells
    whose path ends in another pit or a hole are outside the basin; nothing is guaranteed about the number of cells
    draining into one cell (up to 8).  No nodata frame is needed because the basin never fills the raster."""
    rng = np.random.default_rng(seed)
    oc, orow = (nc + 1) // 2, (2 * nr) // 3
    cols, rows = np.meshgrid(np.arange(1, nc + 1), np.arange(1, nr + 1), indexing="ij")
    h = np.hypot(cols - oc, rows - orow) + noise * rng.random((nc, nr))
    h[oc - 1, orow - 1] = -1.0
    hole = rng.random((nc, nr)) < p_hole
    hole[oc - 1, orow - 1] = False
    h = np.where(hole, np.inf, h)
    pad = np.full((nc + 2, nr + 2), np.inf)
    pad[1:-1, 1:-1] = h
    lower, codes = [], []
    for mx, my in _MOVES:
        nb = pad[1 + mx: 1 + mx + nc, 1 + my: 1 + my + nr]
        lower.append(nb < h)
        codes.append(5 - 3 * my + mx)
    lower = np.stack(lower)                                  # (8, nc, nr)
    cnt = lower.sum(0)
    pick = (rng.integers(0, 1 << 30, (nc, nr)) % np.maximum(cnt, 1)).astype(np.int64)
    order = np.cumsum(lower, 0) - 1                          # rank of each admissible move
    sel = lower & (order == pick[None])
    DIR = np.full((nc, nr), int(NODATA), np.int32)
    for k, code in enumerate(codes):
        DIR = np.where(sel[k], code, DIR).astype(np.int32)
    DIR = np.where(hole | (cnt == 0), int(NODATA), DIR).astype(np.int32)
    return np.asfortranarray(DIR), (oc, orow)


def loguniform(rng, lo, hi, n):
    return np.exp(rng.uniform(np.log(lo), np.log(hi), n))


DEFAULT_CALIB = np.array([0.000325, 2, 2, 0, 0.2, 0.00012, 0.01, 0.005, 1, 1, 1], dtype=np.float32)


def make_shia_inputs(structure, acum, hill_long, slope, *, dxp, dt=3600.0, n_reg=100, seed=0,
                     thresholds=(30, 500), speed_type=(1, 1, 1), n_control=16,
                     zero_frac=0.6, rain_peak_mm=25.0, sediments=False, slides=False,
                     retorno_gr=0, retorno_aq=0, n_control_h=0) -> dict:
    """Synthetic SHIA/TETIS inputs for a basin table ``structure`` (3,N) (SURVEY 8d ranges).

    Returns a dict keyed by the lower-case ``models`` attribute names plus ``calib``, ``rain``
    (n_records,N) int32 mm*1000 with record 1 == zeros (modelosv2.f90:444-445), ``pos_evento``.
    """
    structure = np.asarray(structure)
    N = structure.shape[1]
    rng = np.random.default_rng(seed)
    acum = np.asarray(acum, dtype=np.int32)
    unit_type = np.ones(N, dtype=np.int32)
    for k, u in enumerate(thresholds):         # basin_stream_type, cuencas.f90:1455-1462
        unit_type[acum > u] = k + 2
    inp = {
        "drena": np.asfortranarray(structure, dtype=np.int32),
        "nceldas": N, "n_reg": int(n_reg), "dt": float(dt), "dxp": float(dxp),
        "unit_type": unit_type[None, :],
        "hill_long": np.asarray(hill_long, np.float32)[None, :],
        "hill_slope": np.asarray(slope, np.float32)[None, :],
        "stream_long": np.asarray(hill_long, np.float32)[None, :],      # cells mode, wmf.py:3700-3708
        "stream_slope": np.asarray(slope, np.float32)[None, :],
        "stream_width": np.ones((1, N), np.float32),
        "elem_area": np.full((1, N), float(dxp) ** 2, np.float32),
        "speed_type": list(speed_type), "calc_niter": 5,
        "retorno_gr": retorno_gr, "retorno_aq": retorno_aq, "storage_constant": 0.001,
        "calib": DEFAULT_CALIB.copy(),
    }
    ks = loguniform(rng, 1e-3, 1e-2, N)
    inp["v_coef"] = np.asfortranarray(np.stack([rng.uniform(1e-4, 1e-3, N), ks, ks / 10.0, np.zeros(N)])
                                      .astype(np.float32))
    inp["v_exp"] = np.ones((4, N), np.float32, order="F")
    inp["h_coef"] = np.asfortranarray(np.stack([rng.uniform(0.1, 1.0, N), rng.uniform(1e-3, 1e-2, N),
                                                rng.uniform(1e-4, 1e-3, N), rng.uniform(0.5, 2.0, N)])
                                      .astype(np.float32))
    inp["h_exp"] = np.asfortranarray(np.tile(np.array([[1.0], [2.0], [1.0], [0.32]], np.float32), (1, N)))
    cap = rng.uniform(20, 120, N).astype(np.float32)
    gra = rng.uniform(50, 300, N).astype(np.float32)
    inp["max_capilar"] = cap[None, :]
    inp["max_gravita"] = gra[None, :]
    inp["max_aquifer"] = (10.0 * gra)[None, :]
    inp["storage"] = np.asfortranarray(np.stack([0.5 * cap, np.full(N, 1e-6), np.full(N, 2.0),
                                                 np.full(N, 35.0), np.full(N, 1e-5)]).astype(np.float32))
    inp["evpserie"] = np.ones(n_reg, np.float32)
    # control points: n_control hash-chosen channel cells (outlet row is implicit)
    control = np.zeros(N, np.int32)
    chan = np.nonzero(unit_type > 1)[0]
    if n_control > 0 and chan.size > 0:
        pick = rng.choice(chan, size=min(n_control, chan.size), replace=False)
        control[pick] = np.arange(1, pick.size + 1) * 10
    inp["control"] = control[None, :]
    control_h = np.zeros(N, np.int32)
    if n_control_h > 0:
        pick = rng.choice(N, size=min(n_control_h, N), replace=False)
        control_h[pick] = np.arange(1, pick.size + 1)
    inp["control_h"] = control_h[None, :]
    # rainfall: three Gaussian-in-time pulses times smooth spatial fields; ~zero_frac of the steps dry
    col = structure[1].astype(np.float64)
    row = structure[2].astype(np.float64)
    cn = (col - col.min()) / max(1.0, np.ptp(col))
    rn = (row - row.min()) / max(1.0, np.ptp(row))
    t = np.arange(n_reg, dtype=np.float64)
    centres = np.array([0.15, 0.45, 0.8]) * n_reg
    width = max(1.0, (1.0 - zero_frac) * n_reg / 3.0 / 3.3)
    G = np.exp(-((t[:, None] - centres[None, :]) / width) ** 2)          # (T,3)
    G[G < 0.065] = 0.0
    F = np.stack([0.55 + 0.45 * np.sin(2 * np.pi * (a * cn + b * rn + ph))
                  for a, b, ph in ((1.0, 0.5, 0.1), (0.3, 1.3, 0.6), (1.7, 0.9, 0.3))])  # (3,N)
    wet = np.nonzero(G.sum(1) > 0)[0]
    rain = np.zeros((wet.size + 1, N), np.int32)
    amp = rain_peak_mm * np.array([1.0, 0.6, 0.8])
    blk = 64
    for s in range(0, wet.size, blk):
        w = wet[s:s + blk]
        rain[1 + s:1 + s + w.size] = np.rint(1000.0 * ((G[w] * amp[None, :]) @ F)).astype(np.int32)
    pos = np.ones(n_reg, np.int32)
    pos[wet] = np.arange(2, wet.size + 2)
    inp["rain"] = rain
    inp["pos_evento"] = pos
    if sediments:
        inp.update({
            "sim_sediments": 1, "sed_factor": 1.0,
            "wi": [0.036, 2.2e-4, 8.6e-7], "diametro": [0.35, 0.016, 0.001],   # wmf.py:4046-4047
            "krus": rng.uniform(0.1, 0.4, N).astype(np.float32)[None, :],
            "crus": rng.uniform(0.001, 0.3, N).astype(np.float32)[None, :],
            "prus": np.ones((1, N), np.float32),
            "parliac": np.asfortranarray(np.tile(np.array([[40.0], [40.0], [20.0]], np.float32), (1, N))),
        })
    if slides:
        inp.update({
            "sim_slides": 1, "sl_fs": 1.0, "sl_gammaw": 9.8, "sl_gullienogullie": 1,
            "sl_zs": rng.uniform(0.2, 1.0, N).astype(np.float32)[None, :],
            "sl_gammas": np.full((1, N), 18.0, np.float32),
            "sl_cohesion": rng.uniform(2, 10, N).astype(np.float32)[None, :],
            "sl_frictionangle": np.deg2rad(rng.uniform(25, 35, N)).astype(np.float32)[None, :],
            # synthetic DEM slopes are far too flat to be susceptible: draw the slope angle directly
            "sl_radslope": np.deg2rad(rng.uniform(20, 50, N)).astype(np.float32)[None, :],
        })
    return inp


def make_rain_types(inp: dict, seed: int = 0) -> dict:
    """Rain-type tables for ``separate_rain`` (modelosv2.f90:452-456, 471-474): for every rain record a convective and a
    stratiform int32 map, stored like the rain (record r at row r-1, record 1 = zeros) and addressed by the same
    per-step record numbers.  A cell counts as convective / stratiform when ``value / 1000`` (integer division) is not
    zero: the maps hold 1000 (yes), 0 (no) and a few 500 (-> 0) and 2000 (-> factor 2) to pin that arithmetic."""
    rain = np.asarray(inp["rain"])
    R, N = rain.shape
    cells = np.arange(N, dtype=np.int64)
    conv = np.zeros((R, N), np.int32)
    stra = np.zeros((R, N), np.int32)
    for r in range(1, R):
        u = u01(seed + 101, np.full(N, r, np.int64), cells)
        wet = rain[r] > 0
        c = wet & (u < 0.45)
        conv[r, c] = 1000
        stra[r, wet & ~c] = 1000
        conv[r, wet & (u > 0.97)] = 500
        stra[r, wet & (u < 0.02)] = 2000
    return {"conv": conv, "stra": stra, "pos_conv": np.array(inp["pos_evento"], np.int32, copy=True),
            "pos_stra": np.array(inp["pos_evento"], np.int32, copy=True), "separate_rain": 1}


def make_stations(nc: int, nr: int, dx: float, n_stations: int = 12, n_reg: int = 100, seed: int = 0, corners: bool = True):
    """Synthetic rain gauges over the raster's extent: (coord (2,ns) fp32, rain (ns,n_reg) fp32 mm per step).  A third of
    the steps are dry everywhere, single readings are missing (-999, as the notebooks' series are) or zero, and one step has
    every gauge missing.  ``corners`` appends the four map corners with the series of the nearest gauge, as
    rain_interpolate_mit does (wmf.py:3298-3307), so that a Delaunay mesh of the gauges covers every cell."""
    rng = np.random.default_rng(seed)
    W, H = nc * dx, nr * dx
    coord = np.vstack([rng.uniform(0.05 * W, 0.95 * W, n_stations), rng.uniform(0.05 * H, 0.95 * H, n_stations)])
    rain = np.zeros((n_stations, n_reg), np.float32)
    for t in range(n_reg):
        if t % 3 == 0:
            continue
        rain[:, t] = rng.gamma(0.8, 4.0, n_stations)
        rain[rng.random(n_stations) < 0.15, t] = -999.0
        rain[rng.random(n_stations) < 0.2, t] = 0.0
    if n_reg > 7:
        rain[:, 7] = -999.0
    if corners:
        for cx, cy in ((0.0, 0.0), (0.0, H), (W, H), (W, 0.0)):
            near = int(np.argmin((coord[0] - cx) ** 2 + (coord[1] - cy) ** 2))
            coord = np.hstack([coord, [[cx], [cy]]])
            rain = np.vstack([rain, rain[near]])
    return np.asfortranarray(coord, dtype=np.float32), np.asfortranarray(rain, dtype=np.float32)


def make_flood_inputs(inp: dict, elev, slope, half_width: int = 4, seed: int = 0) -> dict:
    """HydroFlash inputs (wmf.py:3768-3805, modelosv2.f90:1711-1822): channel width, d50, slope, and for every cell a cross
    section of ``2*half_width+1`` elevations (valley shape around the cell's own elevation) with the cells it lies on.
    The velocity factor / threshold are chosen so that a good share of the type-3 cells are evaluated."""
    N = int(inp["nceldas"])
    rng = np.random.default_rng(seed + 77)
    S = 2 * half_width + 1
    k = np.abs(np.arange(S) - half_width)[:, None].astype(np.float32)
    rise = rng.uniform(0.2, 1.5, (1, N)).astype(np.float32)
    sections = (np.asarray(elev, np.float32).reshape(1, N) + k * rise * rng.uniform(0.6, 1.4, (S, N)).astype(np.float32))
    offs = (np.arange(S) - half_width)[:, None] * rng.integers(1, 40, (1, N))
    cells = np.clip(np.arange(1, N + 1)[None, :] + offs, 1, N).astype(np.float32)
    return {"sim_floods": 1, "flood_dw": 1000.0, "flood_dsed": 2600.0, "flood_av": 100.0, "flood_cmax": 0.75,
            "flood_umbral": 0.7, "flood_step": 0.25, "flood_hmax": 5.0, "flood_max_iter": 10,
            "flood_w": rng.uniform(2.0, 10.0, (1, N)).astype(np.float32), "flood_d50": rng.uniform(0.01, 0.1, (1, N)).astype(np.float32),
            "flood_slope": np.sin(np.arctan(np.maximum(np.asarray(slope, np.float32), 1e-3))).reshape(1, N).astype(np.float32),
            "flood_sections": np.asfortranarray(sections, dtype=np.float32),
            "flood_sec_cells": np.asfortranarray(cells, dtype=np.float32)}


def make_ensemble_calibs(n_members: int, seed: int = 0) -> np.ndarray:
    """(M,11) calibration sets sampled uniformly in the Barbosa NSGA-II notebook ranges
    (Examples/Calibracion_Barbosa_NSGAII.ipynb cell 9), extended to the 11 entries shia_v1 needs."""
    rng = np.random.default_rng(seed)
    lo = np.array([1e-4, 0.5, 0.5, 0.0, 0.05, 5e-5, 1e-3, 1e-3, 0.5, 0.5, 0.5])
    hi = np.array([1e-3, 5.0, 5.0, 0.0, 1.0, 5e-4, 5e-2, 2e-2, 2.0, 2.0, 2.0])
    c = rng.uniform(lo, hi, size=(n_members, 11)).astype(np.float32)
    c[0] = DEFAULT_CALIB
    return c


def apply_to_models(models, inp: dict) -> None:
    """Load a ``make_shia_inputs`` dict into a ``models`` module object the way wmf.py's set_* methods do
    (attribute assignment, wmf.py:3024-3045, 3687-3695, 3882, 3952, 4046-4163)."""
    skip = {"calib", "rain", "pos_evento", "n_reg", "nceldas", "stoin", "hspeedin", "conv", "stra", "pos_conv", "pos_stra"}
    for k, v in inp.items():
        if k in skip:
            continue
        setattr(models, k, v)
    models.nceldas = inp["nceldas"]
