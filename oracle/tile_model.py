"""CPU MODEL of the tile-decomposed basin_find (test infrastructure, NOT product code).

`wmf_b200/csrc/pathfind.cu` replaces the level-by-level breadth-first search of `basin_find` (cuencas.f90:525-591) by a
tile decomposition whose every stage is a streaming pass.  This file restates that decomposition with plain Python loops
(tiny tiles, tiny rasters) so that the formulas -- local terminals, boundary graph, sibling offsets, preorder, the final
(level, preorder) ranking -- are checked against the oracle's FIFO search without a GPU, and so that the intermediate
arrays of the CUDA path can be compared stage by stage when a kernel misbehaves (tools/debug_pathfind.py).

The identity behind it: the FIFO queue visits cells by level, and inside a level in the order of the DFS preorder taken
with the children in the 3x3 probe order (kf outer, kc inner).  With size(v) = cells draining through v,

    off(c)  = 1 + sum of size(s) over the siblings s probed before c
    pre(v)  = pre(parent) + off(v),   depth(v) = depth(parent) + 1,   pre(outlet) = depth(outlet) = 0
    queue position of v = rank of (depth(v), pre(v)),   basin_f row 1 of v = queue position of its parent + 1
"""
from __future__ import annotations

import numpy as np

# probe order of cuencas.f90:555-571: (kf, kc) = rows top..bottom, columns left..right, centre skipped;
# the neighbour at (dc, dr) is a child iff its DIR equals CODE
PROBES = [(-1, -1, 3), (0, -1, 2), (1, -1, 1), (-1, 0, 6), (1, 0, 4), (-1, 1, 9), (0, 1, 8), (1, 1, 7)]   # (dc, dr, code)
ROOT_OUT, DEAD = -1, -2


def move(code):
    return (code - 1) % 3 - 1, 1 - (code - 1) // 3          # (dc, dr), rows grow downwards


def perim_index(ly, lx, ts):
    if ly == 0:
        return lx
    if ly == ts - 1:
        return ts + lx
    if lx == 0:
        return 2 * ts + (ly - 1)
    assert lx == ts - 1
    return 2 * ts + (ts - 2) + (ly - 1)


def tile_find(DIR_cr, col0, row0, ts=8):
    """DIR_cr: (ncols, nrows) int32 as wmf.py passes it; (col0,row0) 0-based outlet.  Returns (q_lin, q_par, q_acum, stages)."""
    D = np.ascontiguousarray(np.asarray(DIR_cr).T)           # [row, col]
    nr, nc = D.shape
    lin0 = row0 * nc + col0
    ntx, nty = (nc + ts - 1) // ts, (nr + ts - 1) // ts
    slots = 4 * ts                                            # >= perimeter cells of a tile
    # effective direction code: 0 = no downstream (invalid, off the map, or the outlet)
    ed = np.zeros((nr, nc), np.int32)
    for r in range(nr):
        for c in range(nc):
            d = int(D[r, c])
            if 1 <= d <= 9 and d != 5 and (r * nc + c) != lin0:
                dc, dr = move(d)
                if 0 <= c + dc < nc and 0 <= r + dr < nr:
                    ed[r, c] = d

    def down(r, c):
        dc, dr = move(int(ed[r, c]))
        return r + dr, c + dc

    def node_of(r, c):
        return ((r // ts) * ntx + c // ts) * slots + perim_index(r % ts, c % ts, ts)

    def same_tile(r, c, r2, c2):
        return r // ts == r2 // ts and c // ts == c2 // ts

    def local_terminal(r, c):
        """(terminal cell, hops) following in-tile links; None for an in-tile cycle."""
        hops, seen = 0, 0
        while ed[r, c] != 0:
            r2, c2 = down(r, c)
            if not same_tile(r, c, r2, c2):
                break
            r, c = r2, c2
            hops += 1
            seen += 1
            if seen > ts * ts:
                return None, 0
        return (r, c), hops

    n_nodes = ntx * nty * slots
    # ---- stage A: per tile, terminals of the perimeter cells and the local size of every exit cell ----
    pterm = np.full(n_nodes, DEAD, np.int64)      # node -> node of its local terminal (exit cell) | ROOT_OUT | DEAD
    bdown = np.full(n_nodes, -1, np.int64)        # exit node -> node of the cell it drains to
    lsize = np.zeros(n_nodes, np.int64)
    for r in range(nr):
        for c in range(nc):
            t, _ = local_terminal(r, c)
            if t is None:
                code = DEAD
            elif ed[t] != 0:                       # exit cell
                code = node_of(*t)
                lsize[code] += 1
                if (r, c) == t:
                    bdown[code] = node_of(*down(*t))
            else:
                code = ROOT_OUT if t[0] * nc + t[1] == lin0 else DEAD
            ly, lx = r % ts, c % ts
            if ly in (0, ts - 1) or lx in (0, ts - 1):
                pterm[node_of(r, c)] = code
    # ---- stage B: boundary graph, global size of the exit cells ----
    is_exit = bdown >= 0
    bnext = np.full(n_nodes, DEAD, np.int64)
    bnext[is_exit] = pterm[bdown[is_exit]]
    G = lsize.copy()
    order = [x for x in np.nonzero(is_exit)[0]]
    # leaves first: repeat until stable (tiny inputs)
    pend = np.zeros(n_nodes, np.int64)
    for x in order:
        if bnext[x] >= 0:
            pend[bnext[x]] += 1
    ready = [x for x in order if pend[x] == 0]
    while ready:
        x = ready.pop()
        y = bnext[x]
        if y >= 0:
            G[y] += G[x]
            pend[y] -= 1
            if pend[y] == 0:
                ready.append(y)
    # ---- stage C: per tile, global sizes, sibling offsets, local downward sums ----
    S = np.zeros((nr, nc), np.int64)
    # accumulate in-tile from leaves: process cells by decreasing local hops
    val = np.ones((nr, nc), np.int64)
    for r in range(nr):
        for c in range(nc):
            if ed[r, c] != 0:
                r2, c2 = down(r, c)
                if not same_tile(r, c, r2, c2):
                    val[r2, c2] += G[node_of(r, c)]     # external inflow entering at (r2, c2)
    hops_all = {}
    for r in range(nr):
        for c in range(nc):
            t, h = local_terminal(r, c)
            hops_all[(r, c)] = (t, h)
    S[:] = val
    for (r, c), (t, h) in sorted(hops_all.items(), key=lambda kv: -kv[1][1]):
        if t is None or ed[r, c] == 0:
            continue
        r2, c2 = down(r, c)
        if same_tile(r, c, r2, c2):
            S[r2, c2] += S[r, c]
    off = np.zeros((nr, nc), np.int64)
    nchild = np.zeros((nr, nc), np.int64)
    for r in range(nr):
        for c in range(nc):
            run = 1
            for dc, dr, code in PROBES:
                r2, c2 = r + dr, c + dc
                if 0 <= r2 < nr and 0 <= c2 < nc and ed[r2, c2] == code:
                    off[r2, c2] = run
                    run += S[r2, c2] if same_tile(r, c, r2, c2) else G[node_of(r2, c2)]
                    nchild[r, c] += 1
    LT = {}                                         # cell -> (terminal code, hops, offsum)
    eh = np.zeros(n_nodes, np.int64)
    eo = np.zeros(n_nodes, np.int64)
    boff = np.zeros(n_nodes, np.int64)
    for r in range(nr):
        for c in range(nc):
            t, h = hops_all[(r, c)]
            if t is None:
                LT[(r, c)] = (DEAD, 0, 0)
                continue
            o, rr, cc = 0, r, c
            while (rr, cc) != t:
                o += off[rr, cc]
                rr, cc = down(rr, cc)
            code = node_of(*t) if ed[t] != 0 else (ROOT_OUT if t[0] * nc + t[1] == lin0 else DEAD)
            LT[(r, c)] = (code, h, o)
            ly, lx = r % ts, c % ts
            if ly in (0, ts - 1) or lx in (0, ts - 1):
                eh[node_of(r, c)], eo[node_of(r, c)] = h, o
            if (r, c) == t and ed[t] != 0:
                boff[node_of(r, c)] = off[r, c]
    # ---- stage D: depth / preorder of the exit cells by walking the boundary graph ----
    Dn = np.zeros(n_nodes, np.int64)
    Pn = np.zeros(n_nodes, np.int64)
    alive = np.zeros(n_nodes, bool)
    for x in order:
        d = p = 0
        y, steps = x, 0
        while y >= 0 and steps <= n_nodes:
            d += 1 + eh[bdown[y]]
            p += boff[y] + eo[bdown[y]]
            y = bnext[y]
            steps += 1
        if y == ROOT_OUT:
            Dn[x], Pn[x], alive[x] = d, p, True
    # ---- stage E: (depth, preorder) of every basin cell, ranking ----
    cells = []
    for r in range(nr):
        for c in range(nc):
            code, h, o = LT[(r, c)]
            if code == DEAD:
                continue
            if code == ROOT_OUT:
                cells.append((h, o, r * nc + c))
            elif alive[code]:
                cells.append((Dn[code] + h, Pn[code] + o, r * nc + c))
    cells.sort()
    n = len(cells)
    assert n == S[row0, col0], (n, S[row0, col0])
    assert sorted(p for _, p, _ in cells) == list(range(n)), "preorder is not a permutation"
    q_lin = np.array([l for _, _, l in cells], np.int64)
    q_acum = np.array([S[l // nc, l % nc] for l in q_lin], np.int64)
    q_par = np.zeros(n, np.int64)
    cs = 1
    for j, l in enumerate(q_lin):
        k = nchild[l // nc, l % nc]
        q_par[cs:cs + k] = j + 1
        cs += k
    assert cs == n
    stages = dict(ed=ed, pterm=pterm, bdown=bdown, lsize=lsize, bnext=bnext, G=G, S=S, off=off, nchild=nchild, eh=eh, eo=eo,
                  boff=boff, Dn=Dn, Pn=Pn, alive=alive, LT=LT)
    return q_lin, q_par, q_acum, stages


def basin_table(q_lin, q_par, nc):
    """basin_f (3,N) as basin_cut returns it (cuencas.f90:592-607): the queue reversed."""
    n = q_lin.size
    out = np.empty((3, n), np.int32)
    out[0] = q_par[::-1]
    out[1] = q_lin[::-1] % nc + 1
    out[2] = q_lin[::-1] // nc + 1
    return out
