"""CPU ORACLE, second opinion for Path A -- test infrastructure, NOT product code (parity pinned against the reference's own compiled Fortran, see wmf_oracle.h).

Pure-Python transliteration of basin_find / basin_cut / basin_acum written from wmf/cuencas.f90:77-90, 525-607, 659-680
(not from oracle/wmf_oracle_cu.c); tests require both restatements to agree exactly on small random rasters.
"""
import math

import numpy as np

f32 = np.float32


def coord2fil_col(x, y, xll, yll, dx, ncols, nrows):
    """cuencas.f90:77-90 in fp32"""
    col = abs(int(math.ceil(f32(f32(f32(x) - f32(xll)) / f32(dx)))))
    fil = nrows - int(math.floor(f32(f32(f32(y) - f32(yll)) / f32(dx))))
    if col > ncols or col <= 0:
        col = -999
    if fil > nrows or fil <= 0:
        fil = -999
    return col, fil


def basin_find_cut(x, y, DIR, xll, yll, dx):
    """DIR: (ncols, nrows) int array, DIR[c-1, f-1] = code of (col c, row f).  Returns basin_f (3, N)."""
    ncols, nrows = DIR.shape
    coli, fili = coord2fil_col(x, y, xll, yll, dx, ncols, nrows)
    celTot = ncols * nrows
    temp = -np.ones((3, celTot + 1), np.int64)          # 1-based second index, slot 0 = the Fortran's out-of-bounds read
    temp[:, celTot] = (0, coli, fili)
    tenia, cont2 = 1, 1
    col2, row2 = coli, fili
    while col2 > 0:
        res = []
        for kf in (1, 2, 3):
            for kc in (1, 2, 3):
                c, f = col2 + kc - 2, row2 + kf - 2
                if 0 < c <= ncols and 0 < f <= nrows and DIR[c - 1, f - 1] == 3 * kf - kc + 1:
                    res.append((c, f))
        for i, (c, f) in enumerate(res, start=1):
            temp[:, celTot - tenia + 1 - i] = (cont2, c, f)
        tenia += len(res)
        paso = celTot - cont2
        col2, row2 = (temp[1, paso], temp[2, paso]) if paso >= 1 else (-1, -1)
        cont2 += 1
    n = tenia
    return np.asfortranarray(temp[:, celTot - n + 1: celTot + 1].astype(np.int32))   # basin_cut :592-607


def basin_acum(basin_f):
    """cuencas.f90:659-680"""
    n = basin_f.shape[1]
    acum = np.ones(n, np.int64)
    for i in range(n):
        d = int(basin_f[0, i])
        if d != 0:
            acum[n - d] += acum[i]
    return acum.astype(np.int32)
