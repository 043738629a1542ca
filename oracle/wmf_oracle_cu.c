/*
 * wmf_oracle_cu.c -- CPU ORACLE (test infrastructure, NOT product code).
 * Literal restatement of the Path-A routines of wmf/cuencas.f90 (module cu).
 * See wmf_oracle.h for the status of this file (pinned against oracle/_ref).
 *
 * Fenced quirks (reference behaviour is undefined / non-terminating there):
 *  - basin_find: the reference's res(2,7) scratch holds 7 children although 8
 *    are possible (cuencas.f90:537); the oracle accepts 8.
 *  - basin_find: the centre probe (kf=2,kc=2, DIR==5) would make a cell its own
 *    child and never terminate (cuencas.f90:562); the oracle skips the centre.
 *  - basin_find: termination reads basin_temp(2,celTot-cont2), index 0 when the
 *    basin fills the whole raster (cuencas.f90:584-586); the oracle stops.
 *  - basin_basics: pend(i-1) for an outlet at i==1 (cuencas.f90:640) -> 0.
 *  - basin_map2basin: column/row computed with ceiling() can land one past the
 *    map edge through fp32 rounding (cuencas.f90:1171-1174); clamped.
 *  - geo_hand: a non-channel outlet indexes N+1 (cuencas.f90:2156) -> 0,0,0.
 */
#include "wmf_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define BF(k, i) basin_f[3 * (size_t)(i) + (k)] /* k = 0 drain, 1 col, 2 row; i 0-based */

/* cuencas.f90:77-90 */
void wo_coord2fil_col(const wo_geo *g, float x, float y, int32_t *col, int32_t *fil)
{
    int32_t c = (int32_t)ceilf((x - g->xll) / g->dx);
    if (c < 0) c = -c;
    int32_t f = g->nrows - (int32_t)floorf((y - g->yll) / g->dx);
    if (c > g->ncols || c <= 0) c = -999;
    if (f > g->nrows || f <= 0) f = -999;
    *col = c;
    *fil = f;
}

/* cuencas.f90:525-591 */
int32_t wo_basin_find(const wo_geo *g, float x, float y, const int32_t *dir, int32_t nc, int32_t nr,
                      int32_t *bt)
{
    const int32_t ncols = g->ncols, nrows = g->nrows;
    const int64_t celTot = (int64_t)ncols * nrows;
    int32_t coli, fili;
    wo_coord2fil_col(g, x, y, &coli, &fili);
    for (int64_t i = 0; i < 3 * celTot; ++i) bt[i] = -1;
#define BT(k, p) bt[3 * ((p)-1) + ((k)-1)] /* 1-based (k,p) */
    BT(1, celTot) = 0;
    BT(2, celTot) = coli;
    BT(3, celTot) = fili;
    int64_t tenia = 1, cont2 = 1;
    int32_t col2 = coli, row2 = fili;
    int32_t res[2][8];
    while (col2 > 0) {
        int cont3 = 0;
        for (int kf = 1; kf <= 3; ++kf) {
            for (int kc = 1; kc <= 3; ++kc) {
                if (kf == 2 && kc == 2) continue; /* fenced: DIR==5 self loop */
                int32_t c = col2 + kc - 2, f = row2 + kf - 2;
                if (c <= ncols && c > 0 && f <= nrows && f > 0) {
                    if (dir[(size_t)(f - 1) * nc + (c - 1)] == 3 * kf - kc + 1) {
                        res[0][cont3] = c;
                        res[1][cont3] = f;
                        ++cont3;
                    }
                }
            }
        }
        for (int i = 1; i <= cont3; ++i) {
            int64_t p = celTot - tenia + 1 - i;
            if (p < 1) return -1; /* cyclic DIR: queue overflow */
            BT(1, p) = (int32_t)cont2;
            BT(2, p) = res[0][i - 1];
            BT(3, p) = res[1][i - 1];
        }
        tenia += cont3;
        int64_t paso = celTot - cont2;
        if (paso < 1) break; /* fenced: basin fills the raster */
        col2 = BT(2, paso);
        row2 = BT(3, paso);
        ++cont2;
    }
#undef BT
    (void)nr;
    return (int32_t)tenia;
}

/* cuencas.f90:592-607 */
void wo_basin_cut(const wo_geo *g, const int32_t *bt, int32_t n, int32_t *basin_f)
{
    const int64_t celTot = (int64_t)g->ncols * g->nrows;
    memcpy(basin_f, bt + 3 * (celTot - n), sizeof(int32_t) * 3 * (size_t)n);
}

/* cuencas.f90:659-680 */
void wo_basin_acum(const wo_geo *g, const int32_t *basin_f, int32_t n, int32_t *acum, float *area)
{
    for (int32_t i = 0; i < n; ++i) acum[i] = 1;
    for (int32_t i = 0; i < n; ++i) {
        int32_t d = BF(0, i);
        if (d != 0) {
            int32_t drenaid = n - d; /* 0-based */
            acum[drenaid] = acum[drenaid] + acum[i];
        }
    }
    if (area) *area = (float)n * (g->dxp * g->dxp) / 1e6f;
}

/* cuencas.f90:681-701 ; drain is the first row only */
void wo_basin_acum_var(const int32_t *drain, int32_t n, const float *var, float *acum)
{
    for (int32_t i = 0; i < n; ++i) acum[i] = var[i];
    for (int32_t i = 0; i < n; ++i) {
        if (drain[i] != 0) {
            int32_t drenaid = n - drain[i];
            acum[drenaid] = acum[drenaid] + acum[i];
        }
    }
}

/* cuencas.f90:797-808 */
void wo_basin_coordxy(const wo_geo *g, const int32_t *basin_f, int32_t n, float *x, float *y)
{
    for (int32_t i = 0; i < n; ++i) {
        x[i] = g->xll + g->dx * ((float)BF(1, i) - 0.5f);
        y[i] = g->yll + g->dx * ((float)(g->nrows - BF(2, i)) + 0.5f);
    }
}

static int cmp_float(const void *a, const void *b)
{
    float fa = *(const float *)a, fb = *(const float *)b;
    return (fa > fb) - (fa < fb);
}

/* cuencas.f90:608-658 */
void wo_basin_basics(const wo_geo *g, const int32_t *basin_f, const float *dem, const int32_t *dir,
                     int32_t n, int32_t *acum, float *lng, float *pend, float *elev,
                     wo_basin_scalars *sc)
{
    const float dxp = g->dxp;
    for (int32_t i = 0; i < n; ++i) {
        elev[i] = dem[i];
        acum[i] = 1;
    }
    for (int32_t i = 0; i < n; ++i) {
        int32_t d = BF(0, i);
        int32_t drenaid = n - d;
        float prueba = (float)(dir[i] % 2);
        if (prueba == 0.0f)
            lng[i] = dxp;
        else
            lng[i] = dxp * sqrtf(2.0f);
        if (d != 0) {
            acum[drenaid] = acum[drenaid] + acum[i];
            pend[i] = fabsf(dem[i] - dem[drenaid]) / lng[i];
        } else {
            pend[i] = (i > 0) ? pend[i - 1] : 0.0f; /* fenced for N==1 */
        }
        if (pend[i] == 0.0f) pend[i] = 0.001f;
    }
    if (!sc) return;
    sc->area = (float)n * (dxp * dxp) / 1e6f;
    float s = 0.0f;
    for (int32_t i = 0; i < n; ++i) s += pend[i];
    sc->pend_media = s / (float)n;
    s = 0.0f;
    for (int32_t i = 0; i < n; ++i) s += dem[i];
    sc->elevacion = s / (float)n;
    float *X = (float *)malloc(sizeof(float) * (size_t)n), *Y = (float *)malloc(sizeof(float) * (size_t)n);
    wo_basin_coordxy(g, basin_f, n, X, Y);
    qsort(X, (size_t)n, sizeof(float), cmp_float);
    qsort(Y, (size_t)n, sizeof(float), cmp_float);
    int32_t k = (n % 2 == 0) ? n / 2 : (n + 1) / 2; /* 1-based */
    sc->centrox = X[k - 1];
    sc->centroy = Y[k - 1];
    free(X);
    free(Y);
}

/* cuencas.f90:1144-1184 */
void wo_basin_map2basin(const wo_geo *g, const int32_t *basin_f, int32_t n, const float *mapa,
                        float xllm, float yllm, int32_t ncolsm, int32_t nrowsm, float dxm, float dym,
                        float nodatam, float *vec)
{
    int64_t cantidad = 0;
    float media = 0.0f;
    const int64_t tot = (int64_t)ncolsm * nrowsm;
    for (int64_t k = 0; k < tot; ++k)
        if (mapa[k] != nodatam) {
            ++cantidad;
            media += mapa[k];
        }
    media = media / (float)(int32_t)cantidad;
    for (int32_t i = 0; i < n; ++i) {
        vec[i] = nodatam;
        float xpos = g->xll + g->dx * ((float)BF(1, i) - 0.5f);
        float ypos = g->yll + g->dy * ((float)(g->nrows - BF(2, i)) + 0.5f);
        if (xpos > xllm && xpos < (xllm + dxm * (float)ncolsm) && ypos > yllm &&
            ypos < (yllm + (float)nrowsm * dym)) {
            int32_t columna = (int32_t)ceilf((xpos - xllm) / dxm);
            int32_t fila = (int32_t)ceilf((ypos - yllm) / dym);
            fila = nrowsm - fila + 1;
            if (columna < 1) columna = 1; /* fenced */
            if (columna > ncolsm) columna = ncolsm;
            if (fila < 1) fila = 1;
            if (fila > nrowsm) fila = nrowsm;
            vec[i] = mapa[(size_t)(fila - 1) * ncolsm + (columna - 1)];
            if (vec[i] == nodatam) vec[i] = media;
        }
    }
}

/* cuencas.f90:1185-1201 */
void wo_basin_2map_find(const int32_t *basin_f, int32_t n, int32_t *map_ncols, int32_t *map_nrows)
{
    int32_t cmin = BF(1, 0), cmax = BF(1, 0), fmin = BF(2, 0), fmax = BF(2, 0);
    for (int32_t i = 1; i < n; ++i) {
        if (BF(1, i) < cmin) cmin = BF(1, i);
        if (BF(1, i) > cmax) cmax = BF(1, i);
        if (BF(2, i) < fmin) fmin = BF(2, i);
        if (BF(2, i) > fmax) fmax = BF(2, i);
    }
    *map_ncols = cmax - cmin + 1;
    *map_nrows = fmax - fmin + 1;
}

/* cuencas.f90:1202-1229 */
void wo_basin_2map(const wo_geo *g, const int32_t *basin_f, const float *var, int32_t n,
                   int32_t map_ncols, int32_t map_nrows, float *mapa, float *map_xll, float *map_yll)
{
    int32_t cmin = BF(1, 0), fmin = BF(2, 0), fmax = BF(2, 0);
    for (int32_t i = 1; i < n; ++i) {
        if (BF(1, i) < cmin) cmin = BF(1, i);
        if (BF(2, i) < fmin) fmin = BF(2, i);
        if (BF(2, i) > fmax) fmax = BF(2, i);
    }
    *map_xll = g->xll + g->dx * (float)(cmin - 1);
    *map_yll = g->yll + g->dx * (float)(g->nrows - fmax);
    for (int64_t k = 0; k < (int64_t)map_ncols * map_nrows; ++k) mapa[k] = g->nodata;
    for (int32_t i = 0; i < n; ++i) {
        int32_t col_rel = BF(1, i) - cmin + 1, fil_rel = BF(2, i) - fmin + 1;
        mapa[(size_t)(fil_rel - 1) * map_ncols + (col_rel - 1)] = var[i];
    }
}

/* cuencas.f90:1037-1051 */
void wo_basin_float_var2map(const wo_geo *g, const int32_t *basin_f, const float *vect, int32_t n,
                            int32_t nc, int32_t nf, float *mapa)
{
    for (int64_t k = 0; k < (int64_t)nc * nf; ++k) mapa[k] = g->nodata;
    for (int32_t i = 0; i < n; ++i) mapa[(size_t)(BF(2, i) - 1) * nc + (BF(1, i) - 1)] = vect[i];
}

/* cuencas.f90:1340-1379 */
void wo_basin_stream_nod(const int32_t *basin_f, const int32_t *acum, int32_t n, int32_t umbral,
                         int32_t *cauce, int32_t *nodos, int32_t *trazado, int32_t *n_nodos,
                         int32_t *n_cauce)
{
    int32_t nca = 0;
    for (int32_t i = 0; i < n; ++i) {
        cauce[i] = (acum[i] >= umbral) ? 1 : 0;
        nca += cauce[i];
        nodos[i] = 0;
    }
    *n_cauce = nca;
    for (int32_t i = 0; i < n; ++i) {
        int32_t d = BF(0, i);
        if (d != 0 && cauce[i] == 1) nodos[n - d] += 1;
    }
    for (int32_t i = 0; i < n; ++i)
        if (nodos[i] == 0 && cauce[i] == 1) nodos[i] = 3;
    nodos[n - 1] = 2;
    int32_t nn = 0;
    for (int32_t i = 0; i < n; ++i) {
        if (nodos[i] < 2) nodos[i] = 0;
        if (nodos[i] > 0) ++nn;
    }
    *n_nodos = nn;
    for (int32_t i = 0; i < n; ++i) trazado[i] = 0;
    trazado[n - 1] = 1;
    for (int32_t i = 0; i < n; ++i) {
        int32_t d = BF(0, i);
        if (d != 0 && nodos[n - d] != 0 && cauce[i] == 1) trazado[i] = 1;
    }
}

/* cuencas.f90:1439-1463 */
void wo_basin_stream_type(const int32_t *acum, int32_t n, const int32_t *umbrales, int32_t numbrales,
                          int32_t *stream_types)
{
    for (int32_t j = 0; j < n; ++j) stream_types[j] = 1;
    for (int32_t i = 0; i < numbrales; ++i)
        for (int32_t j = 0; j < n; ++j)
            if (acum[j] > umbrales[i]) stream_types[j] = i + 2;
}

/* cuencas.f90:2120-2129 */
void wo_geo_acum_to_cauce(const int32_t *acum, int32_t ncells, int32_t umbral, int32_t *cauce)
{
    for (int32_t k = 0; k < ncells; ++k) cauce[k] = (acum[k] > umbral) ? 1 : 0;
}

/* cuencas.f90:2131-2174 ; a_quien is 1-based cell id (0 = none), channel cells get 0 */
void wo_geo_hand(const int32_t *basin_f, const float *elev, const float *lng, const int32_t *cauce,
                 int32_t n, float *hand, float *hdnd, int32_t *a_quien)
{
    for (int32_t i = 0; i < n; ++i) {
        hand[i] = 0.0f;
        hdnd[i] = (cauce[i] == 1) ? 0.0f : lng[i];
        a_quien[i] = 0;
    }
    for (int32_t i = 0; i < n; ++i) {
        if (cauce[i] == 1) continue;
        int32_t it = i;
        float lsum = lng[i];
        for (;;) {
            int32_t d = BF(0, it);
            if (d == 0) { /* fenced: non-channel outlet */
                hand[i] = 0.0f;
                hdnd[i] = 0.0f;
                a_quien[i] = 0;
                break;
            }
            int32_t drenaid = n - d;
            if (cauce[drenaid] == 1) {
                hand[i] = elev[i] - elev[drenaid];
                hdnd[i] = lsum;
                a_quien[i] = drenaid + 1;
                break;
            } else if (BF(0, drenaid) == 0) {
                hand[i] = 0.0f;
                hdnd[i] = 0.0f;
                a_quien[i] = 0;
                break;
            } else {
                it = drenaid;
                lsum = lsum + lng[drenaid];
            }
        }
    }
}

/* cuencas.f90:2679-2703 */
static int drain_colfil(int32_t d, int32_t *dc, int32_t *df)
{
    switch (d) {
    case 7: *dc = -1; *df = -1; return 1;
    case 8: *dc = 0; *df = -1; return 1;
    case 9: *dc = 1; *df = -1; return 1;
    case 4: *dc = -1; *df = 0; return 1;
    case 6: *dc = 1; *df = 0; return 1;
    case 1: *dc = -1; *df = 1; return 1;
    case 2: *dc = 0; *df = 1; return 1;
    case 3: *dc = 1; *df = 1; return 1;
    default: return 0; /* reference adds noData to col/fil -> leaves the map */
    }
}

/* cuencas.f90:2175-2219 */
void wo_geo_hand_global(const wo_geo *g, const float *dem, const int32_t *dir, const int32_t *red,
                        int32_t nc, int32_t nr, float *hand)
{
    const float nodata = g->nodata;
#define M(a, c, f) a[(size_t)((f)-1) * nc + ((c)-1)]
    for (int64_t k = 0; k < (int64_t)nc * nr; ++k) hand[k] = (red[k] == 1) ? 0.0f : nodata;
    for (int32_t i = 1; i <= nc; ++i) {
        for (int32_t j = 1; j <= nr; ++j) {
            if ((float)M(dir, i, j) != nodata && M(red, i, j) == 0) {
                int32_t col = i, fil = j;
                int flag = 1;
                int64_t guard = 0;
                while (flag == 1 && (float)M(dir, col, fil) != nodata) {
                    int32_t dc, df;
                    if (!drain_colfil(M(dir, col, fil), &dc, &df)) {
                        M(hand, i, j) = nodata;
                        break;
                    }
                    col += dc;
                    fil += df;
                    if (col > 0 && col <= g->ncols && fil > 0 && fil <= g->nrows && col <= nc && fil <= nr) {
                        if (M(red, col, fil) == 1) {
                            M(hand, i, j) = M(dem, i, j) - M(dem, col, fil);
                            flag = 0;
                        } else if ((float)M(red, col, fil) == nodata) {
                            M(hand, i, j) = nodata;
                            flag = 0;
                        }
                    } else {
                        M(hand, i, j) = nodata;
                        flag = 0;
                    }
                    if (++guard > (int64_t)nc * nr) break; /* fenced: cyclic DIR */
                }
            }
        }
    }
#undef M
}

/* cuencas.f90:724-756 */
void wo_basin_time_to_out(const int32_t *basin_f, const float *lng, const float *speed, int32_t n,
                          float *time)
{
    float *tc = (float *)malloc(sizeof(float) * (size_t)n);
    for (int32_t i = 0; i < n; ++i) tc[i] = lng[i] / speed[i];
    for (int32_t i = 1; i <= n; ++i) {
        int32_t drenaid = n - BF(0, i - 1) + 1; /* 1-based */
        if (drenaid < n) {
            float acc = tc[i - 1];
            int flag = 1;
            while (flag) {
                acc = acc + tc[drenaid - 1];
                drenaid = n - BF(0, drenaid - 1) + 1;
                if (drenaid >= n) flag = 0;
            }
            time[i - 1] = acc;
        } else {
            time[i - 1] = tc[i - 1];
        }
    }
    free(tc);
}

/* cuencas.f90:1811-1833 */
void wo_basin_propagate(const int32_t *basin_f, const float *var, int32_t n, float *var_prop)
{
    for (int32_t i = 0; i < n; ++i) var_prop[i] = var[i];
    for (int32_t i = 0; i < n; ++i) {
        if (var_prop[i] > 0.0f) {
            int32_t drenaid = n - BF(0, i) + 1; /* 1-based */
            if (drenaid <= n) var_prop[drenaid - 1] = var_prop[drenaid - 1] + var_prop[i];
        }
    }
}

/* ---- sub-basin (hills) topology, cuencas.f90:1380-1438, 1835-2118 ---- */

/* basin_stream_slope, cuencas.f90:1380-1438.  celdas_tramo(2) is not written when the cell below a node is itself a node
 * (:1414-1419), yet the assignment loop runs to cont = 2 (:1432-1435): the reference then writes this reach's values into
 * whichever cell an earlier, longer reach left in celdas_tramo(2).  Before any such reach the entry is uninitialised
 * (index 0 with a zeroed heap = one element before the arrays): that write is dropped here. */
void wo_basin_stream_slope(const wo_geo *g, const int32_t *basin_f, const float *elev, const float *lng,
                           const int32_t *nodos, int32_t n_cauce, int32_t n, float *stream_s, float *stream_l)
{
    const int32_t cap = n_cauce > 2 ? n_cauce : 2;
    int32_t *tramo = (int32_t *)calloc((size_t)cap, sizeof(int32_t));   /* 1-based cell ids, 0 = never written */
    float *ed = (float *)calloc(2 * (size_t)cap, sizeof(float));
    for (int32_t i = 0; i < n; ++i) { stream_s[i] = g->nodata; stream_l[i] = g->nodata; }
    for (int32_t i = 0; i < n; ++i) {
        if (nodos[i] > 0 && BF(0, i) != 0) {
            ed[0] = elev[i];
            ed[1] = lng[i];
            tramo[0] = i + 1;
            int flag = 1;
            int32_t cont = 1;
            int32_t d = n - BF(0, i); /* 0-based */
            while (flag == 1) {
                if (nodos[d] == 0) {
                    cont = cont + 1;
                    ed[2 * (cont - 1)] = elev[d];
                    ed[2 * (cont - 1) + 1] = ed[2 * (cont - 2) + 1] + lng[d];
                    tramo[cont - 1] = d + 1;
                } else {
                    flag = 0;
                    if (cont == 1) {
                        cont = cont + 1;
                        ed[2 * (cont - 1)] = elev[d];
                        ed[2 * (cont - 1) + 1] = ed[2 * (cont - 2) + 1] + lng[d];
                    }
                }
                d = n - BF(0, d); /* may step past the outlet (index n): never read again, flag is 0 by then */
                if (d >= n) d = n - 1;
            }
            float S = fabsf((ed[0] - ed[2 * (cont - 1)]) / (ed[1] - ed[2 * (cont - 1) + 1]));
            if (S <= 0.0f) S = 0.001f;
            /* L_tramo is implicitly typed (module cu has no `implicit none`): a name starting with L is INTEGER, so the
               reach length is truncated, and so is dxP in the fallback */
            int32_t Li = (int32_t)ed[2 * (cont - 1) + 1];
            if (Li <= 0) Li = (int32_t)g->dxp;
            const float L = (float)Li;
            for (int32_t j = 0; j < cont; ++j)
                if (tramo[j] >= 1) { stream_s[tramo[j] - 1] = S; stream_l[tramo[j] - 1] = L; }
        }
    }
    free(tramo);
    free(ed);
}

/* basin_subbasin_nod, cuencas.f90:1835-1920 (+ the module global sub_basins_temp(2,n_nodos) that basin_subbasin_cut
 * :1921-1930 returns): sub_basins must hold 2*n entries (only the first 2*n_nodos are meaningful). */
void wo_basin_subbasin_nod(const int32_t *basin_f, const int32_t *acum, int32_t n, int32_t umbral, int32_t *cauce,
                           int32_t *nodos_fin, int32_t *n_nodos, int32_t *sub_basins)
{
    int32_t *nodos = (int32_t *)calloc((size_t)n, sizeof(int32_t));
    for (int32_t i = 0; i < n; ++i) cauce[i] = (acum[i] >= umbral) ? 1 : 0;
    for (int32_t i = 0; i < n; ++i)
        if (BF(0, i) != 0 && cauce[i] == 1) nodos[n - BF(0, i)] += 1;
    for (int32_t i = 0; i < n; ++i)
        if (nodos[i] < 2) nodos[i] = 0;
    for (int32_t i = 0; i < n; ++i) nodos_fin[i] = 0;
    nodos_fin[n - 1] = 1;
    for (int32_t i = 0; i < n; ++i)
        if (cauce[i] == 1) {
            const int32_t d = n - BF(0, i); /* 0-based; == n for the outlet */
            if (d < n && nodos[d] != 0) nodos_fin[i] = 1;
        }
    int32_t nn = 0;
    for (int32_t i = 0; i < n; ++i) nn += nodos_fin[i];
    *n_nodos = nn;
#define SB(k, j) sub_basins[2 * (size_t)((j) - 1) + ((k) - 1)] /* sub_basins_temp(k,j), 1-based */
    for (int32_t j = 1; j <= nn; ++j) { SB(1, j) = 0; SB(2, j) = 0; }
    SB(1, nn) = n;
    SB(2, nn) = 0;
    int32_t cont = 1, cont2 = 0;
    int flag1 = 1;
    while (flag1) {
        if (SB(1, nn - cont + 1) != 0) {
            const int32_t posi = SB(1, nn - cont + 1); /* 1-based cell */
            for (int32_t j = 1; j <= posi - 1; ++j) {
                const int32_t c = posi - j; /* 1-based */
                if (nodos_fin[c - 1] != 0) {
                    int32_t d = n - BF(0, c - 1) + 1; /* 1-based */
                    int flag2 = 1;
                    while (flag2) {
                        if (nodos_fin[d - 1] == 0) {
                            if (d != posi) d = n - BF(0, d - 1) + 1;
                        } else if (d == posi) {
                            cont2 = cont2 + 1;
                            SB(1, nn - cont2) = c;
                            SB(2, nn - cont2) = cont;
                            nodos_fin[c - 1] = cont2 + 1;
                            flag2 = 0;
                        } else {
                            flag2 = 0;
                        }
                    }
                }
            }
            cont = cont + 1;
            if (cont > nn) flag1 = 0;
        } else {
            flag1 = 0;
        }
    }
#undef SB
    free(nodos);
}

/* basin_subbasin_find, cuencas.f90:1979-2010 (sub_basin is declared intent(out) but never assigned: not returned here) */
void wo_basin_subbasin_find(const int32_t *basin_f, const int32_t *nodos, int32_t n, int32_t *sub_pert)
{
    for (int32_t i = 0; i < n; ++i) sub_pert[i] = nodos[i];
    for (int32_t i = 0; i < n; ++i)
        if (sub_pert[i] == 0) {
            int32_t d = n - BF(0, i);
            while (nodos[d] == 0) d = n - BF(0, d);
            sub_pert[i] = nodos[d];
        }
}

/* basin_subbasin_horton, cuencas.f90:1931-1978.  sub_basins (2,n_nodos); nod_horton (nceldas) is only written at the node
 * cells (the Fortran leaves the rest uninitialised: zeros here); at most 10 tributaries are looked at (drenantes(10)). */
void wo_basin_subbasin_horton(const int32_t *sub_basins, int32_t n_nodos, int32_t n, int32_t *sub_horton, int32_t *nod_horton)
{
#define SB(k, j) sub_basins[2 * (size_t)((j) - 1) + ((k) - 1)]
    for (int32_t i = 0; i < n; ++i) nod_horton[i] = 0;
    for (int32_t i = 1; i <= n_nodos; ++i) sub_horton[i - 1] = 0;
    for (int32_t i = 1; i <= n_nodos; ++i) {
        const int32_t posi = n_nodos - i + 1;
        int32_t conteo = 0;
        for (int32_t j = 1; j <= n_nodos; ++j) conteo += (SB(2, j) == posi);
        if (conteo == 0) {
            sub_horton[i - 1] = 1;
            nod_horton[SB(1, i) - 1] = 1;
        }
    }
    for (int32_t i = 1; i <= n_nodos; ++i)
        if (sub_horton[i - 1] != 1) {
            const int32_t posi = n_nodos - i + 1;
            int32_t cont = 0, dren[10] = {0};
            for (int32_t j = 1; j <= i - 1; ++j)
                if (SB(2, j) == posi) {
                    cont = cont + 1;
                    if (cont <= 10) dren[cont - 1] = sub_horton[j - 1];
                }
            if (cont > 10) cont = 10;
            int32_t mx = dren[0], mn = dren[0];
            for (int32_t k = 1; k < cont; ++k) { if (dren[k] > mx) mx = dren[k]; if (dren[k] < mn) mn = dren[k]; }
            sub_horton[i - 1] = (mx == mn) ? mx + 1 : mx;
            nod_horton[SB(1, i) - 1] = sub_horton[i - 1];
        }
#undef SB
}

/* basin_subbasin_long, cuencas.f90:2011-2051.  `long` is declared INTEGER there: f2py truncates the cell lengths the caller
 * passes (wmf.py:3670 hands over hill_long, fp32) before the sums. */
void wo_basin_subbasin_long(const int32_t *sub_pert, const int32_t *cauce, const int32_t *lng_int, const int32_t *sub_basin,
                            const int32_t *sub_horton, int32_t n_nodos, int32_t n, float *sub_basin_long, float *max_long,
                            int32_t *nodo_max_long)
{
#define SB(k, j) sub_basin[2 * (size_t)((j) - 1) + ((k) - 1)]
    for (int32_t i = 1; i <= n_nodos; ++i) {
        const int32_t posi = n_nodos - i + 1;
        float s = 0.0f;
        for (int32_t c = 0; c < n; ++c)
            if (cauce[c] * sub_pert[c] == posi) s += (float)(cauce[c] * lng_int[c]);
        sub_basin_long[i - 1] = s;
    }
    *max_long = 0.0f;
    *nodo_max_long = 0;
    for (int32_t i = 1; i <= n_nodos; ++i)
        if (sub_horton[i - 1] == 1) {
            float acc = sub_basin_long[i - 1];
            int32_t d = n_nodos - SB(2, i) + 1;
            while (d >= 1 && d <= n_nodos && SB(2, d) >= 1) {
                acc = acc + sub_basin_long[d - 1];
                d = n_nodos - SB(2, d) + 1;
            }
            if (acc > *max_long) { *max_long = acc; *nodo_max_long = n_nodos - i + 1; }
        }
#undef SB
}

/* basin_subbasin_stream_prop, cuencas.f90:2093-2115 (indexes the hills by i, not by n_nodos-i+1; 0/0 = NaN for a hill
 * without channel cells) */
void wo_basin_subbasin_stream_prop(const int32_t *sub_pert, const int32_t *cauce, const float *lng, const float *slope,
                                   int32_t n, int32_t n_nodos, float *stream_slope, float *stream_long)
{
    for (int32_t i = 1; i <= n_nodos; ++i) {
        float sl = 0.0f, ss = 0.0f;
        int32_t cnt = 0;
        for (int32_t c = 0; c < n; ++c)
            if (sub_pert[c] == i && cauce[c] == 1) { sl += lng[c]; ss += slope[c]; ++cnt; }
        stream_long[i - 1] = sl;
        stream_slope[i - 1] = ss / (float)cnt;
    }
}

/* basin_subbasin_map2subbasin, cuencas.f90:2052-2092 */
void wo_basin_subbasin_map2subbasin(const int32_t *sub_pert, const float *var, int32_t n_nodos, int32_t n,
                                    const int32_t *cauce, int32_t sum_mean, float *out)
{
    for (int32_t i = 1; i <= n_nodos; ++i) {
        const int32_t posi = n_nodos - i + 1;
        float s = 0.0f, mx = -3.4028235e38f; /* maxval of an empty set is -huge */
        int32_t cnt = 0;
        for (int32_t c = 0; c < n; ++c)
            if (sub_pert[c] == posi && cauce[c] == 1) { s += var[c]; ++cnt; if (var[c] > mx) mx = var[c]; }
        if (cnt > 0) out[i - 1] = (sum_mean == 0) ? s / (float)cnt : (sum_mean == 1) ? s : (sum_mean == 2) ? mx : 0.0f;
        else out[i - 1] = 0.0f;
    }
}

/* stream_find_to_corr, cuencas.f90:372-417: walk down DIR from (x,y) until a cell of the channel map (or the edge of the
 * map / a nodata direction) is reached and return that cell's centre.  `dire` changes inside the 3x3 scan, so one sweep
 * can take several steps; lat/lon stay unassigned (0 here) when the start cell matches no direction. */
void wo_stream_find_to_corr(const wo_geo *g, float x, float y, const int32_t *dir, const int32_t *cauce, int32_t nc, int32_t nf,
                            float *lat, float *lon)
{
    int32_t col, fil;
    wo_coord2fil_col(g, x, y, &col, &fil);
    *lat = 0.0f;
    *lon = 0.0f;
#define RM(a, c, f) a[(size_t)((f) - 1) * nc + ((c) - 1)]
    if (col < 1 || col > nc || fil < 1 || fil > nf) return;
    int32_t dire = RM(dir, col, fil);
    int flag = 1;
    while (flag == 1) {
        for (int kf = 1; kf <= 3; ++kf)
            for (int kc = 1; kc <= 3; ++kc) {
                const int32_t envia = 9 - 3 * kf + kc;
                if (dire == envia) {
                    col = col + kc - 2;
                    fil = fil + kf - 2;
                    if (col < 1 || col > nc || fil < 1 || fil > nf) { dire = -1; } /* the Fortran reads outside the map here */
                    else dire = RM(dir, col, fil);
                    *lat = g->xll + g->dx * ((float)col - 0.5f);
                    *lon = g->yll + g->dx * ((float)(g->nrows - fil) + 0.5f);
                }
            }
        if (col > nc || fil > nf || col <= 0 || fil <= 0 || (float)dire == g->nodata || dire <= 0 || RM(cauce, col, fil) != 0) flag = 0;
    }
#undef RM
}

/* find_xy_in_basin, cuencas.f90:2659-2678: 1-based position of (col,fil) in the basin table, 0 when outside */
static int32_t find_xy_in_basin(const int32_t *basin_f, int32_t col, int32_t fil, int32_t n)
{
    for (int32_t i = 0; i < n; ++i)
        if (BF(1, i) == col && BF(2, i) == fil) return i + 1;
    return 0;
}

/* basin_point2var, cuencas.f90:1230-1261 (basin_pts is never zeroed by the Fortran: zeros here) */
void wo_basin_point2var(const wo_geo *g, const int32_t *basin_f, const int32_t *id_coord, const float *xy_coord, int32_t ncoord,
                        int32_t n, int32_t *res_coord, int32_t *basin_pts)
{
    for (int32_t i = 0; i < n; ++i) basin_pts[i] = 0;
    for (int32_t i = 0; i < ncoord; ++i) {
        const float x = (xy_coord[2 * i] - g->xll) / g->dx;
        const int32_t x_col = (int32_t)ceilf(x);
        const float y = (float)g->nrows - (xy_coord[2 * i + 1] - g->yll) / g->dx;
        const int32_t y_fil = (int32_t)ceilf(y);
        const int32_t posit = find_xy_in_basin(basin_f, x_col, y_fil, n);
        if (posit > 0) { res_coord[i] = 0; basin_pts[posit - 1] = id_coord[i]; }
        else res_coord[i] = 1;
    }
}

/* basin_stream_point2stream, cuencas.f90:1524-1577 */
void wo_basin_stream_point2stream(const wo_geo *g, const int32_t *basin_f, const int32_t *cauce, const int32_t *id_coord,
                                  const float *xy_coord, int32_t ncoord, int32_t n, int32_t *res_coord, int32_t *basin_pts,
                                  float *xy_new)
{
    for (int32_t i = 0; i < n; ++i) basin_pts[i] = 0;
    for (int32_t i = 0; i < 2 * ncoord; ++i) xy_new[i] = g->nodata;
    for (int32_t i = 0; i < ncoord; ++i) {
        const float x = (xy_coord[2 * i] - g->xll) / g->dx;
        const int32_t x_col = (int32_t)ceilf(x);
        const float y = (float)g->nrows - (xy_coord[2 * i + 1] - g->yll) / g->dx;
        const int32_t y_fil = (int32_t)ceilf(y);
        int32_t posit = find_xy_in_basin(basin_f, x_col, y_fil, n);
        if (posit > 0) {
            res_coord[i] = 1;
            if (cauce[posit - 1] > 0) {
                res_coord[i] = 2;
                basin_pts[posit - 1] = id_coord[i];
            } else {
                while (cauce[posit - 1] == 0 && posit < n) posit = n + 1 - BF(0, posit - 1);
                if (posit < n) { basin_pts[posit - 1] = id_coord[i]; res_coord[i] = 2; }
            }
        } else {
            res_coord[i] = 0;
        }
        if (res_coord[i] == 2) {
            xy_new[2 * i] = g->xll + g->dx * ((float)BF(1, posit - 1) - 0.5f);
            xy_new[2 * i + 1] = g->yll + g->dx * ((float)(g->nrows - BF(2, posit - 1)) + 0.5f);
        }
    }
}

/* basin_stream_sections, cuencas.f90:1464-1523: for every channel cell a cross section of 2*num_celdas+1 raster cells
 * perpendicular-ish to its flow direction: elevations and the basin positions of those cells (0 = not in the basin,
 * -9999 = outside the map).  find_xy_in_basin is answered from a raster lookup built once (same result, first match). */
void wo_basin_stream_sections(const int32_t *basin_f, const int32_t *cauces, const int32_t *directions, const float *dem,
                              int32_t num_celdas, int32_t n, int32_t ncols, int32_t nrows, float *secciones, int32_t *secciones_cel)
{
    const int32_t S = 2 * num_celdas + 1;
    int32_t *look = (int32_t *)calloc((size_t)ncols * nrows, sizeof(int32_t));
    for (int32_t i = n - 1; i >= 0; --i) {
        const int32_t c = BF(1, i), f = BF(2, i);
        if (c >= 1 && c <= ncols && f >= 1 && f <= nrows) look[(size_t)(f - 1) * ncols + (c - 1)] = i + 1;
    }
    for (size_t k = 0; k < (size_t)S * n; ++k) { secciones[k] = 0.0f; secciones_cel[k] = 0; }
    for (int32_t i = 0; i < n; ++i) {
        if (cauces[i] != 1) continue;
        int32_t colMov = 0, filMov = 0;
        const int32_t d = directions[i];
        if (d == 4 || d == 6) filMov = 1;
        else if (d == 8 || d == 2) colMov = 1;
        else if (d == 7 || d == 3) { colMov = 1; filMov = -1; }
        else if (d == 1 || d == 9) { colMov = 1; filMov = 1; }
        const int32_t col = BF(1, i), fil = BF(2, i);
        int32_t cont = 0;
        for (int32_t j = -num_celdas; j <= num_celdas; ++j, ++cont) {
            const int32_t c = col + j * colMov, f = fil + j * filMov;
            if (c >= 1 && c <= ncols && f >= 1 && f <= nrows) {
                secciones[(size_t)S * i + cont] = dem[(size_t)(f - 1) * ncols + (c - 1)];
                secciones_cel[(size_t)S * i + cont] = look[(size_t)(f - 1) * ncols + (c - 1)];
            } else {
                secciones[(size_t)S * i + cont] = -9999.0f;
                secciones_cel[(size_t)S * i + cont] = -9999;
            }
        }
    }
    free(look);
}
