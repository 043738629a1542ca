/*
 * wmf_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A literal, sequential C restatement of the reference's Fortran for the two
 * hot paths (nicolas998/WMF: wmf/cuencas.f90 module `cu`, wmf/modelosv2.f90
 * module `models`).  Same loop order, float / int32_t arithmetic, powf from
 * libm, left-associated sums.  Build: -O2 -fno-fast-math -ffp-contract=off.
 *
 * PARITY PINNED against the reference's own compiled Fortran: oracle/_ref
 * (ref_loader.c + ref.py) runs the COFF objects the reference tree ships, and
 * tests/test_reference_pin.py requires this restatement to be BIT-IDENTICAL to
 * them on every routine below and every sub-model switch.  Also held by KAT-1
 * (tests/golden/kat1.json), invariants and closed forms (tests/test_oracle_*.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 *
 * Array layout follows f2py / Fortran column-major: X(k,N) is stored
 * k-fastest, i.e. element (k,i) (1-based) lives at [(i-1)*K + (k-1)].
 * Rasters M(nc,nr): element (col,row) (1-based) at [(row-1)*nc + (col-1)].
 */
#ifndef WMF_ORACLE_H
#define WMF_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* module cu globals, cuencas.f90:49-54 */
typedef struct {
    int32_t ncols, nrows;
    float xll, yll, dx, dy, dxp, nodata;
} wo_geo;

/* results written to module scalars by basin_basics/basin_acum, cuencas.f90:60-62 */
typedef struct {
    float area, pend_media, elevacion, centrox, centroy;
} wo_basin_scalars;

/* ---- Path A: cuencas.f90 ------------------------------------------------ */
void wo_coord2fil_col(const wo_geo *g, float x, float y, int32_t *col, int32_t *fil);
/* basin_temp must hold 3*ncols*nrows int32; returns nceldas */
int32_t wo_basin_find(const wo_geo *g, float x, float y, const int32_t *dir,
                      int32_t nc, int32_t nr, int32_t *basin_temp);
void wo_basin_cut(const wo_geo *g, const int32_t *basin_temp, int32_t nceldas, int32_t *basin_f);
void wo_basin_acum(const wo_geo *g, const int32_t *basin_f, int32_t n, int32_t *acum, float *area);
void wo_basin_acum_var(const int32_t *drain, int32_t n, const float *var, float *acum);
void wo_basin_basics(const wo_geo *g, const int32_t *basin_f, const float *dem, const int32_t *dir,
                     int32_t n, int32_t *acum, float *lng, float *pend, float *elev,
                     wo_basin_scalars *sc);
void wo_basin_coordxy(const wo_geo *g, const int32_t *basin_f, int32_t n, float *x, float *y);
void wo_basin_map2basin(const wo_geo *g, const int32_t *basin_f, int32_t n, const float *mapa,
                        float xllm, float yllm, int32_t ncolsm, int32_t nrowsm, float dxm, float dym,
                        float nodatam, float *vec);
void wo_basin_2map_find(const int32_t *basin, int32_t n, int32_t *map_ncols, int32_t *map_nrows);
void wo_basin_2map(const wo_geo *g, const int32_t *basin, const float *var, int32_t n,
                   int32_t map_ncols, int32_t map_nrows, float *mapa, float *map_xll, float *map_yll);
void wo_basin_float_var2map(const wo_geo *g, const int32_t *basin_f, const float *vect, int32_t n,
                            int32_t nc, int32_t nf, float *mapa);
void wo_basin_stream_nod(const int32_t *basin_f, const int32_t *acum, int32_t n, int32_t umbral,
                         int32_t *cauce, int32_t *nodos, int32_t *trazado, int32_t *n_nodos,
                         int32_t *n_cauce);
void wo_basin_stream_type(const int32_t *acum, int32_t n, const int32_t *umbrales, int32_t numbrales,
                          int32_t *stream_types);
void wo_geo_acum_to_cauce(const int32_t *acum, int32_t ncells, int32_t umbral, int32_t *cauce);
void wo_geo_hand(const int32_t *basin_f, const float *elev, const float *lng, const int32_t *cauce,
                 int32_t n, float *hand, float *hdnd, int32_t *a_quien);
void wo_geo_hand_global(const wo_geo *g, const float *dem, const int32_t *dir, const int32_t *red,
                        int32_t nc, int32_t nr, float *hand);
void wo_basin_time_to_out(const int32_t *basin_f, const float *lng, const float *speed, int32_t n,
                          float *time);
void wo_basin_propagate(const int32_t *basin_f, const float *var, int32_t n, float *var_prop);

/* ---- Path B: modelosv2.f90 ---------------------------------------------- */
typedef struct {
    /* sizes */
    int32_t n_cel, n_reg, n_cont, n_conth;
    /* scalar config (module globals modelosv2.f90:71-97) */
    float dt, dxp, retorno_gr, retorno_aq, storage_constant;
    int32_t speed_type[3];
    int32_t calc_niter;
    int32_t sim_sediments, sim_slides;
    int32_t show_storage, show_speed, show_mean_speed, show_mean_retorno, save_retorno;
    int32_t save_vfluxes, save_rc;
    int32_t separate_rain;  /* module separate_rain (:91): shadow storages by rain type, Qsep_byrain */
    int32_t separate_fluxes;   /* modelosv2.f90:90, shadow shares of runoff / subsurface / base flow in the channel */
    float calib[11];
    /* topology + statics, all (1,N) or (K,N) column-major */
    const int32_t *drena;      /* (N) drain ids = structure row 1 */
    const int32_t *unit_type;  /* (N) */
    const float *hill_long, *hill_slope, *stream_long, *stream_slope, *elem_area; /* (N) */
    const float *v_coef, *h_coef, *h_exp;  /* (4,N) */
    const float *max_capilar, *max_gravita, *max_aquifer; /* (N) */
    const float *evpserie;     /* (n_reg) */
    const float *storage;      /* (5,N) module Storage */
    const int32_t *control, *control_h; /* (N) */
    const float *stoin;        /* (5,N) or NULL */
    const float *hspeedin;     /* (4,N) or NULL */
    /* rain held in memory: records are 1-based, record r at rain + (r-1)*N;
       posEvento(t)==1 means "no rain" (modelosv2.f90:444-445) */
    const int32_t *rain;       /* (n_records, N) int32 mm*1000 */
    int32_t n_records;
    const int32_t *pos_evento; /* (n_reg) */
    /* sediments (modelosv2.f90:115-125) */
    float sed_factor, wi[3], diametro[3];
    const float *krus, *crus, *prus; /* (N) */
    const float *parliac;      /* (3,N) */
    float *vd_init;            /* (3,N) or NULL: Vd persists across runs (sed_allocate :1301-1304) */
    /* slides (modelosv2.f90:135-150) */
    int32_t sl_gullienogullie;
    float sl_fs, sl_gammaw;
    const float *sl_zs, *sl_gammas, *sl_cohesion, *sl_frictionangle, *sl_radslope; /* (N) */
    /* separate_rain (:346-359, :452-456): rain-type tables like `rain` (int32, record r at conv + (r-1)*N, a cell is
       convective / stratiform when its value / 1000 is non-zero) and the record of every step (n_reg, 1-based) */
    const int32_t *conv, *stra;
    int32_t n_conv_records, n_stra_records;
    const int32_t *pos_conv, *pos_stra;
    /* HydroFlash flood sub-model (modelosv2.f90:728-743, 1711-1822): module switches and maps */
    int32_t sim_floods, flood_sec_tam;  /* flood_sec_tam = rows of flood_sections */
    float flood_dw, flood_dsed, flood_av, flood_cmax, flood_umbral, flood_step, flood_hmax;
    const float *flood_w, *flood_d50, *flood_slope;   /* (N) */
    const float *flood_sections, *flood_sec_cells;    /* (flood_sec_tam, N): cross-section elevations and the cells they lie on */
} wo_shia_in;

typedef struct {
    float *q;           /* (n_cont, n_reg) column-major: [ (t)*n_cont + c ] */
    float *qsed;        /* (n_cont, 3, n_reg) */
    float *qseparated;  /* (n_cont, 3, n_reg), separate_fluxes = 1 (may be NULL) */
    float *hum, *st1, *st3; /* (n_conth, n_reg) */
    float *balance;     /* (n_reg) literal fp32 sequential sums */
    double *balance64;  /* (n_reg) same terms accumulated in double (diagnostic, may be NULL) */
    float *speed, *areacontrol; /* (n_cont, n_reg) */
    float *stoout;      /* (5,N) */
    float *hspeed_out;  /* (4,N) final horizontal speeds (Speed_map analogue) */
    float *mean_rain;   /* (n_reg) */
    float *acum_rain;   /* (N) */
    float *mean_storage;/* (5,n_reg) or NULL */
    float *mean_speed;  /* (4,n_reg) or NULL */
    float *mean_retorno;/* (n_reg) or NULL */
    float *retorned;    /* (N) or NULL */
    float *mean_vfluxes;/* (4,n_reg) or NULL */
    float *rc_coef;     /* (2,N) or NULL */
    /* fp64-accumulated twins of the three per-step means (diagnostic, may be NULL): the reference's
       sequential fp32 sums carry an error of order N*eps, these do not */
    double *mean_storage64, *mean_speed64, *mean_retorno64;
    double *mean_vfluxes64;  /* (4,n_reg) fp64 twin of mean_vfluxes (may be NULL) */
    float *vfluxes_last;     /* (4,N) the vfluxes map after the last step (what a dump of that step holds), may be NULL */
    float *retorned_last;    /* (N) Retorned after the last step, before the per-step reset (:841-843) = the record save_retorno writes for that step; may be NULL */
    /* sediments */
    float *vs, *vd, *vsc, *vdc; /* (3,N) */
    float *volero, *voldepo;    /* (N) */
    float *erot, *dept;         /* (3) */
    double *erot64, *dept64;    /* (3) fp64-accumulated twins (diagnostic, may be NULL) */
    /* slides */
    float *sl_zmin, *sl_zcrit, *sl_zmax, *sl_bo, *sl_riskvector, *sl_slideocurrence; /* (N) */
    int32_t *sl_slideacumulate; /* (N) */
    int32_t *sl_slidencelltime; /* (n_reg) */
    float *qsep_byrain;      /* (n_cont,2,n_reg) when separate_rain (:713-715, :767-769), may be NULL */
    float *storage_conv, *storage_stra; /* (5,N) final shadow storages (module Storage_conv / Storage_stra), may be NULL */
    /* floods: module flood_flood (N) int, flood_h / ufr / Cr / Q / Qsed / speed (N), flood_profundidad (sec_tam,N),
       flood_scalars = {flood_rdf, flood_area, flood_diff} after the last evaluation; all may be NULL */
    int32_t *flood_flood;
    float *flood_h, *flood_ufr, *flood_cr, *flood_q, *flood_qsed, *flood_speed, *flood_profundidad, *flood_scalars;
} wo_shia_out;

int wo_shia_v1(const wo_shia_in *in, wo_shia_out *out);
/* modelosv2.f90:1278-1293; returns area, updates *speed */
float wo_calc_speed(float sm, float coef, float expo, float elem_long, float *speed, float dt,
                    int32_t calc_niter);

/* basin_point2var :1230-1261, basin_stream_point2stream :1524-1577 (control points onto the basin / the stream network) */
void wo_basin_point2var(const wo_geo *g, const int32_t *basin_f, const int32_t *id_coord, const float *xy_coord, int32_t ncoord,
                        int32_t n, int32_t *res_coord, int32_t *basin_pts);
void wo_basin_stream_point2stream(const wo_geo *g, const int32_t *basin_f, const int32_t *cauce, const int32_t *id_coord,
                                  const float *xy_coord, int32_t ncoord, int32_t n, int32_t *res_coord, int32_t *basin_pts,
                                  float *xy_new);
/* basin_stream_sections, cuencas.f90:1464-1523 */
void wo_basin_stream_sections(const int32_t *basin_f, const int32_t *cauces, const int32_t *directions, const float *dem,
                              int32_t num_celdas, int32_t n, int32_t ncols, int32_t nrows, float *secciones, int32_t *secciones_cel);
/* stream_find_to_corr, cuencas.f90:372-417 */
void wo_stream_find_to_corr(const wo_geo *g, float x, float y, const int32_t *dir, const int32_t *cauce, int32_t nc, int32_t nf,
                            float *lat, float *lon);
/* rain_pre_mit, modelosv2.f90:1068-1101 */
void wo_rain_pre_mit(const float *xy_basin, const int32_t *tin, const float *coord, int32_t nceldas, int32_t ntin,
                     int32_t *tin_perte);
/* sub-basin (hills) topology: cuencas.f90:1380-1438, 1835-2118 (see wmf_oracle_cu.c for the quirks reproduced) */
void wo_basin_stream_slope(const wo_geo *g, const int32_t *basin_f, const float *elev, const float *lng,
                           const int32_t *nodos, int32_t n_cauce, int32_t n, float *stream_s, float *stream_l);
void wo_basin_subbasin_nod(const int32_t *basin_f, const int32_t *acum, int32_t n, int32_t umbral, int32_t *cauce,
                           int32_t *nodos_fin, int32_t *n_nodos, int32_t *sub_basins /* 2*n ints */);
void wo_basin_subbasin_find(const int32_t *basin_f, const int32_t *nodos, int32_t n, int32_t *sub_pert);
void wo_basin_subbasin_horton(const int32_t *sub_basins, int32_t n_nodos, int32_t n, int32_t *sub_horton, int32_t *nod_horton);
void wo_basin_subbasin_long(const int32_t *sub_pert, const int32_t *cauce, const int32_t *lng_int, const int32_t *sub_basin,
                            const int32_t *sub_horton, int32_t n_nodos, int32_t n, float *sub_basin_long, float *max_long,
                            int32_t *nodo_max_long);
void wo_basin_subbasin_stream_prop(const int32_t *sub_pert, const int32_t *cauce, const float *lng, const float *slope,
                                   int32_t n, int32_t n_nodos, float *stream_slope, float *stream_long);
void wo_basin_subbasin_map2subbasin(const int32_t *sub_pert, const float *var, int32_t n_nodos, int32_t n,
                                    const int32_t *cauce, int32_t sum_mean, float *out);

/* rain_idw (modelosv2.f90:1195-1271) and rain_mit (:1104-1194), cells model (maskVector all ones).  Instead of writing
 * records to `ruta`, the wet fields go to table (capacity rows x nceldas int32, mm*1000 truncated): row 0 = the all-zero
 * record 1, row k-1 = record k.  mean_rain / pos_ids (nreg) as the Fortran returns them; mean_rain64 (may be NULL) is the
 * fp64 twin of the sequential fp32 sum.  Returns the number of records written (>= 1), or -1 when capacity is too small. */
int32_t wo_rain_idw(const float *xy_basin, const float *coord, const float *rain, float pp, int32_t nceldas,
                    int32_t ncoord, int32_t nreg, float umbral, int32_t *table, int32_t capacity, float *mean_rain,
                    int32_t *pos_ids, double *mean_rain64, const int32_t *mask /* hill of every cell or NULL */, int32_t nhills);
int32_t wo_rain_mit(const float *xy_basin, const float *coord, const float *rain, const int32_t *tin,
                    const int32_t *tin_perte, int32_t nceldas, int32_t ncoord, int32_t ntin, int32_t nreg, float umbral,
                    int32_t *table, int32_t capacity, float *mean_rain, int32_t *pos_ids, double *mean_rain64,
                    const int32_t *mask, int32_t nhills);

#ifdef __cplusplus
}
#endif
#endif
