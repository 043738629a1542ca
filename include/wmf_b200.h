/*
 * wmf_b200.h -- C ABI of libwmf_b200.so: the B200 (sm_100a) drop-in for the two
 * data-parallel hot paths of nicolas998/WMF.
 *
 * What this replaces.  The reference reaches its Fortran through two f2py
 * extension modules (`cu` from wmf/cuencas.f90, `models` from
 * wmf/modelosv2.f90; setup.py:5-8, generated wrappers
 * build/src.win-amd64-3.7/{cumodule.c,modelsmodule.c}).  Every entry point
 * below names the f2py-wrapped Fortran subroutine it stands in for.  The
 * Python side (wmf_b200/_cu.py, wmf_b200/_models.py) loads this library with
 * ctypes and re-creates the `cu` / `models` module objects with the reference's
 * positional signatures.
 *
 * Conventions
 *  - Plain C: opaque context handle, pointers and sizes only.
 *  - Every array pointer may be a HOST pointer or a DEVICE pointer of the
 *    context's GPU (detected with cudaPointerGetAttributes).  Host inputs are
 *    staged host->device and host outputs device->host on the context stream
 *    inside the call; device pointers are used in place (zero copies), which is
 *    how resident pipelines and the benchmark's kernel-only number run.
 *  - Layouts are the Fortran/f2py ones: X(k,N) is k-fastest
 *    ([(i-1)*K + (k-1)]); a raster M(nc,nr) stores (col,row) at
 *    [(row-1)*nc + (col-1)]; cell ids, columns and rows are 1-based.
 *  - Every function returns a status (0 = WMF_OK) and never throws; the text of
 *    the last error is available from wmf_last_error().  The Fortran printed
 *    and carried on instead (cuencas.f90:605, modelosv2.f90:889).
 *  - There is NO CPU fallback: without a CUDA device wmf_ctx_create fails.
 *  - A context is not thread-safe; use one per thread / per GPU.  (The Fortran
 *    modules are process-global and non-reentrant.)
 */
#ifndef WMF_B200_H
#define WMF_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WMF_ABI_VERSION 14

enum {
    WMF_OK = 0,
    WMF_ERR_ARG = 1,      /* bad size / null pointer / unsupported switch */
    WMF_ERR_CUDA = 2,     /* CUDA runtime error, see wmf_last_error */
    WMF_ERR_TOPOLOGY = 3, /* drain table is cyclic, not upstream-first or not level-ordered */
    WMF_ERR_STATE = 4,    /* call order violated (e.g. basin_cut before basin_find) */
    WMF_ERR_NOMEM = 5
};

typedef struct wmf_ctx wmf_ctx;

/* ------------------------------------------------------------------ context */
int wmf_abi_version(void);
int wmf_ctx_create(int device, wmf_ctx **out);
int wmf_ctx_destroy(wmf_ctx *ctx);
const char *wmf_last_error(const wmf_ctx *ctx);
/* Run on an existing CUDA stream (e.g. torch.cuda.current_stream().cuda_stream); NULL = own stream. */
int wmf_set_stream(wmf_ctx *ctx, void *cuda_stream);
int wmf_synchronize(wmf_ctx *ctx);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
int64_t wmf_launch_count(const wmf_ctx *ctx);

/* module `cu` globals (cuencas.f90:49-54), set by wmf.read_map_raster (wmf.py:319-333) */
int wmf_set_geo(wmf_ctx *ctx, int32_t ncols, int32_t nrows, float xll, float yll, float dx, float dy,
                float dxp, float nodata);
/* module `cu` result scalars (cuencas.f90:60-62): area, pend_media, elevacion, centrox, centroy */
int wmf_get_basin_scalars(wmf_ctx *ctx, float out5[5]);

/* ------------------------------------------------------------ Path A: `cu` */
/* coord2fil_col, cuencas.f90:77-90 (host arithmetic in fp32, bit-exact) */
int wmf_coord2fil_col(wmf_ctx *ctx, float x, float y, int32_t *col, int32_t *fil);
/* basin_find, cuencas.f90:525-591.  dir: int32 (nc,nr).  Keeps the BFS queue in the context
 * (the reference keeps it in the module global basin_temp). */
int wmf_basin_find(wmf_ctx *ctx, float x, float y, const int32_t *dir, int32_t nc, int32_t nr,
                   int32_t *nceldas);
/* basin_cut, cuencas.f90:592-607.  basin_f: int32 (3,nceldas) */
int wmf_basin_cut(wmf_ctx *ctx, int32_t nceldas, int32_t *basin_f);
/* number of BFS levels (tree depth + 1) of the last basin_find; diagnostic */
int wmf_basin_find_depth(wmf_ctx *ctx, int32_t *depth);
/* basin_acum, cuencas.f90:659-680 (also sets the `area` scalar) */
int wmf_basin_acum(wmf_ctx *ctx, const int32_t *basin_f, int32_t nceldas, int32_t *acum);
/* basin_acum_var, cuencas.f90:681-701.  drain: int32 (nceldas) = basin_f row 1 */
int wmf_basin_acum_var(wmf_ctx *ctx, const int32_t *drain, int32_t nceldas, const float *var, float *acum);
/* basin_basics, cuencas.f90:608-658 (sets all five scalars) */
int wmf_basin_basics(wmf_ctx *ctx, const int32_t *basin_f, const float *dem, const int32_t *dir,
                     int32_t nceldas, int32_t *acum, float *lng, float *pend, float *elev);
/* basin_coordXY, cuencas.f90:797-808 */
int wmf_basin_coordxy(wmf_ctx *ctx, const int32_t *basin_f, int32_t nceldas, float *x, float *y);
/* basin_map2basin, cuencas.f90:1144-1184.  mapa: fp32 (ncolsm,nrowsm) */
int wmf_basin_map2basin(wmf_ctx *ctx, const int32_t *basin_f, int32_t nceldas, const float *mapa,
                        float xllm, float yllm, int32_t ncolsm, int32_t nrowsm, float dxm, float dym,
                        float nodatam, float *vec);
/* basin_2map_find, cuencas.f90:1185-1201 */
int wmf_basin_2map_find(wmf_ctx *ctx, const int32_t *basin, int32_t nceldas, int32_t *map_ncols,
                        int32_t *map_nrows);
/* basin_2map, cuencas.f90:1202-1229.  mapa: fp32 (map_ncols,map_nrows) */
int wmf_basin_2map(wmf_ctx *ctx, const int32_t *basin, const float *var, int32_t nceldas,
                   int32_t map_ncols, int32_t map_nrows, float *mapa, float *map_xll, float *map_yll);
/* basin_float_var2map, cuencas.f90:1037-1051.  mapa: fp32 (nc,nf) */
int wmf_basin_float_var2map(wmf_ctx *ctx, const int32_t *basin_f, const float *vect, int32_t nceldas,
                            int32_t nc, int32_t nf, float *mapa);
/* channel mask of basin_stream_nod, cuencas.f90:1351-1353: cauce = (acum >= umbral) */
int wmf_basin_stream_mask(wmf_ctx *ctx, const int32_t *acum, int32_t nceldas, int32_t umbral,
                          int32_t *cauce, int32_t *n_cauce);
/* basin_stream_nod, cuencas.f90:1340-1379: channel mask, hydrological nodes (3 = head, 2+ = confluence / outlet), tracing
 * points */
int wmf_basin_stream_nod(wmf_ctx *ctx, const int32_t *basin_f, const int32_t *acum, int32_t nceldas, int32_t umbral,
                         int32_t *cauce, int32_t *nodos, int32_t *trazado, int32_t *n_nodos, int32_t *n_cauce);
/* basin_stream_type, cuencas.f90:1439-1463 */
int wmf_basin_stream_type(wmf_ctx *ctx, const int32_t *acum, int32_t nceldas, const int32_t *umbrales,
                          int32_t numbrales, int32_t *stream_types);
/* geo_acum_to_cauce, cuencas.f90:2120-2129 (strict >) over ncells raster cells */
int wmf_geo_acum_to_cauce(wmf_ctx *ctx, const int32_t *acum, int64_t ncells, int32_t umbral, int32_t *cauce);
/* geo_hand, cuencas.f90:2131-2174 */
int wmf_geo_hand(wmf_ctx *ctx, const int32_t *basin_f, const float *elev, const float *lng,
                 const int32_t *cauce, int32_t nceldas, float *hand, float *hdnd, int32_t *a_quien);
/* geo_hand_global, cuencas.f90:2175-2219.  rasters (nc,nr) */
int wmf_geo_hand_global(wmf_ctx *ctx, const float *dem, const int32_t *dir, const int32_t *red,
                        int32_t nc, int32_t nr, float *hand);
/* basin_time_to_out, cuencas.f90:724-756 */
int wmf_basin_time_to_out(wmf_ctx *ctx, const int32_t *basin_f, const float *lng, const float *speed,
                          int32_t nceldas, float *time);
/* basin_propagate, cuencas.f90:1811-1833 */
int wmf_basin_propagate(wmf_ctx *ctx, const int32_t *basin_f, const float *var, int32_t nceldas,
                        float *var_prop);

/* basin_point2var (cuencas.f90:1230-1261) and basin_stream_point2stream (:1524-1577): control points onto the basin table /
 * the stream network (set_Control -> Points_Points2Stream, wmf.py:2096-2103, 4021).  xy_coord (2,ncoord), xy_new (2,ncoord). */
int wmf_basin_point2var(wmf_ctx *ctx, const int32_t *basin_f, const int32_t *id_coord, const float *xy_coord,
                        int32_t ncoord, int32_t nceldas, int32_t *res_coord, int32_t *basin_pts);
int wmf_basin_stream_point2stream(wmf_ctx *ctx, const int32_t *basin_f, const int32_t *cauce, const int32_t *id_coord,
                                  const float *xy_coord, int32_t ncoord, int32_t nceldas, int32_t *res_coord,
                                  int32_t *basin_pts, float *xy_new);
/* basin_stream_sections, cuencas.f90:1464-1523: the cross sections HydroFlash reads (GetGeo_Sections, wmf.py:1564-1587):
 * secciones / secciones_cel (2*num_celdas+1, nceldas); dem is the (ncols,nrows) raster */
int wmf_basin_stream_sections(wmf_ctx *ctx, const int32_t *basin_f, const int32_t *cauces, const int32_t *directions,
                              const float *dem, int32_t num_celdas, int32_t nceldas, int32_t ncols, int32_t nrows,
                              float *secciones, int32_t *secciones_cel);
/* stream_find_to_corr, cuencas.f90:372-417 (SimuBasin.__init__ with useCauceMap, wmf.py:2998-3000) */
int wmf_stream_find_to_corr(wmf_ctx *ctx, float x, float y, const int32_t *dir, const int32_t *cauce, int32_t nc,
                            int32_t nf, float *lat, float *lon);

/* Sub-basin (hills) topology: the routines SimuBasin.__init__ / set_Geomorphology call next to the basin cut
 * (wmf.py:3019-3023, 3664-3676).  Layouts as f2py: sub_basins (2,n_nodos), vectors (nceldas) or (n_nodos). */
/* basin_subbasin_nod, cuencas.f90:1835-1920; keeps sub_basins_temp in the context for basin_subbasin_cut (:1921-1930) */
int wmf_basin_subbasin_nod(wmf_ctx *ctx, const int32_t *basin_f, const int32_t *acum, int32_t nceldas, int32_t umbral,
                           int32_t *cauce, int32_t *nodos_fin, int32_t *n_nodos);
int wmf_basin_subbasin_cut(wmf_ctx *ctx, int32_t n_nodos, int32_t *sub_basins);
/* basin_subbasin_find, cuencas.f90:1979-2010 (its second output sub_basin is never assigned by the Fortran) */
int wmf_basin_subbasin_find(wmf_ctx *ctx, const int32_t *basin_f, const int32_t *nodos, int32_t nceldas, int32_t *sub_pert);
/* basin_subbasin_horton, cuencas.f90:1931-1978 */
int wmf_basin_subbasin_horton(wmf_ctx *ctx, const int32_t *sub_basins, int32_t n_nodos, int32_t nceldas,
                              int32_t *sub_horton, int32_t *nod_horton);
/* basin_subbasin_long, cuencas.f90:2011-2051 (`lng` is truncated to integer as the Fortran's INTEGER dummy does) */
int wmf_basin_subbasin_long(wmf_ctx *ctx, const int32_t *sub_pert, const int32_t *cauce, const float *lng,
                            const int32_t *sub_basin, const int32_t *sub_horton, int32_t n_nodos, int32_t nceldas,
                            float *sub_basin_long, float *max_long, int32_t *nodo_max_long);
/* basin_subbasin_stream_prop, cuencas.f90:2093-2115 */
int wmf_basin_subbasin_stream_prop(wmf_ctx *ctx, const int32_t *sub_pert, const int32_t *cauce, const float *lng,
                                   const float *slope, int32_t nceldas, int32_t n_nodos, float *stream_slope,
                                   float *stream_long);
/* basin_subbasin_map2subbasin, cuencas.f90:2052-2092 (sum_mean: 0 mean, 1 sum, 2 max over the hill's channel cells) */
int wmf_basin_subbasin_map2subbasin(wmf_ctx *ctx, const int32_t *sub_pert, const float *basin_var, int32_t n_nodos,
                                    int32_t nceldas, const int32_t *cauce, int32_t sum_mean, float *subbasin_sum);
/* basin_stream_slope, cuencas.f90:1380-1438 (uses the context's nodata and dxp) */
int wmf_basin_stream_slope(wmf_ctx *ctx, const int32_t *basin_f, const float *basin_elev, const float *basin_long,
                           const int32_t *nodos, int32_t nceldas, float *stream_s, float *stream_l);

/* -------------------------------------------------------- Path B: `models` */
/* Scalar switches = module `models` globals (modelosv2.f90:71-97) + shia_v1 sizes (:191-216). */
typedef struct {
    int32_t n_cel, n_reg, n_cont, n_conth;
    float dt, dxp, retorno_gr, retorno_aq, storage_constant;
    int32_t speed_type[3];
    int32_t calc_niter;
    int32_t sim_sediments, sim_slides;
    int32_t show_storage, show_speed, show_mean_speed, show_mean_retorno, save_retorno;
    /* sediments (modelosv2.f90:115-125, wmf.py:4046-4104) */
    float sed_factor, wi[3], diametro[3];
    /* slides (modelosv2.f90:135-150, wmf.py:4106-4163) */
    int32_t sl_gullienogullie;
    float sl_fs, sl_gammaw;
    /* ensemble: number of members sharing topology, maps and rain (1 = the reference's single run) */
    int32_t n_members;
    /* state dumps (modelosv2.f90:799-803, :815-818): the library returns the snapshots, the caller writes the files */
    int32_t save_storage, save_speed;
    /* module separate_fluxes (modelosv2.f90:90): track the runoff / subsurface / base-flow shares of the channel water */
    int32_t separate_fluxes;
    /* module save_vfluxes / save_rc (modelosv2.f90:83, :392-426): vertical-flux dumps + mean_vfluxes, and the two
     * runoff-coefficient accumulators; save_retorno (above) returns the per-step return-flow map */
    int32_t save_vfluxes, save_rc;
    /* module separate_rain (modelosv2.f90:91): convective / stratiform shadow storages and Qsep_byrain */
    int32_t separate_rain;
    /* HydroFlash flood sub-model (modelosv2.f90:79, :160-183, :728-743, :1711-1822): module sim_floods and flood_* scalars;
     * flood_sec_tam = rows of flood_sections */
    int32_t sim_floods, flood_sec_tam;
    float flood_dw, flood_dsed, flood_av, flood_cmax, flood_umbral, flood_step, flood_hmax;
    /* storm-scenario ensembles (BASELINE configs[4]; the reference forks one process per scenario, wmf.py:872-878): when 1,
     * pos_evento is (n_members, n_reg) row-major and member m reads the rain record pos_evento[m][t] of the shared table */
    int32_t pos_per_member;
} wmf_shia_cfg;

/* Arrays = module `models` allocatables (modelosv2.f90:53-113) in their f2py layouts. */
typedef struct {
    const int32_t *drena;     /* (N) drain ids = structure row 1 */
    int32_t drena_stride;     /* 1, or 3 when `drena` points at a (3,N) structure table */
    const int32_t *unit_type; /* (N) */
    const float *hill_long, *hill_slope, *stream_long, *stream_slope, *elem_area; /* (N) */
    const float *v_coef, *h_coef, *h_exp; /* (4,N) */
    const float *max_capilar, *max_gravita, *max_aquifer; /* (N) */
    const float *evpserie;    /* (n_reg) */
    const float *storage;     /* (5,N) module Storage (used when stoin is NULL or stoin[0] <= 0) */
    const int32_t *control, *control_h; /* (N), may be NULL */
    const float *calib;       /* (11) ; ensembles: (n_members, 11) row-major */
    const float *stoin;       /* (5,N) or NULL   -- shia_v1 optional StoIn   */
    const float *hspeedin;    /* (4,N) or NULL   -- shia_v1 optional HspeedIn */
    /* rainfall held in memory instead of the per-step direct-access read (modelosv2.f90:444-449):
     * record r (1-based) at rain + (r-1)*N, int32 mm*1000; pos_evento(t)==1 means "dry step". */
    const int32_t *rain;
    int32_t n_records;
    const int32_t *pos_evento; /* (n_reg) */
    /* sediments */
    const float *krus, *crus, *prus; /* (N) */
    const float *parliac;            /* (3,N) */
    const float *vd_init;            /* (3,N) or NULL: Vd carried over from a previous run (:1301-1304) */
    /* slides */
    const float *sl_zs, *sl_gammas, *sl_cohesion, *sl_frictionangle, *sl_radslope; /* (N) */
    /* module guarda_cond (n_reg): > 0 = dump StoOut after that step (the value is the record number the caller writes
     * it to, wmf.py:4497-4505); needed when save_storage = 1 */
    const int32_t *guarda_cond;
    /* module guarda_vfluxes (n_reg), same convention, for save_vfluxes = 1; NULL = no step is dumped (:399-401) */
    const int32_t *guarda_vfluxes;
    /* separate_rain = 1: the rain-type maps, held in memory like `rain` (record r at conv + (r-1)*N, int32; a cell is
     * convective / stratiform at a step when value / 1000 -- integer division -- is non-zero, modelosv2.f90:471-474) and
     * the record of every step, posConv / posStra (n_reg, 1-based; read on wet steps only, :452-456) */
    const int32_t *conv, *stra;
    int32_t n_conv_records, n_stra_records;
    const int32_t *pos_conv, *pos_stra;
    /* sim_floods = 1: module flood_w, flood_d50, flood_slope (N); flood_sections, flood_sec_cells (flood_sec_tam, N) */
    const float *flood_w, *flood_d50, *flood_slope, *flood_sections, *flood_sec_cells;
} wmf_shia_inputs;

/* Outputs.  Any pointer may be NULL (= not wanted).  Shapes as returned by f2py (modelsmodule.c:375-408).
 * For ensembles (n_members = M > 1) q is (M, n_cont, n_reg) with the member slowest, stoout /
 * hspeed_out are (M, 5|4, N); the remaining per-run outputs must be NULL. */
typedef struct {
    float *q;            /* (n_cont, n_reg) */
    float *qsed;         /* (n_cont, 3, n_reg) */
    float *hum, *st1, *st3; /* (n_conth, n_reg) */
    float *balance;      /* (n_reg); accumulated in fp64 on the GPU, rounded once */
    float *speed, *areacontrol; /* (n_cont, n_reg) */
    float *stoout;       /* (5,N) */
    float *hspeed_out;   /* (4,N) final horizontal speeds (feeds HspeedIn of the next run) */
    float *mean_rain;    /* (n_reg)  module Mean_Rain */
    float *acum_rain;    /* (N)      module Acum_rain */
    float *mean_storage; /* (5,n_reg) when show_storage */
    float *mean_speed;   /* (4,n_reg) when show_mean_speed */
    float *mean_retorno; /* (n_reg)   when retorno_gr & show_mean_retorno */
    float *retorned;     /* (N)       when retorno_gr */
    float *vs, *vd, *vsc, *vdc; /* (3,N) sediments */
    float *volero, *voldepo;    /* (N) */
    float *erot, *dept;         /* (3) */
    float *sl_zmin, *sl_zcrit, *sl_zmax, *sl_bo, *sl_riskvector, *sl_slideocurrence; /* (N) */
    int32_t *sl_slideacumulate; /* (N) */
    int32_t *sl_slidencelltime; /* (n_reg) */
    /* save_storage = 1: StoOut after every step with guarda_cond > 0, in step order: (n_snap, 5, N) -- what
     * write_float_basin(ruta_storage, StoOut, guarda_cond(t), N_cel, 5) puts on disk (modelosv2.f90:799-803) */
    float *sto_snap;
    /* save_speed = 1: hspeed after every step: (n_reg, 4, N) (modelosv2.f90:815-818) */
    float *speed_snap;
    /* separate_fluxes = 1: Qseparated (n_cont, 3, n_reg), modelosv2.f90:697-704, 752-759 */
    float *qseparated;
    /* save_vfluxes = 1: mean_vfluxes (4, n_reg) = sum(vfluxes,dim=2)/N_cel of every step (:807-809; fp64 sums on the
     * GPU, rounded once) and the vfluxes map (4,N) of every step with guarda_vfluxes > 0, in step order:
     * (n_vsnap, 4, N) -- what write_float_basin(ruta_vfluxes, vfluxes, guarda_vfluxes(t), N_cel, 4) writes (:811-813) */
    float *mean_vfluxes, *vfl_snap;
    /* save_rc = 1: rc_coef (2,N) after the final clamp rc(1) = min(rc(1), rc(2)) (:521-524, :863-867) */
    float *rc_coef;
    /* save_retorno = 1: Retorned of every step, (n_reg, N) -- record t of ruta_retorno (:820-823) */
    float *ret_snap;
    /* separate_rain = 1: Qsep_byrain (n_cont, 2, n_reg) (modelosv2.f90:713-716, :767-770) and the final shadow storages,
     * module Storage_conv / Storage_stra (5,N) */
    float *qsep_byrain, *storage_conv, *storage_stra;
    /* sim_floods = 1: module flood_flood (N) int32 (0 dry, 1 flooded by a section, 2 evaluated channel cell), the maps of
     * the evaluated cells' last evaluation flood_h / ufr / Cr / Q / Qsed / speed (N), flood_profundidad (flood_sec_tam, N)
     * and flood_scalars[3] = {flood_rdf, flood_area (never assigned by the Fortran: 0), flood_diff} of the last evaluation */
    int32_t *flood_flood;
    float *flood_h, *flood_ufr, *flood_cr, *flood_q, *flood_qsed, *flood_speed, *flood_profundidad, *flood_scalars;
} wmf_shia_outputs;

/* shia_v1, modelosv2.f90:191-869 (+ calc_speed :1278-1293, sed_* :1298-1608, slide_* :1613-1707). */
int wmf_shia_run(wmf_ctx *ctx, const wmf_shia_cfg *cfg, const wmf_shia_inputs *in, wmf_shia_outputs *out);

/* Objective scores of an ensemble against one observed series: wmf.py:749-754 (__eval_nash__) and
 * wmf.py:775-781 (__eval_q_pico__).  q: (M, n_cont, n_reg) as produced by wmf_shia_run, row = 0-based
 * control row to score (wmf.py:875 uses Qsim[nodo]); scores: (M,2) = [nash, qpico]. */
int wmf_eval_scores(wmf_ctx *ctx, const float *q, int32_t n_members, int32_t n_cont, int32_t n_reg,
                    int32_t row, const float *qobs, float *scores);

/* Rainfall interpolation, cells model: rain_idw (modelosv2.f90:1195-1271) and rain_mit (:1104-1194).
 * xy_basin (2,nceldas), coord (2,ncoord), rain (ncoord,nreg) in their f2py layouts; tin (3,ntin) 1-based station ids and
 * tin_perte (nceldas) 1-based triangle of every cell for the TIN variant.  Instead of writing records to `ruta` the wet
 * fields go to `table` (capacity rows x nceldas int32, mm*1000 truncated; host or device): row 0 = the all-zero record 1,
 * row k-1 = record k -- exactly the table wmf_shia_inputs.rain takes, with pos_ids as pos_evento.  mean_rain / pos_ids
 * (nreg) as the Fortran returns them (the field sums are accumulated in fp64).  *n_records = rows of the finished table.
 * While interpolating the table must hold 1 + (number of steps with a positive reading) rows -- capacity = nreg + 1 always
 * suffices; with table = NULL only the statistics are computed. */
/* rain_pre_mit, modelosv2.f90:1068-1101: tin_perte (nceldas) = 1-based triangle of every cell, 0 = none */
int wmf_rain_pre_mit(wmf_ctx *ctx, const float *xy_basin, const int32_t *tin, const float *coord, int32_t nceldas,
                     int32_t ntin, int32_t ncoord, int32_t *tin_perte);
/* rain_idw also has the hills model: mask_vector (nceldas) = hill of every cell (1..nhills), nhills > 0: the records then hold
 * nhills values, column i = hill nhills-i+1 = the SUM of the hill's cell values (what the Fortran computes, :1214-1216,
 * :1257-1259 with module models' basin_subbasin_map2subbasin(..., sum_mean = 2)).  mask_vector = NULL, nhills = 0: cells.
 * (rain_mit's hills branch needs sum(maskVector) == nhills, :1126, which no hill table satisfies: cells only.) */
int wmf_rain_idw(wmf_ctx *ctx, const float *xy_basin, const float *coord, const float *rain, float pp,
                 int32_t nceldas, int32_t ncoord, int32_t nreg, float umbral, int32_t *table, int32_t capacity,
                 float *mean_rain, int32_t *pos_ids, int32_t *n_records, const int32_t *mask_vector, int32_t nhills);
int wmf_rain_mit(wmf_ctx *ctx, const float *xy_basin, const float *coord, const float *rain, const int32_t *tin,
                 const int32_t *tin_perte, int32_t nceldas, int32_t ncoord, int32_t ntin, int32_t nreg, float umbral,
                 int32_t *table, int32_t capacity, float *mean_rain, int32_t *pos_ids, int32_t *n_records);

/* calc_speed, modelosv2.f90:1278-1293 (one evaluation on the device; mirrors the f2py export) */
int wmf_calc_speed(wmf_ctx *ctx, float sm, float coef, float expo, float elem_long, float dt,
                   int32_t calc_niter, float *speed_inout, float *area_out);

/* time spent inside the last wmf_shia_run between its first and last kernel, CUDA events, ms */
int wmf_last_kernel_ms(wmf_ctx *ctx, float *ms);
/* topology statistics of the last wmf_shia_run: levels (tree depth+1) and waves launched */
int wmf_last_shia_stats(wmf_ctx *ctx, int32_t *n_levels, int32_t *n_waves);

#ifdef __cplusplus
}
#endif
#endif /* WMF_B200_H */
