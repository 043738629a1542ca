/*
 * wmf_oracle_models.c -- CPU ORACLE (test infrastructure, NOT product code).
 * Literal sequential restatement of shia_v1 and its sub-models from
 * wmf/modelosv2.f90 (module models).  See wmf_oracle.h (pinned against oracle/_ref).
 *
 * Scope: vertical tanks, hillslope outflow (linear / kinematic), routing,
 * channel kinematic wave, control-point records, per-step means and balance,
 * SED sediment sub-model, landslide sub-model.  Not restated:
 * separate_fluxes, separate_rain, floods, per-step file dumps.
 *
 * Fenced quirks (undefined in the reference; oracle AND kernels do this):
 *  - vspeed is always computed (the reference leaves it uninitialised when
 *    HspeedIn is supplied, modelosv2.f90:285-299).
 *  - speed_type(4) is read out of bounds at :293; hspeed(4,:) starts at 0
 *    (the channel always uses the kinematic solver).
 *  - hflux(4)/section_area are stale for a type-1 control cell (:750,:764);
 *    the oracle records 0 there.
 *  - at the outlet drenaid = N+1 (:468): the unguarded writes
 *    StoOut(4,drenaid) (:617), Vs(i,drena_id) (:1374,:1401,:1418) and
 *    Vsc(i,drena_id) (:1601) are dropped.
 *  - slide_hill2gullie walks past the outlet (:1695); the walk stops there.
 *  - x**2.0 in slide_allocate is evaluated as x*x.
 */
#include "wmf_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define A4(a, k, i) a[4 * (size_t)(i) + (k)]
#define A5(a, k, i) a[5 * (size_t)(i) + (k)]
#define A3(a, k, i) a[3 * (size_t)(i) + (k)]

/* modelosv2.f90:1278-1293 */
float wo_calc_speed(float sm, float coef, float expo, float elem_long, float *speed, float dt,
                    int32_t calc_niter)
{
    float area = 0.0f;
    for (int i = 0; i < calc_niter; ++i) {
        area = sm / (elem_long + *speed * dt);
        float new_speed = coef * powf(area, expo);
        *speed = (2.0f * new_speed + *speed) / 3.0f;
    }
    return area;
}

typedef struct {
    const wo_shia_in *in;
    wo_shia_out *out;
    float qlin_sed;
    float DEP[3];
    float EROt[3], DEPt[3];
    double EROt64[3], DEPt64[3];
} sed_state;

/* modelosv2.f90:1321-1428 ; celda, drena_id 0-based (drena_id == n_cel means "outside") */
static void sed_hillslope(sed_state *S, float alfa, float v2, float S2, float So, int32_t celda,
                          int32_t drena_id, int32_t tipo)
{
    const wo_shia_in *in = S->in;
    wo_shia_out *o = S->out;
    const float dt = in->dt, dxp = in->dxp;
    const int32_t N = in->n_cel;
    float qsSUStot = 0, qsBMtot = 0, qsEROtot = 0, qsSUS[3] = {0, 0, 0}, qsBM[3] = {0, 0, 0},
          qsERO[3] = {0, 0, 0};
    float SUStot = 0, DEPtot = 0, Te[3];
    float *Vs = o->vs, *Vd = o->vd, *Vsc = o->vsc;
    S->DEP[0] = S->DEP[1] = S->DEP[2] = 0;
    float Qskr = alfa * powf(So, 1.664f) * powf(S->qlin_sed, 2.035f) * in->krus[celda] *
                 in->crus[celda] * in->prus[celda] * dxp * dt;
    for (int i = 0; i < 3; ++i) {
        if (S2 / 1000.0f > in->wi[i] * dt)
            Te[i] = in->wi[i] * dt * 1000.0f / S2;
        else
            Te[i] = 1.0f;
        A3(Vd, i, celda) = A3(Vd, i, celda) + Te[i] * A3(Vs, i, celda);
        A3(Vs, i, celda) = A3(Vs, i, celda) - Te[i] * A3(Vs, i, celda);
        S->DEP[i] = Te[i] * A3(Vs, i, celda);
        SUStot = SUStot + A3(Vs, i, celda);
        DEPtot = DEPtot + A3(Vd, i, celda);
    }
    for (int i = 0; i < 3; ++i) {
        if (A3(Vs, i, celda) > 0.0f) {
            if (Qskr < SUStot) {
                float cap = Qskr * A3(Vs, i, celda) / SUStot;
                float Adv = A3(Vs, i, celda) * v2 * dt / (dxp + v2 * dt);
                qsSUS[i] = fmaxf(cap, Adv);
            } else {
                qsSUS[i] = A3(Vs, i, celda);
            }
        }
        qsSUStot = qsSUStot + qsSUS[i];
        A3(Vs, i, celda) = A3(Vs, i, celda) - qsSUS[i];
        if (tipo == 1) {
            if (drena_id < N) A3(Vs, i, drena_id) = A3(Vs, i, drena_id) + qsSUS[i];
        } else {
            A3(Vsc, i, celda) = A3(Vsc, i, celda) + qsSUS[i];
        }
    }
    float totXSScap = fmaxf(0.0f, Qskr - qsSUStot);
    if (totXSScap > 0.0f && DEPtot > 0.0f) {
        for (int i = 0; i < 3; ++i) {
            if (A3(Vd, i, celda) > 0.0f) {
                if (totXSScap < DEPtot)
                    qsBM[i] = totXSScap * A3(Vd, i, celda) / DEPtot;
                else
                    qsBM[i] = A3(Vd, i, celda);
            }
            qsBMtot = qsBMtot + qsBM[i];
            S->DEP[i] = S->DEP[i] - qsBM[i];
            if (S->DEP[i] < 0) S->DEP[i] = 0;
            A3(Vd, i, celda) = A3(Vd, i, celda) - qsBM[i];
            if (tipo == 1) {
                if (drena_id < N) A3(Vs, i, drena_id) = A3(Vs, i, drena_id) + qsBM[i];
            } else {
                A3(Vsc, i, celda) = A3(Vsc, i, celda) + qsBM[i];
            }
        }
    }
    for (int i = 0; i < 3; ++i) {
        S->DEPt[i] = S->DEPt[i] + S->DEP[i];
        S->DEPt64[i] += (double)S->DEP[i];
    }
    float REScap = fmaxf(0.0f, totXSScap - qsBMtot);
    if (REScap > 0.0f) {
        for (int i = 0; i < 3; ++i) {
            qsERO[i] = A3(in->parliac, i, celda) * REScap / 100.0f;
            qsEROtot = qsEROtot + qsERO[i];
            if (tipo == 1) {
                if (drena_id < N) A3(Vs, i, drena_id) = A3(Vs, i, drena_id) + qsERO[i];
            } else {
                A3(Vsc, i, celda) = A3(Vsc, i, celda) + qsERO[i];
            }
        }
    }
    for (int i = 0; i < 3; ++i) {
        S->EROt[i] = S->EROt[i] + qsERO[i];
        S->EROt64[i] += (double)qsERO[i];
    }
    o->volero[celda] = o->volero[celda] + ((qsERO[0] + qsERO[1]) + qsERO[2]);
    o->voldepo[celda] = o->voldepo[celda] + ((S->DEP[0] + S->DEP[1]) + S->DEP[2]);
    (void)qsEROtot;
}

/* modelosv2.f90:1537-1608 */
static void sed_channel(sed_state *S, float S5, float v5, float Q5, float So, float area_sec,
                        int32_t celda, int32_t drena_id, float VolSal[3])
{
    const wo_shia_in *in = S->in;
    wo_shia_out *o = S->out;
    const float dt = in->dt, dxp = in->dxp;
    const int32_t N = in->n_cel;
    const float grav = 9.8f, Gsed = 2.65f;
    float Cw[3], Te[3], qsSUS[3] = {0, 0, 0}, qsBM[3] = {0, 0, 0};
    float L, Rh;
    float *Vsc = o->vsc, *Vdc = o->vdc;
    S->DEP[0] = S->DEP[1] = S->DEP[2] = 0;
    if (S5 > 0.0f) {
        L = area_sec * 1000.0f / S5;
        Rh = L * S5 / (1000.0f * L + 2.0f * S5);
    } else {
        Rh = 0.0f;
    }
    for (int i = 0; i < 3; ++i) {
        float d = in->diametro[i];
        Cw[i] = 0.05f * (Gsed / (Gsed - 1.0f)) * (v5 * So) / (sqrtf((Gsed - 1.0f) * grav * d / 1000.0f)) *
                sqrtf((Rh * So) / ((Gsed - 1.0f) * (d / 1000.0f)));
        if (S5 / 1000.0f > in->wi[i] * dt)
            Te[i] = in->wi[i] * dt * 1000.0f / S5;
        else
            Te[i] = 1.0f;
        A3(Vdc, i, celda) = A3(Vdc, i, celda) + Te[i] * A3(Vsc, i, celda);
        A3(Vsc, i, celda) = A3(Vsc, i, celda) - Te[i] * A3(Vsc, i, celda);
        S->DEP[i] = S->DEP[i] + Te[i] * A3(Vsc, i, celda);
    }
    for (int i = 0; i < 3; ++i) {
        float supply = A3(Vsc, i, celda) + A3(Vdc, i, celda);
        qsSUS[i] = 0.0f;
        if (supply > 0.0f) {
            float VolEH = Q5 * Cw[i] * dt / 2.65f;
            float AdvF = fminf(1.0f, v5 * dt / (dxp + v5 * dt));
            qsSUS[i] = A3(Vsc, i, celda) * AdvF;
            float XSScap = fmaxf(0.0f, VolEH - qsSUS[i]);
            qsBM[i] = A3(Vdc, i, celda) * AdvF;
            qsBM[i] = fminf(XSScap, qsBM[i]);
            S->DEP[i] = S->DEP[i] - qsBM[i];
            if (S->DEP[i] < 0.0f) S->DEP[i] = 0.0f;
            A3(Vsc, i, celda) = A3(Vsc, i, celda) - qsSUS[i];
            A3(Vdc, i, celda) = A3(Vdc, i, celda) - qsBM[i];
            if (drena_id < N) A3(Vsc, i, drena_id) = A3(Vsc, i, drena_id) + qsSUS[i] + qsBM[i];
            VolSal[i] = (qsSUS[i] + qsBM[i]) / dt;
        }
    }
    o->voldepo[celda] = o->voldepo[celda] + ((S->DEP[0] + S->DEP[1]) + S->DEP[2]);
}

/* modelosv2.f90:1613-1648 */
static void slide_allocate(const wo_shia_in *in, wo_shia_out *o)
{
    const int32_t N = in->n_cel;
    const float gw = in->sl_gammaw;
    for (int32_t i = 0; i < N; ++i) {
        float gs = in->sl_gammas[i], co = in->sl_cohesion[i], fa = in->sl_frictionangle[i],
              rs = in->sl_radslope[i], zs = in->sl_zs[i];
        float c2 = cosf(rs) * cosf(rs);
        o->sl_zmin[i] = co / ((gw * c2 * tanf(fa)) + (gs * c2 * (tanf(rs) - tanf(fa))));
        o->sl_zcrit[i] = (gs / gw) * zs * (1.0f - (tanf(rs) / tanf(fa))) + (co / (gw * c2 * tanf(fa)));
        o->sl_zmax[i] = co / ((gs * c2) * (tanf(rs) - tanf(fa)));
        o->sl_bo[i] = atanf(-tanf(fa * (gw - gs) / gs));
        float rv = 2.0f;
        if (in->unit_type[i] == 3) rv = 0.0f;
        if (rs < o->sl_bo[i]) rv = 0.0f;
        if (zs < o->sl_zmin[i] && rs < o->sl_bo[i]) rv = 0.0f;
        if (zs > o->sl_zmax[i] && rs > o->sl_bo[i]) rv = 1.0f;
        o->sl_riskvector[i] = rv;
        o->sl_slideocurrence[i] = 0.0f;
        o->sl_slideacumulate[i] = 0;
    }
    for (int32_t t = 0; t < in->n_reg; ++t) o->sl_slidencelltime[t] = 0;
}

/* modelosv2.f90:1683-1707 ; cell 0-based, t 0-based */
static void slide_hill2gullie(const wo_shia_in *in, wo_shia_out *o, int32_t cell, int32_t t)
{
    const int32_t N = in->n_cel;
    if (in->unit_type[cell] <= 2) {
        int32_t local = cell;
        for (;;) {
            if (in->drena[local] == 0) break; /* fenced: outlet */
            int32_t direction = N - in->drena[local];
            if (in->unit_type[direction] > in->unit_type[local]) break;
            o->sl_slideocurrence[direction] = 3.0f;
            o->sl_slideacumulate[direction] += 1;
            o->sl_slidencelltime[t] += 1;
            local = direction;
        }
    }
}

/* modelosv2.f90:1649-1682 */
static void slide_ocurrence(const wo_shia_in *in, wo_shia_out *o, int32_t t, int32_t cell,
                            float StorageT3, float MaxStoT3)
{
    if (o->sl_riskvector[cell] == 2.0f && o->sl_slideocurrence[cell] == 0.0f) {
        float Zw = (StorageT3 * in->sl_zs[cell]) / MaxStoT3;
        if (Zw >= o->sl_zcrit[cell]) {
            o->sl_slideocurrence[cell] = 1.0f;
            o->sl_slideacumulate[cell] = 1;
            o->sl_slidencelltime[t] += 1;
            if (in->sl_gullienogullie == 1) slide_hill2gullie(in, o, cell, t);
        } else {
            float rs = in->sl_radslope[cell];
            float Num = in->sl_cohesion[cell] + (in->sl_gammas[cell] * in->sl_zs[cell] - Zw * in->sl_gammaw) *
                                                    (cosf(rs) * cosf(rs)) * tanf(in->sl_frictionangle[cell]);
            float Den = in->sl_gammas[cell] * in->sl_zs[cell] * sinf(rs) * cosf(rs);
            if (Num / Den <= in->sl_fs) {
                o->sl_slideocurrence[cell] = 1.0f;
                o->sl_slideacumulate[cell] = 1;
                o->sl_slidencelltime[t] += 1;
                if (in->sl_gullienogullie == 1) slide_hill2gullie(in, o, cell, t);
            }
        }
    }
}

static float sum_f(const float *a, size_t n)
{
    float s = 0.0f;
    for (size_t i = 0; i < n; ++i) s += a[i];
    return s;
}
static double sum_d(const float *a, size_t n)
{
    double s = 0.0;
    for (size_t i = 0; i < n; ++i) s += (double)a[i];
    return s;
}

/* gfortran 5.3's inline MAX(0.0, x) as compiled in the reference object: 0 when x is NaN (an empty capillary tank makes
   :486 a 0/0; checked against oracle/_ref) */
#define WO_FMAX0(x) (((x) > 0.0f) ? (x) : 0.0f)
/* MAX(x, 0.0) with the operands the other way round (rain_idw / rain_mit); NaN behaviour checked against oracle/_ref */
#define WO_MAXX0(x) (((x) > 0.0f) ? (x) : 0.0f)

/* HydroFlash: flood_params (modelosv2.f90:1730-1749) + flood_debris_flow2 (:1786-1822) for one channel cell at one step */
static void flood_cell(const wo_shia_in *in, wo_shia_out *out, int32_t c, float Q, float speed, float *dif_buf)
{
    const int32_t S = in->flood_sec_tam, N = in->n_cel;
    const float d50 = in->flood_d50[c], dw = in->flood_dw, dsed = in->flood_dsed, cmax = in->flood_cmax;
    out->flood_q[c] = Q;
    out->flood_speed[c] = speed;
    float h = Q / (speed * in->flood_w[c]);
    if (h > in->flood_hmax) h = in->flood_hmax;
    out->flood_h[c] = h;
    const float ufr = speed / (5.75f * logf(h / d50) + 6.25f);
    out->flood_ufr[c] = ufr;
    const float cr = cmax * powf(0.06f * h, 0.2f / ufr);
    out->flood_cr[c] = cr;
    const float rdf = (1.0f / d50) * powf((9.81f / 0.0128f) * (cr + (1.0f - cr) * (dw / dsed)), 0.5f) *
                      (powf(cmax / cr, 1.0f / 3.0f) - 1.0f);
    const float Qsed = Q * (1.0f + (cr / (1.0f - cr)));
    out->flood_qsed[c] = Qsed;
    /* flood_debris_flow2 */
    float area = 0.0f, Qp = 0.0f, dif = 0.0f;
    int32_t i = 1;
    const float *sec = in->flood_sections + (size_t)S * c;
    const float fondo = sec[S / 2];
    for (int32_t k = 0; k < S; ++k) dif_buf[k] = 0.0f; /* `diferencias` is uninitialised when the loop never runs: zeros */
    while (Qp < Qsed && (float)i * in->flood_step < 50.0f) {
        dif = (float)i * in->flood_step;
        const float linea = fondo + dif;
        float sum = 0.0f;
        for (int32_t k = 0; k < S; ++k) {
            dif_buf[k] = linea - sec[k];
            if (dif_buf[k] > 0.0f) sum += dif_buf[k];
        }
        area = sum * in->dxp;
        Qp = ((2.0f / 5.0f) * rdf * powf((float)i * in->flood_step, 3.0f / 2.0f) * in->flood_slope[c] * (1.0f / 2.0f)) * area;
        i = i + 1;
    }
    if (out->flood_profundidad) memcpy(out->flood_profundidad + (size_t)S * c, dif_buf, sizeof(float) * (size_t)S);
    const float *cells = in->flood_sec_cells + (size_t)S * c;
    for (int32_t k = 0; k < S; ++k)
        if (dif_buf[k] > 0.0f) {
            const int32_t pos = (int32_t)cells[k];
            if (pos >= 1 && pos <= N) out->flood_flood[pos - 1] = 1;
        }
    out->flood_flood[c] = 2;
    /* flood_debris_flow2 never assigns its first output `areas` (it computes the local `area`): flood_area keeps its value */
    (void)area;
    if (out->flood_scalars) { out->flood_scalars[0] = rdf; out->flood_scalars[2] = dif; }
}

/* modelosv2.f90:191-869 */
int wo_shia_v1(const wo_shia_in *in, wo_shia_out *out)
{
    const int32_t N = in->n_cel, T = in->n_reg, NC = in->n_cont, NH = in->n_conth;
    const float dt = in->dt;
    if (N <= 0 || T <= 0 || NC < 1 || in->calc_niter < 1) return 1;
    float *Sto = out->stoout;
    float *vspeed = (float *)malloc(sizeof(float) * 4 * (size_t)N);
    float *hspeed = out->hspeed_out;
    float *H = (float *)malloc(sizeof(float) * 3 * (size_t)N);
    float *m3_mm = (float *)malloc(sizeof(float) * (size_t)N);
    float *rain = (float *)malloc(sizeof(float) * (size_t)N);
    float *vfl_map = in->save_vfluxes ? (float *)malloc(sizeof(float) * 4 * (size_t)N) : NULL;
    /* Fluxes(3,N_cel+1) :340-344 (one spare column: the outlet's out-of-range receiver is never written here) */
    float *Fluxes = (in->separate_fluxes == 1) ? (float *)calloc(3 * ((size_t)N + 1), sizeof(float)) : NULL;
    if (Fluxes && out->qseparated) memset(out->qseparated, 0, sizeof(float) * 3 * (size_t)NC * T);
    /* separate_rain :346-359: shadow storages (one spare column for the outlet's out-of-range receiver), Conv / Stra = 0 */
    const int sepr = in->separate_rain == 1;
    float *Sc = sepr ? (float *)calloc(5 * ((size_t)N + 1), sizeof(float)) : NULL;
    float *Ss = sepr ? (float *)calloc(5 * ((size_t)N + 1), sizeof(float)) : NULL;
    int32_t *Conv = sepr ? (int32_t *)calloc((size_t)N, sizeof(int32_t)) : NULL;
    int32_t *Stra = sepr ? (int32_t *)calloc((size_t)N, sizeof(int32_t)) : NULL;
    float hflux_c[4] = {0, 0, 0, 0}, hflux_s[4] = {0, 0, 0, 0};
    if (sepr && out->qsep_byrain) memset(out->qsep_byrain, 0, sizeof(float) * 2 * (size_t)NC * T);
    float *flood_dif = NULL;
    if (in->sim_floods == 1) { /* flood_allocate :1711-1729: flood_eval = flood_flood = 0; the other maps start at zero here */
        flood_dif = (float *)calloc((size_t)(in->flood_sec_tam > 0 ? in->flood_sec_tam : 1), sizeof(float));
        for (int32_t i = 0; i < N; ++i) {
            out->flood_flood[i] = 0;
            out->flood_h[i] = out->flood_ufr[i] = out->flood_cr[i] = out->flood_q[i] = out->flood_qsed[i] = out->flood_speed[i] = 0.0f;
        }
        if (out->flood_profundidad) memset(out->flood_profundidad, 0, sizeof(float) * (size_t)in->flood_sec_tam * N);
        if (out->flood_scalars) out->flood_scalars[0] = out->flood_scalars[1] = out->flood_scalars[2] = 0.0f;
    }
    sed_state SS;
    memset(&SS, 0, sizeof SS);
    SS.in = in;
    SS.out = out;

    /* setup :257-304 */
    for (int32_t t = 0; t < T; ++t) out->mean_rain[t] = 0.0f;
    for (int32_t i = 0; i < N; ++i) out->acum_rain[i] = 0.0f;
    for (int32_t i = 0; i < N; ++i) m3_mm[i] = in->elem_area[i] / 1000.0f;
    for (size_t k = 0; k < (size_t)NC * T; ++k) out->q[k] = 0.0f;
    if (out->speed) memset(out->speed, 0, sizeof(float) * (size_t)NC * T);
    if (out->areacontrol) memset(out->areacontrol, 0, sizeof(float) * (size_t)NC * T);
    if (out->hum) memset(out->hum, 0, sizeof(float) * (size_t)NH * T);
    if (out->st1) memset(out->st1, 0, sizeof(float) * (size_t)NH * T);
    if (out->st3) memset(out->st3, 0, sizeof(float) * (size_t)NH * T);
    if (in->stoin && in->stoin[0] > 0.0f)
        memcpy(Sto, in->stoin, sizeof(float) * 5 * (size_t)N);
    else
        for (size_t k = 0; k < 5 * (size_t)N; ++k) Sto[k] = in->storage[k] + in->storage_constant;
    float entradas = 0.0f, salidas = 0.0f;
    double entradas64 = 0.0, salidas64 = 0.0;
    for (int32_t t = 0; t < T; ++t) out->balance[t] = 0.0f;
    for (int32_t i = 0; i < N; ++i)
        for (int k = 0; k < 4; ++k) A4(vspeed, k, i) = A4(in->v_coef, k, i) * in->calib[k] * dt;
    if (in->hspeedin && in->hspeedin[0] > 0.0f) {
        memcpy(hspeed, in->hspeedin, sizeof(float) * 4 * (size_t)N);
    } else {
        for (int32_t i = 0; i < N; ++i) {
            for (int k = 0; k < 3; ++k)
                A4(hspeed, k, i) = (in->speed_type[k] == 1) ? A4(in->h_coef, k, i) * in->calib[k + 4] : 0.0f;
            A4(hspeed, 3, i) = 0.0f;
        }
    }
    for (int32_t i = 0; i < N; ++i) {
        A3(H, 0, i) = in->max_capilar[i] * in->calib[8];
        A3(H, 1, i) = in->max_gravita[i] * in->calib[9];
        A3(H, 2, i) = in->max_aquifer[i] * in->calib[10];
    }
    /* optional setup :310-426 */
    if (in->retorno_gr == 1.0f && out->retorned)
        for (int32_t i = 0; i < N; ++i) out->retorned[i] = 0.0f;
    if (in->retorno_gr == 1.0f && in->show_mean_retorno == 1 && out->mean_retorno)
        for (int32_t t = 0; t < T; ++t) out->mean_retorno[t] = 0.0f;
    if (in->sim_sediments == 1) {
        /* sed_allocate :1298-1319 */
        for (size_t k = 0; k < 3 * (size_t)N; ++k) {
            out->vd[k] = in->vd_init ? in->vd_init[k] : 0.0f;
            out->vs[k] = 0.0f;
            out->vsc[k] = 0.0f;
            out->vdc[k] = 0.0f;
        }
        for (int32_t i = 0; i < N; ++i) out->volero[i] = out->voldepo[i] = 0.0f;
        for (size_t k = 0; k < (size_t)NC * 3 * T; ++k) out->qsed[k] = 0.0f;
    }
    if (in->sim_slides == 1) slide_allocate(in, out);
    if (in->save_rc == 1 && out->rc_coef)
        for (size_t k = 0; k < 2 * (size_t)N; ++k) out->rc_coef[k] = 0.0f;

    /* time loop :433 */
    for (int32_t t = 0; t < T; ++t) {
        int32_t control_cont = 2, controlh_cont = 1;
        float StoAtras = sum_f(Sto, 5 * (size_t)N);
        double StoAtras64 = sum_d(Sto, 5 * (size_t)N);
        if (in->pos_evento[t] == 1) {
            for (int32_t i = 0; i < N; ++i) rain[i] = 0.0f;
        } else {
            const int32_t *rec = in->rain + (size_t)(in->pos_evento[t] - 1) * N;
            for (int32_t i = 0; i < N; ++i) rain[i] = (float)rec[i] / 1000.0f;
            /* :452-456 -- the rain types are read on wet steps only: a dry step keeps the previous step's Conv / Stra;
               a record outside the file fails the read and leaves the vector as it was */
            if (sepr) {
                const int32_t rc = in->pos_conv[t], rs = in->pos_stra[t];
                if (rc >= 1 && rc <= in->n_conv_records) memcpy(Conv, in->conv + (size_t)(rc - 1) * N, sizeof(int32_t) * (size_t)N);
                if (rs >= 1 && rs <= in->n_stra_records) memcpy(Stra, in->stra + (size_t)(rs - 1) * N, sizeof(int32_t) * (size_t)N);
            }
        }
        float rain_sum = 0.0f;
        for (int32_t i = 0; i < N; ++i) out->acum_rain[i] = out->acum_rain[i] + rain[i];
        float hflux[4] = {0, 0, 0, 0}, vflux[4];
        float section_area = 0.0f;
        float Vsal_sed[3] = {0, 0, 0};

        for (int32_t c = 0; c < N; ++c) {
            const int32_t drenaid = N - in->drena[c]; /* 0-based; == N at the outlet */
            const int32_t ut = in->unit_type[c];
            entradas = entradas + rain[c];
            entradas64 += (double)rain[c];
            rain_sum = rain_sum + rain[c];
            float Co = 0.0f, St = 0.0f;
            if (sepr) { Co = (float)(Conv[c] / 1000); St = (float)(Stra[c] / 1000); } /* integer division, :471-474 */
            /* vertical :478-512 */
            vflux[0] = fmaxf(0.0f, rain[c] - A3(H, 0, c) + A5(Sto, 0, c));
            A5(Sto, 0, c) = A5(Sto, 0, c) + rain[c] - vflux[0];
            float Evp_loss = fminf(in->evpserie[t] * A4(vspeed, 0, c) *
                                       powf(A5(Sto, 0, c) / A3(H, 0, c), 0.6f),
                                   A5(Sto, 0, c));
            if (sepr) { /* :485-488, against the capillary tank BEFORE the evaporation is taken out; max(0., NaN) = 0 as the reference's compiled MAX */
                const float ec = A5(Sc, 0, c) - Evp_loss * A5(Sc, 0, c) / A5(Sto, 0, c);
                const float es = A5(Ss, 0, c) - Evp_loss * A5(Ss, 0, c) / A5(Sto, 0, c);
                A5(Sc, 0, c) = WO_FMAX0(ec);
                A5(Ss, 0, c) = WO_FMAX0(es);
            }
            A5(Sto, 0, c) = A5(Sto, 0, c) - Evp_loss;
            for (int i = 0; i < 3; ++i) {
                vflux[i + 1] = fminf(vflux[i], A4(vspeed, i + 1, c));
                A5(Sto, i + 1, c) = A5(Sto, i + 1, c) + vflux[i] - vflux[i + 1];
            }
            if (in->retorno_aq > 0.0f) {
                float Ret_aq = fmaxf(0.0f, A5(Sto, 3, c) - A3(H, 2, c));
                A5(Sto, 2, c) = A5(Sto, 2, c) + Ret_aq;
                A5(Sto, 3, c) = A5(Sto, 3, c) - Ret_aq;
            }
            float Ret_sep = 0.0f;
            if (in->retorno_gr > 0.0f) {
                float Ret = fmaxf(0.0f, A5(Sto, 2, c) - A3(H, 1, c));
                Ret_sep = Ret;
                A5(Sto, 1, c) = A5(Sto, 1, c) + Ret;
                A5(Sto, 2, c) = A5(Sto, 2, c) - Ret;
                if (out->retorned) out->retorned[c] = out->retorned[c] + Ret;
                vflux[1] = vflux[1] + Ret;
                vflux[2] = vflux[2] - Ret;
            }
            if (sepr) { /* :530-547; vflux(2:3) already carry the return flow (:510-511) and Ret is added again here */
                A5(Sc, 0, c) = A5(Sc, 0, c) + (rain[c] - vflux[0]) * Co;
                A5(Ss, 0, c) = A5(Ss, 0, c) + (rain[c] - vflux[0]) * St;
                for (int i = 0; i < 3; ++i) {
                    A5(Sc, i + 1, c) = A5(Sc, i + 1, c) + (vflux[i] - vflux[i + 1]) * Co;
                    A5(Ss, i + 1, c) = A5(Ss, i + 1, c) + (vflux[i] - vflux[i + 1]) * St;
                }
                if (in->retorno_gr > 0.0f) {
                    A5(Sc, 1, c) = A5(Sc, 1, c) + Ret_sep * Co;
                    A5(Sc, 2, c) = A5(Sc, 2, c) - Ret_sep * Co;
                    A5(Ss, 1, c) = A5(Ss, 1, c) + Ret_sep * St;
                    A5(Ss, 2, c) = A5(Ss, 2, c) - Ret_sep * St;
                }
            }
            if (vfl_map)
                for (int i = 0; i < 4; ++i) A4(vfl_map, i, c) = vflux[i];
            if (in->save_rc == 1 && out->rc_coef) {
                out->rc_coef[2 * (size_t)c] = out->rc_coef[2 * (size_t)c] + vflux[1] - vflux[2];
                out->rc_coef[2 * (size_t)c + 1] = out->rc_coef[2 * (size_t)c + 1] + rain[c];
            }
            salidas = salidas + vflux[3] + Evp_loss;
            salidas64 += (double)vflux[3] + (double)Evp_loss;
            /* hillslope outflow :553-595 */
            for (int i = 0; i < 3; ++i) {
                if (in->speed_type[i] == 1) {
                    hflux[i] = (1.0f - in->hill_long[c] / (A4(hspeed, i, c) * dt + in->hill_long[c])) *
                               A5(Sto, i + 1, c);
                } else if (in->speed_type[i] == 2) {
                    section_area = wo_calc_speed(A5(Sto, i + 1, c) * m3_mm[c],
                                                 A4(in->h_coef, i, c) * in->calib[i + 4], A4(in->h_exp, i, c),
                                                 in->hill_long[c], &A4(hspeed, i, c), dt, in->calc_niter);
                    hflux[i] = fminf(section_area * A4(hspeed, i, c) * dt / m3_mm[c], A5(Sto, i + 1, c));
                    if (in->sim_sediments == 1 && i == 0) {
                        SS.qlin_sed = hflux[i] * m3_mm[c] / (dt * in->dxp);
                        sed_hillslope(&SS, in->sed_factor, A4(hspeed, 0, c), A5(Sto, 1, c), in->hill_slope[c],
                                      c, drenaid, ut);
                    }
                }
                if (sepr) { /* :580-592 */
                    if (A5(Sto, i + 1, c) > 0.0f) {
                        hflux_c[i] = hflux[i] * A5(Sc, i + 1, c) / A5(Sto, i + 1, c);
                        hflux_s[i] = hflux[i] * A5(Ss, i + 1, c) / A5(Sto, i + 1, c);
                    } else {
                        hflux_c[i] = 0.0f;
                        hflux_s[i] = 0.0f;
                    }
                    A5(Sc, i + 1, c) = A5(Sc, i + 1, c) - hflux_c[i];
                    A5(Ss, i + 1, c) = A5(Ss, i + 1, c) - hflux_s[i];
                }
                A5(Sto, i + 1, c) = A5(Sto, i + 1, c) - hflux[i];
            }
            /* routing :599-719 */
            if (ut == 1) {
                if (in->drena[c] != 0) {
                    for (int i = 0; i < 3; ++i) A5(Sto, i + 1, drenaid) = A5(Sto, i + 1, drenaid) + hflux[i];
                    if (sepr)
                        for (int i = 0; i < 3; ++i) { /* :605-608 */
                            A5(Sc, i + 1, drenaid) = A5(Sc, i + 1, drenaid) + hflux_c[i];
                            A5(Ss, i + 1, drenaid) = A5(Ss, i + 1, drenaid) + hflux_s[i];
                        }
                } else {
                    float s3 = (hflux[0] + hflux[1]) + hflux[2];
                    out->q[(size_t)t * NC + 0] = out->q[(size_t)t * NC + 0] + s3 * m3_mm[c];
                    salidas = salidas + s3;
                    salidas64 += (double)s3;
                }
            } else if (ut > 1) {
                if (drenaid < N) A5(Sto, 3, drenaid) = A5(Sto, 3, drenaid) + hflux[2] * (float)(3 - ut);
                A5(Sto, 4, c) = A5(Sto, 4, c) + (hflux[0] + hflux[1]) + hflux[2] * (float)(ut - 2);
                if (sepr) { /* :623-628 (the spare column takes the outlet's out-of-range write) */
                    A5(Sc, 3, drenaid) = A5(Sc, 3, drenaid) + hflux_c[2] * (float)(3 - ut);
                    A5(Ss, 3, drenaid) = A5(Ss, 3, drenaid) + hflux_s[2] * (float)(3 - ut);
                    A5(Sc, 4, c) = A5(Sc, 4, c) + (hflux_c[0] + hflux_c[1]) + hflux_c[2] * (float)(ut - 2);
                    A5(Ss, 4, c) = A5(Ss, 4, c) + (hflux_s[0] + hflux_s[1]) + hflux_s[2] * (float)(ut - 2);
                }
                section_area = wo_calc_speed(A5(Sto, 4, c) * m3_mm[c], A4(in->h_coef, 3, c) * in->calib[7],
                                             A4(in->h_exp, 3, c), in->stream_long[c], &A4(hspeed, 3, c), dt,
                                             in->calc_niter);
                hflux[3] = fminf(section_area * A4(hspeed, 3, c) * dt / m3_mm[c], A5(Sto, 4, c));
                if (in->sim_sediments == 1) {
                    Vsal_sed[0] = Vsal_sed[1] = Vsal_sed[2] = 0.0f;
                    sed_channel(&SS, A5(Sto, 4, c), A4(hspeed, 3, c), hflux[3] * m3_mm[c] / dt,
                                in->stream_slope[c], section_area, c, drenaid, Vsal_sed);
                }
                if (sepr) { /* :645-655 */
                    if (A5(Sto, 4, c) != 0.0f) {
                        hflux_c[3] = hflux[3] * A5(Sc, 4, c) / A5(Sto, 4, c);
                        hflux_s[3] = hflux[3] * A5(Ss, 4, c) / A5(Sto, 4, c);
                    } else {
                        hflux_c[3] = 0.0f;
                        hflux_s[3] = 0.0f;
                    }
                    A5(Sc, 4, c) = A5(Sc, 4, c) - hflux_c[3];
                    A5(Ss, 4, c) = A5(Ss, 4, c) - hflux_s[3];
                }
                /* separate_fluxes :658-665 -- note the ratio uses StoOut(5) BEFORE the outflow is taken out */
                if (Fluxes) {
                    for (int i = 0; i < 3; ++i) Fluxes[3 * (size_t)c + i] = Fluxes[3 * (size_t)c + i] + hflux[i];
                    if (A5(Sto, 4, c) > 0.0f) {
                        const float ratio = hflux[3] / A5(Sto, 4, c);
                        for (int i = 0; i < 3; ++i)
                            Fluxes[3 * (size_t)c + i] = Fluxes[3 * (size_t)c + i] - Fluxes[3 * (size_t)c + i] * ratio;
                    } else {
                        for (int i = 0; i < 3; ++i) Fluxes[3 * (size_t)c + i] = 0.0f;
                    }
                }
                A5(Sto, 4, c) = A5(Sto, 4, c) - hflux[3];
                if (in->drena[c] != 0) {
                    A5(Sto, 4, drenaid) = A5(Sto, 4, drenaid) + hflux[3];
                    if (sepr) { /* :683-686 */
                        A5(Sc, 4, drenaid) = A5(Sc, 4, drenaid) + hflux_c[3];
                        A5(Ss, 4, drenaid) = A5(Ss, 4, drenaid) + hflux_s[3];
                    }
                    /* :674-680 -- and here AFTER; an empty channel zeroes the receiver's shares */
                    if (Fluxes) {
                        if (A5(Sto, 4, c) > 0.0f) {
                            const float ratio = hflux[3] / A5(Sto, 4, c);
                            for (int i = 0; i < 3; ++i)
                                Fluxes[3 * (size_t)drenaid + i] = Fluxes[3 * (size_t)drenaid + i] + Fluxes[3 * (size_t)c + i] * ratio;
                        } else {
                            for (int i = 0; i < 3; ++i) Fluxes[3 * (size_t)drenaid + i] = 0.0f;
                        }
                    }
                } else {
                    if (Fluxes && out->qseparated) { /* :697-704 */
                        for (int i = 0; i < 3; ++i)
                            out->qseparated[((size_t)t * 3 + i) * NC + 0] =
                                (A5(Sto, 4, c) > 0.0f) ? Fluxes[3 * (size_t)c + i] * (hflux[3] / A5(Sto, 4, c)) * m3_mm[c] / dt : 0.0f;
                    }
                    if (sepr && out->qsep_byrain) { /* :713-716 */
                        out->qsep_byrain[((size_t)t * 2 + 0) * NC + 0] = hflux_c[3] * m3_mm[c] / dt;
                        out->qsep_byrain[((size_t)t * 2 + 1) * NC + 0] = hflux_s[3] * m3_mm[c] / dt;
                    }
                    out->q[(size_t)t * NC + 0] = hflux[3] * m3_mm[c] / dt;
                    salidas = salidas + hflux[3];
                    salidas64 += (double)hflux[3];
                    if (in->sim_sediments == 1)
                        for (int i = 0; i < 3; ++i) out->qsed[((size_t)t * 3 + i) * NC + 0] = Vsal_sed[i];
                    if (in->show_speed == 1) {
                        out->speed[(size_t)t * NC + 0] = A4(hspeed, 3, c);
                        out->areacontrol[(size_t)t * NC + 0] = section_area;
                    }
                }
            }
            /* slides :723 */
            if (in->sim_slides == 1) slide_ocurrence(in, out, t, c, A5(Sto, 2, c), A3(H, 1, c));
            /* floods :728-743 -- channel cells of type 3 flowing faster than the threshold */
            if (in->sim_floods == 1 && ut == 3 && A4(hspeed, 3, c) * in->flood_av > in->flood_umbral)
                flood_cell(in, out, c, hflux[3] * m3_mm[c] / dt, A4(hspeed, 3, c) * in->flood_av, flood_dif);
            /* records :749-787 */
            if (in->control[c] != 0) {
                if (control_cont <= NC) {
                    size_t k = (size_t)t * NC + (control_cont - 1);
                    if (ut > 1) {
                        out->q[k] = hflux[3] * m3_mm[c] / dt;
                        if (sepr && out->qsep_byrain) { /* :767-770 */
                            out->qsep_byrain[((size_t)t * 2 + 0) * NC + (control_cont - 1)] = hflux_c[3] * m3_mm[c] / dt;
                            out->qsep_byrain[((size_t)t * 2 + 1) * NC + (control_cont - 1)] = hflux_s[3] * m3_mm[c] / dt;
                        }
                        if (Fluxes && out->qseparated) /* :752-759 */
                            for (int i = 0; i < 3; ++i)
                                out->qseparated[((size_t)t * 3 + i) * NC + (control_cont - 1)] =
                                    (A5(Sto, 4, c) > 0.0f) ? Fluxes[3 * (size_t)c + i] * (hflux[3] / A5(Sto, 4, c)) * m3_mm[c] / dt : 0.0f;
                        if (in->show_speed == 1) {
                            out->speed[k] = A4(hspeed, 3, c);
                            out->areacontrol[k] = section_area;
                        }
                        if (in->sim_sediments == 1)
                            for (int i = 0; i < 3; ++i)
                                out->qsed[((size_t)t * 3 + i) * NC + (control_cont - 1)] = Vsal_sed[i];
                    } else { /* fenced: stale hflux(4) for a hillslope control cell */
                        out->q[k] = 0.0f;
                    }
                }
                control_cont = control_cont + 1;
            }
            if (in->control_h[c] != 0) {
                if (controlh_cont <= NH) {
                    size_t k = (size_t)t * NH + (controlh_cont - 1);
                    out->hum[k] = A5(Sto, 0, c) + A5(Sto, 2, c);
                    out->st1[k] = A5(Sto, 0, c);
                    out->st3[k] = A5(Sto, 2, c);
                }
                controlh_cont = controlh_cont + 1;
            }
        }
        /* per-step epilogue :797-848 */
        out->mean_rain[t] = rain_sum / (float)N;
        if (vfl_map && out->mean_vfluxes)
            for (int k = 0; k < 4; ++k) {
                float s = 0.0f;
                double s64 = 0.0;
                for (int32_t i = 0; i < N; ++i) { s += A4(vfl_map, k, i); s64 += (double)A4(vfl_map, k, i); }
                out->mean_vfluxes[4 * (size_t)t + k] = s / (float)N;
                if (out->mean_vfluxes64) out->mean_vfluxes64[4 * (size_t)t + k] = s64 / (double)N;
            }
        if (vfl_map && out->vfluxes_last && t == T - 1) memcpy(out->vfluxes_last, vfl_map, sizeof(float) * 4 * (size_t)N);
        if (in->show_storage == 1 && out->mean_storage)
            for (int k = 0; k < 5; ++k) {
                float s = 0.0f;
                for (int32_t i = 0; i < N; ++i) s += A5(Sto, k, i);
                out->mean_storage[5 * (size_t)t + k] = s / (float)N;
                if (out->mean_storage64) {
                    double sd = 0.0;
                    for (int32_t i = 0; i < N; ++i) sd += (double)A5(Sto, k, i);
                    out->mean_storage64[5 * (size_t)t + k] = sd / (double)N;
                }
            }
        if (in->show_mean_speed == 1 && out->mean_speed)
            for (int k = 0; k < 4; ++k) {
                float s = 0.0f;
                for (int32_t i = 0; i < N; ++i) s += A4(hspeed, k, i);
                out->mean_speed[4 * (size_t)t + k] = s / (float)N;
                if (out->mean_speed64) {
                    double sd = 0.0;
                    for (int32_t i = 0; i < N; ++i) sd += (double)A4(hspeed, k, i);
                    out->mean_speed64[4 * (size_t)t + k] = sd / (double)N;
                }
            }
        if (in->retorno_gr == 1.0f && in->show_mean_retorno == 1 && out->mean_retorno && out->retorned)
        {
            out->mean_retorno[t] = sum_f(out->retorned, (size_t)N) / (float)N;
            if (out->mean_retorno64) out->mean_retorno64[t] = sum_d(out->retorned, (size_t)N) / (double)N;
        }
        if (out->retorned_last && out->retorned && t == T - 1) memcpy(out->retorned_last, out->retorned, sizeof(float) * (size_t)N);
        if ((in->show_mean_retorno == 1 || in->save_retorno == 1) && out->retorned && in->retorno_gr == 1.0f)
            for (int32_t i = 0; i < N; ++i) out->retorned[i] = 0.0f;
        out->balance[t] = sum_f(Sto, 5 * (size_t)N) - StoAtras - entradas + salidas;
        if (out->balance64) out->balance64[t] = sum_d(Sto, 5 * (size_t)N) - StoAtras64 - entradas64 + salidas64;
        entradas = 0.0f;
        salidas = 0.0f;
        entradas64 = 0.0;
        salidas64 = 0.0;
    }
    if (in->save_rc == 1 && out->rc_coef)
        for (int32_t i = 0; i < N; ++i)
            if (out->rc_coef[2 * (size_t)i] > out->rc_coef[2 * (size_t)i + 1])
                out->rc_coef[2 * (size_t)i] = out->rc_coef[2 * (size_t)i + 1];
    if (in->sim_sediments == 1) {
        if (out->erot) memcpy(out->erot, SS.EROt, sizeof SS.EROt);
        if (out->dept) memcpy(out->dept, SS.DEPt, sizeof SS.DEPt);
        if (out->erot64) memcpy(out->erot64, SS.EROt64, sizeof SS.EROt64);
        if (out->dept64) memcpy(out->dept64, SS.DEPt64, sizeof SS.DEPt64);
    }
    free(vspeed);
    free(H);
    free(m3_mm);
    free(Fluxes);
    free(rain);
    if (sepr) {
        if (out->storage_conv) memcpy(out->storage_conv, Sc, sizeof(float) * 5 * (size_t)N);
        if (out->storage_stra) memcpy(out->storage_stra, Ss, sizeof(float) * 5 * (size_t)N);
    }
    free(flood_dif);
    free(Sc);
    free(Ss);
    free(Conv);
    free(Stra);
    free(vfl_map);
    return 0;
}

/* ---- rain interpolation, modelosv2.f90:1104-1271 (cells model) ---- */
/* what both routines do with a finished field (:1166-1190, :1244-1268) */
/* hills mode (maskVector = hill id of every cell): basin_subbasin_map2subbasin of module models (:1841-1877) is called with
 * sum_mean = celdas_hills = 2, i.e. the hill's value is the SUM of its cells' rainfall, added in ascending cell order; column
 * i of the record is hill n_hills - i + 1 */
static void rain_to_hills(const float *campo, int32_t n, const int32_t *mask, int32_t nhills, int32_t *rec)
{
    for (int32_t i = 1; i <= nhills; ++i) {
        const int32_t posi = nhills - i + 1;
        float s = 0.0f;
        int32_t cnt = 0;
        for (int32_t c = 0; c < n; ++c)
            if (mask[c] == posi) { s += campo[c]; ++cnt; }
        rec[i - 1] = (int32_t)((cnt > 0 ? s : 0.0f) * 1000.0f);
    }
}

static int rain_store(const float *campo, int32_t n, float umbral, int32_t t, int32_t *cont, int32_t *table,
                      int32_t capacity, float *mean_rain, int32_t *pos_ids, double *mean_rain64, const int32_t *mask, int32_t nhills)
{
    float s = 0.0f;
    double s64 = 0.0;
    int32_t n_umbral = 0, n_pos = 0;
    for (int32_t i = 0; i < n; ++i) {
        s += campo[i];
        s64 += (double)campo[i];
        if (campo[i] > umbral) ++n_umbral;
        if (campo[i] > 0.0f) ++n_pos;
    }
    if (s > umbral && n_umbral > 0) {
        mean_rain[t] = s / (float)n_pos;
        if (mean_rain64) mean_rain64[t] = s64 / (double)n_pos;
        if (*cont > capacity) return -1;
        int32_t *rec = table + (size_t)(*cont - 1) * (nhills > 0 ? nhills : n);
        if (nhills > 0) rain_to_hills(campo, n, mask, nhills, rec);
        else
            for (int32_t i = 0; i < n; ++i) rec[i] = (int32_t)(campo[i] * 1000.0f); /* campoInt = campo*1000 truncates */
        pos_ids[t] = *cont;
        *cont = *cont + 1;
    } else {
        mean_rain[t] = 0.0f;
        if (mean_rain64) mean_rain64[t] = 0.0;
        pos_ids[t] = 1;
    }
    return 0;
}

/* modelosv2.f90:1195-1271 */
int32_t wo_rain_idw(const float *xy_basin, const float *coord, const float *rain, float pp, int32_t nceldas,
                    int32_t ncoord, int32_t nreg, float umbral, int32_t *table, int32_t capacity, float *mean_rain,
                    int32_t *pos_ids, double *mean_rain64, const int32_t *mask, int32_t nhills)
{
    if (capacity < 1) return -1;
    float *W = (float *)malloc(sizeof(float) * (size_t)ncoord * nceldas); /* W(ncoord,nceldas) */
    float *campo = (float *)malloc(sizeof(float) * (size_t)nceldas);
    for (int32_t c = 0; c < (nhills > 0 ? nhills : nceldas); ++c) table[c] = 0; /* record 1 = zeros (:1224) */
    for (int32_t i = 0; i < ncoord; ++i) /* :1227-1229 */
        for (int32_t c = 0; c < nceldas; ++c) {
            const float dxs = coord[2 * i] - xy_basin[2 * (size_t)c], dys = coord[2 * i + 1] - xy_basin[2 * (size_t)c + 1];
            W[(size_t)c * ncoord + i] = 1.0f / powf(sqrtf(dxs * dxs + dys * dys), pp);
        }
    int32_t cont = 2, rc = 0;
    for (int32_t t = 0; t < nreg && rc == 0; ++t) {
        const float *r = rain + (size_t)t * ncoord;
        for (int32_t c = 0; c < nceldas; ++c) { /* :1234-1242 */
            float Wr = 0.0f, den = 0.0f;
            for (int32_t i = 0; i < ncoord; ++i) {
                if (r[i] > 0.0f) Wr = Wr + W[(size_t)c * ncoord + i] * r[i];
                if (r[i] >= 0.0f) den = den + W[(size_t)c * ncoord + i];
            }
            const float q = Wr / den;
            const float valor = WO_MAXX0(q);
            campo[c] = (valor == valor - 1.0f) ? 0.0f : valor;
        }
        rc = rain_store(campo, nceldas, umbral, t, &cont, table, capacity, mean_rain, pos_ids, mean_rain64, mask, nhills);
    }
    free(W);
    free(campo);
    return rc ? -1 : cont - 1;
}

/* modelosv2.f90:1104-1194 */
int32_t wo_rain_mit(const float *xy_basin, const float *coord, const float *rain, const int32_t *tin,
                    const int32_t *tin_perte, int32_t nceldas, int32_t ncoord, int32_t ntin, int32_t nreg, float umbral,
                    int32_t *table, int32_t capacity, float *mean_rain, int32_t *pos_ids, double *mean_rain64,
                    const int32_t *mask, int32_t nhills)
{
    if (capacity < 1) return -1;
    float *campo = (float *)malloc(sizeof(float) * (size_t)nceldas);
    for (int32_t c = 0; c < (nhills > 0 ? nhills : nceldas); ++c) table[c] = 0;
    int32_t cont = 2, rc = 0;
    for (int32_t t = 0; t < nreg && rc == 0; ++t) {
        const float *r = rain + (size_t)t * ncoord;
        for (int32_t c = 0; c < nceldas; ++c) {
            const int32_t tri = tin_perte[c] - 1;
            const int32_t ia = tin[3 * tri] - 1, ib = tin[3 * tri + 1] - 1, ic = tin[3 * tri + 2] - 1;
            const float ax = coord[2 * ia], ay = coord[2 * ia + 1], bx = coord[2 * ib], by = coord[2 * ib + 1];
            const float cx = coord[2 * ic], cy = coord[2 * ic + 1];
            const float coef1 = (bx - ax) * (cy - ay) - (cx - ax) * (by - ay); /* :1150 */
            const float Cel_x = xy_basin[2 * (size_t)c], Cel_y = xy_basin[2 * (size_t)c + 1];
            const float az = fmaxf(r[ia], 0.0f), bz = fmaxf(r[ib], 0.0f), cz = fmaxf(r[ic], 0.0f);
            const float det1 = (cx - ax) * (Cel_y - ay), det2 = (by - ay) * (Cel_x - ax);
            const float det3 = (cy - ay) * (Cel_x - ax), det4 = (Cel_y - ay) * (bx - ax);
            const float coef2 = det1 * (bz - az) + det2 * (cz - az) - det3 * (bz - az) - det4 * (cz - az);
            const float q = az - coef2 / coef1;
            campo[c] = WO_MAXX0(q);
        }
        rc = rain_store(campo, nceldas, umbral, t, &cont, table, capacity, mean_rain, pos_ids, mean_rain64, mask, nhills);
    }
    free(campo);
    return rc ? -1 : cont - 1;
}

/* rain_pre_mit, modelosv2.f90:1068-1101: the triangle of every cell; the test is only s > 0 and t > 0 (not s + t < 1) and the
 * last triangle that passes wins; a cell no triangle claims keeps 0 (the Fortran leaves it unassigned) */
void wo_rain_pre_mit(const float *xy_basin, const int32_t *tin, const float *coord, int32_t nceldas, int32_t ntin,
                     int32_t *tin_perte)
{
    for (int32_t c = 0; c < nceldas; ++c) {
        const float px = xy_basin[2 * (size_t)c], py = xy_basin[2 * (size_t)c + 1];
        int32_t own = 0;
        for (int32_t k = 0; k < ntin; ++k) {
            const int32_t i0 = tin[3 * k] - 1, i1 = tin[3 * k + 1] - 1, i2 = tin[3 * k + 2] - 1;
            const float p0x = coord[2 * i0], p1x = coord[2 * i1], p2x = coord[2 * i2];
            const float p0y = coord[2 * i0 + 1], p1y = coord[2 * i1 + 1], p2y = coord[2 * i2 + 1];
            const float Area = fabsf(1.0f / 2.0f * (-p1y * p2x + p0y * (-p1x + p2x) + p0x * (p1y - p2y) + p1x * p2y));
            const float s = 1.0f / (2.0f * Area) * (p0y * p2x - p0x * p2y + (p2y - p0y) * px + (p0x - p2x) * py);
            const float t = 1.0f / (2.0f * Area) * (p0x * p1y - p0y * p1x + (p0y - p1y) * px + (p1x - p0x) * py);
            if (s > 0.0f && t > 0.0f) own = k + 1;
        }
        tin_perte[c] = own;
    }
}
