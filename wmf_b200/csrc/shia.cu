// shia.cu -- Path B: the SHIA/TETIS time loop (module `models`, wmf/modelosv2.f90:191-869) on sm_100a.
//
// Execution model: (level x time) skewed topological wavefront.
//   Cells are bucketed by hop distance to the outlet ("level"; basin_find's ordering makes each level a
//   contiguous index range and the children of a cell a contiguous index range).  Cell (level L, step t)
//   depends only on its children at the same step (level L+1) and on itself at step t-1, so all (L,t) with
//   t + (Lmax - L) == w are independent: wave w is ONE kernel launch over a contiguous cell range.
//   Each thread owns one (cell, member): it gathers its children's outflows of the same step in ascending
//   cell order (== the Fortran's sequential fp32 addition order, no atomics on water), runs the vertical
//   tank cascade and the three hillslope outflows, solves the channel kinematic wave, and publishes its own
//   outflow into a ring double-buffered on the parity of the time block.
// Two kernels implement it:
//   k_shia_tb    the product path: K steps per launch per thread (state in registers for the block, parameters
//                derived once per block, programmatic dependent launch between waves, hillslope and channel cells
//                in different warps); template flags add slides (marks handed downstream through a per-block ring, no
//                walk), state dumps, flux separation and the staging of what the sediments read.
//                AoS state: float4 st4[cell][member] (tanks 1-4) + s5; ring float4 outb4[2][K][cell][member].
//   k_sed_tb     the SED sediments of a time block (sed_hillslope, sed_channel): launched behind every wave of k_shia_tb
//                on a second stream, next to the hydrology of the following wave; same tiles, same time blocks.
//   k_shia_wave  the first generation, one step per launch, SoA state[tank][cell][member]; serves the per-step
//                mean outputs (show_storage, show_mean_speed, mean_retorno) and cross-checks k_shia_tb in the tests.
// Static maps are read in their f2py (K,N) layout (one aligned float4 per cell); rainfall records stay resident in HBM
// (or stream in from pinned host memory next to the sweep, see RainStream) and are indexed per step.
// Reductions (balance, means) are accumulated in fp64 per step slot.
#include <limits.h>
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <chrono>

#include "common.cuh"
#include "wmf_pow.cuh"

constexpr int WAVE_THREADS = 256;
#ifndef WAVE_MIN_BLOCKS
#define WAVE_MIN_BLOCKS 4
#endif
constexpr int RED_SUB = 8;  // sub-slots per (step, quantity) to spread atomic traffic
// reduction quantities per step
enum { RQ_SAL = 0, RQ_STO1 = 1 /* ..5 */, RQ_SPD1 = 6 /* ..9 */, RQ_RET = 10, RQ_VFL1 = 11 /* ..14 */, RQ_COUNT = 15 };

constexpr uint32_t TW_CTRLQ = 1u << 6, TW_CTRLH = 1u << 7;

struct ShiaP {
    int N, T, M, Lmax, n_cont, n_conth;
    float dt, dxp, retorno_gr, retorno_aq;
    int st[3];
    int niter;
    int hs_state;  // HspeedIn supplied: linear tanks read their (constant) speed from the state array
    int show_speed, show_storage, show_mean_speed, want_retorned, want_red;
    // topology
    const uint32_t *tw;   // level<<8 | ctrlH<<7 | ctrlQ<<6 | type<<4 | nchild
    const int32_t *cs;    // first child
    const int32_t *ctrl_cells, *ctrlh_cells;  // sorted cell ids of the control points
    int n_ctrl, n_ctrlh;
    // statics (f2py layouts)
    const float *v_coef, *h_coef, *h_exp;  // (4,N): one float4 per cell
    const float *max_capilar, *max_gravita, *max_aquifer, *hill_long, *hill_slope, *stream_long, *stream_slope,
        *elem_area;
    const float *evp;
    const float *calib;  // (M,11)
    const int32_t *rain;
    const int32_t *pos;
    int pos_stride;       // 0: every member reads pos[t]; T: member m reads pos[m*T + t] (storm-scenario ensembles)
    // state (SoA member-minor)
    float *sto;   // [5][N][M]
    float *hsp;   // [4][N][M]
    float *outb;  // [2][4][N][M]
    float *retorned;  // [N] (M==1)
    // outputs
    float *q, *speed, *area, *hum, *st1o, *st3o, *qsed;
    double *red;  // [T][RQ_COUNT][RED_SUB]
    // sediments
    float sed_factor, wi[3], diam[3];
    const float *krus, *crus, *prus, *parliac;  // parliac (3,N)
    float *sed;      // [12][N][M]: Vs(3) Vd(3) Vsc(3) Vdc(3)
    float *sedout;   // [2][12][N][M]: hillslope->parent sus(3) bm(3) ero(3), channel->parent (3)
    float *volero, *voldepo;  // [N][M]
    double *sedtot;  // [6] EROt(3), DEPt(3): one-step kernel (atomics)
    double *sedacc;  // [6][N][M] the same totals per (cell, member), time-blocked kernel (summed at the end, no atomics)
    // slides
    int sl_gullie;
    float sl_fs, sl_gammaw;
    const float *sl_zs, *sl_gammas, *sl_cohesion, *sl_friction, *sl_radslope;
    const float *sl_zcrit, *sl_risk;
    const float *sl_cosr, *sl_sinr, *sl_tanfa;  // cos/sin of the slope angle, tan of the friction angle
    int32_t *sl_slid;       // [N][M] own failure flag
    int32_t *sl_marktime;   // [N][M] first step at which an upstream failure marked this cell (INT_MAX = never)
    int32_t *sl_acum;       // [N][M]
    int32_t *sl_ncelltime;  // [M][T]
    int2 *sl_ring;          // [2][N][M] time-blocked kernel: (earliest step of this block at which a failure walk passes the cell,
                            // number of such walks | unit type << 30), handed downstream like the water (slide_hill2gullie)
    const int32_t *sl_len;  // [N] cells the marking walk that starts at a cell visits (:1683-1707): static
    const int32_t *drena;   // original drain ids (for the downstream marking walk)
    int drena_stride;
    // time-blocked wavefront (k_shia_tb): cell lists by role, AoS state, per-step outflow ring
    int nb;                 // number of time blocks = ceil(T / K)
    const int32_t *hidx, *cidx;  // hillslope (type 1) / channel (type 2,3) cells, ascending cell id
    float4 *st4;            // [N][M] tanks 1..4
    float *s5;              // [N][M] tank 5
    float4 *outb4;          // [2][K][N][M] what a cell sends downstream at each step of its time block
    int bal_direct;         // red[t][RQ_SAL] holds sum(dS + outputs), balance = that - inputs
    float4 *sedout4;        // [2][K][3][N][M] sediment a cell sends downstream: sus(3) bm(3) ero(3) chan(3)
    float4 *sedstg4;        // [2][K][2][N][M] what the sediment kernel needs from the hydrology of a step: (speed of tank 2, tank 2
                            // before its outflow, qlin, 1 = sed_hillslope runs) and (channel tank before its outflow, channel
                            // speed, discharge, section area)
    const float *sed_aso;   // [N] sed_factor * So**1.664 (:1339), constant over a run
    // state dumps (single runs): snap_slot[t] = row of sto_snap that receives StoOut after step t, or -1
    const int32_t *snap_slot;
    float *sto_snap;        // [n_snap][N][5]
    float *spd_snap;        // [T][N][4]
    const int32_t *vsnap_slot;  // like snap_slot, from guarda_vfluxes
    float *vfl_snap;        // [n_vsnap][N][4] vertical fluxes of the flagged steps (save_vfluxes, :806-813)
    int want_vmean;         // accumulate red[t][RQ_VFL1..4] = sum over the cells of the four vertical fluxes
    float *rc;              // [2][N] save_rc accumulators (:521-524)
    float *ret_snap;        // [T][N] return flow of every step (save_retorno, :820-823)
    // separate_rain (:346-359, :452-456, :471-474, :485-488, :530-547, :580-592, :605-608, :623-628, :645-655, :683-686):
    // convective / stratiform shadows of the five tanks
    float *shc, *shs;             // [5][N][M] Storage_conv / Storage_stra
    float4 *cvout4, *srout4;      // [2][K][N][M] shadow water sent downstream at each step of a time block (like outb4)
    const int32_t *conv, *stra;   // rain-type tables, record r at (r-1)*N
    const int32_t *effc, *effs;   // [T] record in force at step t (the last one read on a wet step), 0 = none yet
    float *qsepr;                 // Qsep_byrain (n_cont, 2, n_reg)
    // HydroFlash flood sub-model (:728-743, :1711-1822)
    float fl_dw, fl_dsed, fl_av, fl_cmax, fl_umbral, fl_step, fl_hmax;
    int fl_S;                                   // rows of flood_sections
    const float *fl_w, *fl_d50, *fl_slope;      // [N]
    const float *fl_sections, *fl_cells;        // [N][S] (f2py (S,N))
    float *fl_out;                              // [8][N]: h, ufr, Cr, Q, Qsed, speed, rdf, dif of the cell's last evaluation
    int32_t *fl_last;                           // [N] step of the cell's last evaluation, -1 = never
    float *fl_prof;                             // [N][S] flood_profundidad
    unsigned long long *fl_key;                 // [N] ((step * N + writer cell) << 2 | value): the last writer of flood_flood wins
    // separate_fluxes (:658-680, :697-704, :752-759): shares of runoff / subsurface / base flow in the channel tank
    float *flx;             // [3][N][M]
    float4 *flxout4;        // [2][K][N][M] (share sent downstream x3, flag: 1 add, 0 "receiver's shares := 0", 2 hillslope cell)
    float *qsep;            // (n_cont, 3, n_reg)
};

__device__ __forceinline__ size_t sidx(const ShiaP &P, int k, int cell, int m) {
    return ((size_t)k * P.N + cell) * P.M + m;
}

// x / 3 and x / 1000, correctly rounded without the generic division sequence: q0 = x * (1/c), one fma for the exact
// remainder, one to correct (tools/check_pow.cpp verifies both against x / c over every finite float with |q| >= 1e-30).
__device__ __forceinline__ float div_by_3(float x) {
    if (!(fabsf(x) >= 1e-29f && fabsf(x) < 1e37f)) return x / 3.0f;
    const float rc = 1.0f / 3.0f, q0 = x * rc;
    return fmaf(fmaf(-3.0f, q0, x), rc, q0);
}
__device__ __forceinline__ float div_by_1000(float x) {
    if (!(fabsf(x) >= 1e-26f && fabsf(x) < 1e37f)) return x / 1000.0f;
    const float rc = 1.0f / 1000.0f, q0 = x * rc;
    return fmaf(fmaf(-1000.0f, q0, x), rc, q0);
}

// modelosv2.f90:1278-1293
__device__ __forceinline__ float calc_speed_dev(float sm, float coef, float expo, float elem_long, float &speed,
                                                float dt, int niter) {
    float area = 0.0f;
    for (int i = 0; i < niter; ++i) {
        area = sm / (elem_long + speed * dt);
        const float new_speed = coef * wmf_powf(area, expo);
        speed = div_by_3(2.0f * new_speed + speed);
    }
    return area;
}

struct SedLocal {
    float Vs[3], Vd[3], Vsc[3], Vdc[3];
    float up[12];  // what this cell sends to its parent this step
    float DEP[3];
    float ero_add[3], dep_add[3];  // contributions to EROt / DEPt
    float volero_add, voldepo_add;
};

// modelosv2.f90:1321-1428
// aSo = alfa * So**1.664, the first product of Qskr (:1339): constant per cell, the callers compute it once per time block
__device__ __forceinline__ void sed_hillslope_dev(const ShiaP &P, SedLocal &S, float aSo, float v2, float S2,
                                                  float qlin, int cell, int tipo) {
    const float dt = P.dt, dxp = P.dxp;
    float qsSUStot = 0.f, qsBMtot = 0.f, qsSUS[3] = {0, 0, 0}, qsBM[3] = {0, 0, 0}, qsERO[3] = {0, 0, 0};
    float SUStot = 0.f, DEPtot = 0.f, Te[3];
    S.DEP[0] = S.DEP[1] = S.DEP[2] = 0.f;
    // the two powers go through double precision: totXSScap / REScap below are differences of nearly equal terms, and a
    // 1.4 ulp (wmf_powf) or 4 ulp (CUDA powf) error in Qskr reaches 1e-5 in Vs / VolDEPo (tests/test_pow_emulation.py)
    const float Qskr = aSo * wmf_powf_dd(qlin, 2.035f) * P.krus[cell] * P.crus[cell] * P.prus[cell] * dxp * dt;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (S2 / 1000.0f > P.wi[i] * dt)
            Te[i] = P.wi[i] * dt * 1000.0f / S2;
        else
            Te[i] = 1.0f;
        S.Vd[i] = S.Vd[i] + Te[i] * S.Vs[i];
        S.Vs[i] = S.Vs[i] - Te[i] * S.Vs[i];
        S.DEP[i] = Te[i] * S.Vs[i];
        SUStot = SUStot + S.Vs[i];
        DEPtot = DEPtot + S.Vd[i];
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (S.Vs[i] > 0.0f) {
            if (Qskr < SUStot) {
                const float cap = Qskr * S.Vs[i] / SUStot;
                const float Adv = S.Vs[i] * v2 * dt / (dxp + v2 * dt);
                qsSUS[i] = fmaxf(cap, Adv);
            } else {
                qsSUS[i] = S.Vs[i];
            }
        }
        qsSUStot = qsSUStot + qsSUS[i];
        S.Vs[i] = S.Vs[i] - qsSUS[i];
        if (tipo == 1) S.up[i] = qsSUS[i];
        else S.Vsc[i] = S.Vsc[i] + qsSUS[i];
    }
    const float totXSScap = fmaxf(0.0f, Qskr - qsSUStot);
    if (totXSScap > 0.0f && DEPtot > 0.0f) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (S.Vd[i] > 0.0f) {
                if (totXSScap < DEPtot) qsBM[i] = totXSScap * S.Vd[i] / DEPtot;
                else qsBM[i] = S.Vd[i];
            }
            qsBMtot = qsBMtot + qsBM[i];
            S.DEP[i] = S.DEP[i] - qsBM[i];
            if (S.DEP[i] < 0.f) S.DEP[i] = 0.f;
            S.Vd[i] = S.Vd[i] - qsBM[i];
            if (tipo == 1) S.up[3 + i] = qsBM[i];
            else S.Vsc[i] = S.Vsc[i] + qsBM[i];
        }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) S.dep_add[i] += S.DEP[i];
    const float REScap = fmaxf(0.0f, totXSScap - qsBMtot);
    if (REScap > 0.0f) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            qsERO[i] = P.parliac[3 * (size_t)cell + i] * REScap / 100.0f;
            if (tipo == 1) S.up[6 + i] = qsERO[i];
            else S.Vsc[i] = S.Vsc[i] + qsERO[i];
        }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) S.ero_add[i] += qsERO[i];
    S.volero_add = (qsERO[0] + qsERO[1]) + qsERO[2];
    S.voldepo_add = (S.DEP[0] + S.DEP[1]) + S.DEP[2];
}

// modelosv2.f90:1537-1608 ; returns the second VolDEPo increment of the step through *voldepo2
__device__ __forceinline__ void sed_channel_dev(const ShiaP &P, SedLocal &S, float S5, float v5, float Q5, float So,
                                                float area_sec, float VolSal[3], float *voldepo2) {
    const float dt = P.dt, dxp = P.dxp;
    const float grav = 9.8f, Gsed = 2.65f;
    float Cw[3], Te[3], qsSUS[3] = {0, 0, 0}, qsBM[3] = {0, 0, 0};
    float L, Rh;
    S.DEP[0] = S.DEP[1] = S.DEP[2] = 0.f;
    if (S5 > 0.0f) {
        L = area_sec * 1000.0f / S5;
        Rh = L * S5 / (1000.0f * L + 2.0f * S5);
    } else {
        Rh = 0.0f;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float d = P.diam[i];
        Cw[i] = 0.05f * (Gsed / (Gsed - 1.0f)) * (v5 * So) / (sqrtf((Gsed - 1.0f) * grav * d / 1000.0f)) *
                sqrtf((Rh * So) / ((Gsed - 1.0f) * (d / 1000.0f)));
        if (S5 / 1000.0f > P.wi[i] * dt)
            Te[i] = P.wi[i] * dt * 1000.0f / S5;
        else
            Te[i] = 1.0f;
        S.Vdc[i] = S.Vdc[i] + Te[i] * S.Vsc[i];
        S.Vsc[i] = S.Vsc[i] - Te[i] * S.Vsc[i];
        S.DEP[i] = S.DEP[i] + Te[i] * S.Vsc[i];
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float supply = S.Vsc[i] + S.Vdc[i];
        qsSUS[i] = 0.0f;
        if (supply > 0.0f) {
            const float VolEH = Q5 * Cw[i] * dt / 2.65f;
            const float AdvF = fminf(1.0f, v5 * dt / (dxp + v5 * dt));
            qsSUS[i] = S.Vsc[i] * AdvF;
            const float XSScap = fmaxf(0.0f, VolEH - qsSUS[i]);
            qsBM[i] = S.Vdc[i] * AdvF;
            qsBM[i] = fminf(XSScap, qsBM[i]);
            S.DEP[i] = S.DEP[i] - qsBM[i];
            if (S.DEP[i] < 0.0f) S.DEP[i] = 0.0f;
            S.Vsc[i] = S.Vsc[i] - qsSUS[i];
            S.Vdc[i] = S.Vdc[i] - qsBM[i];
            S.up[9 + i] = qsSUS[i] + qsBM[i];
            VolSal[i] = (qsSUS[i] + qsBM[i]) / dt;
        }
    }
    *voldepo2 = (S.DEP[0] + S.DEP[1]) + S.DEP[2];
}

__device__ __forceinline__ int find_row(const int32_t *__restrict__ cells, int n, int cell) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cells[mid] < cell) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// The channel kinematic-wave solve (calc_niter x {div, powf, div}) is ~10x the cost of everything else a cell does,
// and channel cells are a few percent of a basin scattered among hillslope cells: solved in place, most warps would
// run the whole solver for one or two lanes.  Instead the block compacts its channel cells into shared memory and
// the first ceil(n/32) warps solve them with full lanes; results travel back through shared memory.
template <bool SED, bool SLIDES>
__global__ void __launch_bounds__(WAVE_THREADS, WAVE_MIN_BLOCKS)
k_shia_wave(const __grid_constant__ ShiaP P, const int w, const int lo, const int hi) {
    __shared__ float4 s_in[WAVE_THREADS];  // sm, coef, expo, stream_long
    __shared__ float2 s_io[WAVE_THREADS];  // in: x = speed ; out: x = section area, y = speed
    __shared__ int s_wcnt[WAVE_THREADS / 32];
    const long long g = (long long)blockIdx.x * WAVE_THREADS + threadIdx.x;
    int cell, m;
    if (P.M == 1) { cell = lo + (int)g; m = 0; }
    else { cell = lo + (int)(g / P.M); m = (int)(g % P.M); }
    bool active = cell < hi;
    int t = -1;
    uint32_t tw = 0;
    if (active) {
        tw = __ldg(P.tw + cell);
        t = w - (P.Lmax - (int)(tw >> 8));
        active = (t >= 0) && (t < P.T);
    }
    const int N = P.N, M = P.M;
    const int type = (tw >> 4) & 3, nch = tw & 15;
    const float dt = P.dt;
    const size_t NM = (size_t)N * M;
    const size_t base = (size_t)cell * M + m;
    double r_sal = 0.0, r_ret = 0.0;
    double r_vf[4] = {0.0, 0.0, 0.0, 0.0};  // this step's vertical fluxes (mean_vfluxes)
    float sed_ero[3] = {0, 0, 0}, sed_dep[3] = {0, 0, 0};
    float S[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    float h[4] = {0.f, 0.f, 0.f, 0.f};
    float hs[4] = {0.f, 0.f, 0.f, 0.f};
    float m3 = 1.0f, H2 = 0.f;
    SedLocal SL;
    // =========================== phase A: everything up to the channel solve ===========================
    if (active) {
#pragma unroll
        for (int k = 0; k < 5; ++k) S[k] = P.sto[k * NM + base];
        if (SED) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                SL.Vs[k] = P.sed[(0 + k) * NM + base];
                SL.Vd[k] = P.sed[(3 + k) * NM + base];
                SL.Vsc[k] = P.sed[(6 + k) * NM + base];
                SL.Vdc[k] = P.sed[(9 + k) * NM + base];
            }
#pragma unroll
            for (int k = 0; k < 12; ++k) SL.up[k] = 0.f;
            SL.ero_add[0] = SL.ero_add[1] = SL.ero_add[2] = 0.f;
            SL.dep_add[0] = SL.dep_add[1] = SL.dep_add[2] = 0.f;
            SL.volero_add = SL.voldepo_add = 0.f;
        }
        // ---- gather what the upstream neighbours sent this step, in ascending cell order (:601,:617,:672) ----
        if (nch) {
            const int c0 = __ldg(P.cs + cell);
            const float *ob = P.outb + (size_t)(t & 1) * 4 * NM;
            for (int c = c0; c < c0 + nch; ++c) {
                const size_t cb = (size_t)c * M + m;
                S[1] = S[1] + ob[0 * NM + cb];
                S[2] = S[2] + ob[1 * NM + cb];
                S[3] = S[3] + ob[2 * NM + cb];
                S[4] = S[4] + ob[3 * NM + cb];
                if (SED) {
                    const float *so = P.sedout + (size_t)(t & 1) * 12 * NM;
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        SL.Vs[i] = SL.Vs[i] + so[(0 + i) * NM + cb];
                        SL.Vs[i] = SL.Vs[i] + so[(3 + i) * NM + cb];
                        SL.Vs[i] = SL.Vs[i] + so[(6 + i) * NM + cb];
                        SL.Vsc[i] = SL.Vsc[i] + so[(9 + i) * NM + cb];
                    }
                }
            }
        }
    }
    float chan_coef = 0.f, chan_expo = 0.f;
    if (active) {
        // ---- per-cell parameters (:285-304), recomputed from the shared maps and the member's calibration ----
        float cal[11];
        {
            const float *cp = P.calib + (size_t)m * 11;
#pragma unroll
            for (int k = 0; k < 11; ++k) cal[k] = __ldg(cp + k);
        }
        const float4 vc = __ldg(reinterpret_cast<const float4 *>(P.v_coef) + cell);
        const float4 hc = __ldg(reinterpret_cast<const float4 *>(P.h_coef) + cell);
        const float vs1 = vc.x * cal[0] * dt, vs2 = vc.y * cal[1] * dt, vs3 = vc.z * cal[2] * dt, vs4 = vc.w * cal[3] * dt;
        const float H1 = __ldg(P.max_capilar + cell) * cal[8];
        const float Lh = __ldg(P.hill_long + cell);
        m3 = __ldg(P.elem_area + cell) / 1000.0f;
        // ---- rain (:444-449) ----
        const int p = __ldg(P.pos + t);
        float R = 0.0f;
        if (p != 1) R = (float)__ldg(P.rain + (size_t)(p - 1) * N + cell) / 1000.0f;
        // ---- vertical cascade (:478-512) ----
        float vf1 = fmaxf(0.0f, R - H1 + S[0]);
        S[0] = S[0] + R - vf1;
        const float E = fminf(__ldg(P.evp + t) * vs1 * wmf_powf(S[0] / H1, 0.6f), S[0]);
        S[0] = S[0] - E;
        float vf2 = fminf(vf1, vs2);
        S[1] = S[1] + vf1 - vf2;
        float vf3 = fminf(vf2, vs3);
        S[2] = S[2] + vf2 - vf3;
        const float vf4 = fminf(vf3, vs4);
        S[3] = S[3] + vf3 - vf4;
        if (P.retorno_gr > 0.0f || SLIDES) H2 = __ldg(P.max_gravita + cell) * cal[9];
        if (P.retorno_aq > 0.0f) {
            const float H3 = __ldg(P.max_aquifer + cell) * cal[10];
            const float Ret_aq = fmaxf(0.0f, S[3] - H3);
            S[2] = S[2] + Ret_aq;
            S[3] = S[3] - Ret_aq;
        }
        float ret_step = 0.0f;
        if (P.retorno_gr > 0.0f) {
            const float Ret = fmaxf(0.0f, S[2] - H2);
            S[1] = S[1] + Ret;
            S[2] = S[2] - Ret;
            ret_step = Ret;
            if (P.want_retorned) {
                if (P.retorned) P.retorned[cell] = P.retorned[cell] + Ret;
                r_ret = (double)Ret;
            }
        }
        if (P.want_vmean | (P.vfl_snap != nullptr) | (P.rc != nullptr) | (P.ret_snap != nullptr)) {
            // the dumps that follow the vertical fluxes; vflux(2:3) carry the return flow (:509-524)
            const float vx2 = (P.retorno_gr > 0.0f) ? vf2 + ret_step : vf2;
            const float vx3 = (P.retorno_gr > 0.0f) ? vf3 - ret_step : vf3;
            r_vf[0] = (double)vf1; r_vf[1] = (double)vx2; r_vf[2] = (double)vx3; r_vf[3] = (double)vf4;
            if (P.rc) {
                P.rc[cell] = P.rc[cell] + vx2 - vx3;
                P.rc[(size_t)N + cell] = P.rc[(size_t)N + cell] + R;
            }
            if (P.ret_snap) P.ret_snap[(size_t)t * N + cell] = ret_step;
            if (P.vfl_snap) {
                const int sl = __ldg(P.vsnap_slot + t);
                if (sl >= 0) reinterpret_cast<float4 *>(P.vfl_snap)[(size_t)sl * N + cell] = make_float4(vf1, vx2, vx3, vf4);
            }
        }
        r_sal = (double)vf4 + (double)E;
        // ---- hillslope outflows of tanks 2..4 (:553-595) ----
        const float hcoef[3] = {hc.x, hc.y, hc.z};
        float4 he = make_float4(0.f, 0.f, 0.f, 0.f);
        const bool any_kin = (P.st[0] == 2) | (P.st[1] == 2) | (P.st[2] == 2);
        if (any_kin || type > 1) he = __ldg(reinterpret_cast<const float4 *>(P.h_exp) + cell);
        const float hexp[3] = {he.x, he.y, he.z};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (P.st[k] == 1) {
                const float v = P.hs_state ? P.hsp[k * NM + base] : hcoef[k] * cal[4 + k];
                hs[k] = v;
                h[k] = (1.0f - Lh / (v * dt + Lh)) * S[k + 1];
            } else {
                float v = P.hsp[k * NM + base];
                const float sec = calc_speed_dev(S[k + 1] * m3, hcoef[k] * cal[4 + k], hexp[k], Lh, v, dt, P.niter);
                h[k] = fminf(sec * v * dt / m3, S[k + 1]);
                hs[k] = v;
                P.hsp[k * NM + base] = v;
                if (SED && k == 0) {
                    const float qlin = h[0] * m3 / (dt * P.dxp);
                    sed_hillslope_dev(P, SL, P.sed_factor * wmf_powf_dd(__ldg(P.hill_slope + cell), 1.664f), v, S[1], qlin, cell, type);
                }
            }
            S[k + 1] = S[k + 1] - h[k];
        }
        if (type > 1) {  // channel cell: what stays in the channel tank before the kinematic solve (:617-621)
            S[4] = S[4] + (h[0] + h[1]) + h[2] * (float)(type - 2);
            chan_coef = hc.w * cal[7];
            chan_expo = he.w;
        }
    }
    // =========================== channel solve, compacted over the block (:631-635) ===========================
    const bool is_ch = active && type > 1;
    float sec_area = 0.f;
    {
        const unsigned full = 0xffffffffu;
        const unsigned bal = __ballot_sync(full, is_ch);
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        if (lane == 0) s_wcnt[wid] = __popc(bal);
        __syncthreads();
        int woff = 0, total = 0;
#pragma unroll
        for (int k = 0; k < WAVE_THREADS / 32; ++k) {
            const int c = s_wcnt[k];
            if (k < wid) woff += c;
            total += c;
        }
        if (total > 0) {  // block-uniform
            const int slot = woff + __popc(bal & ((1u << lane) - 1u));
            if (is_ch) {
                s_in[slot] = make_float4(S[4] * m3, chan_coef, chan_expo, __ldg(P.stream_long + cell));
                s_io[slot] = make_float2(P.hsp[3 * NM + base], 0.f);
            }
            __syncthreads();
            for (int j = threadIdx.x; j < total; j += WAVE_THREADS) {
                const float4 a = s_in[j];
                float v = s_io[j].x;
                const float ar = calc_speed_dev(a.x, a.y, a.z, a.w, v, dt, P.niter);
                s_io[j] = make_float2(ar, v);
            }
            __syncthreads();
            if (is_ch) {
                const float2 r = s_io[slot];
                sec_area = r.x;
                hs[3] = r.y;
            }
        }
    }
    // =========================== phase B: routing, records, publish ===========================
    if (active) {
        // ---- routing (:599-719) ----
        const bool outlet = (cell == N - 1);
        float o0, o1, o2, o3;
        float Vsal[3] = {0.f, 0.f, 0.f};
        const size_t qoff = ((size_t)m * P.T + t) * P.n_cont;
        if (type == 1) {
            o0 = h[0]; o1 = h[1]; o2 = h[2]; o3 = 0.0f;
            if (outlet) {
                const float s3 = (h[0] + h[1]) + h[2];
                if (P.q) P.q[qoff] = s3 * m3;
                r_sal += (double)s3;
            }
        } else {
            o0 = 0.0f; o1 = 0.0f;
            o2 = h[2] * (float)(3 - type);
            const float v = hs[3];
            h[3] = fminf(sec_area * v * dt / m3, S[4]);
            P.hsp[3 * NM + base] = v;
            if (SED) {
                float vd2 = 0.f;
                sed_channel_dev(P, SL, S[4], v, h[3] * m3 / dt, __ldg(P.stream_slope + cell), sec_area, Vsal, &vd2);
                // VolDEPo(celda) is incremented twice in a step for a channel cell (:1427 then :1607)
                float vdep = P.voldepo[base];
                vdep = vdep + SL.voldepo_add;
                vdep = vdep + vd2;
                P.voldepo[base] = vdep;
                P.volero[base] = P.volero[base] + SL.volero_add;
                SL.voldepo_add = 0.f;
                SL.volero_add = 0.f;
            }
            S[4] = S[4] - h[3];
            o3 = h[3];
            if (outlet) {
                if (P.q) P.q[qoff] = h[3] * m3 / dt;
                r_sal += (double)h[3];
                if (SED && P.qsed)
#pragma unroll
                    for (int i = 0; i < 3; ++i) P.qsed[(((size_t)m * P.T + t) * 3 + i) * P.n_cont] = Vsal[i];
                if (P.show_speed) {
                    if (P.speed) P.speed[qoff] = v;
                    if (P.area) P.area[qoff] = sec_area;
                }
            }
        }
        if (SED && type == 1 && P.st[0] == 2) {  // hillslope cell: single VolDEPo / VolERO increment (:1426-1427)
            P.voldepo[base] = P.voldepo[base] + SL.voldepo_add;
            P.volero[base] = P.volero[base] + SL.volero_add;
        }
        // ---- landslides (:723, :1649-1707) ----
        if (SLIDES) {
            if (__ldg(P.sl_risk + cell) == 2.0f && P.sl_slid[cell] == 0 && P.sl_marktime[cell] > t) {
                const float zs = __ldg(P.sl_zs + cell);
                const float Zw = (S[2] * zs) / H2;
                bool fail = false;
                if (Zw >= __ldg(P.sl_zcrit + cell)) {
                    fail = true;
                } else {
                    const float gs = __ldg(P.sl_gammas + cell);
                    const float cr = __ldg(P.sl_cosr + cell), sr = __ldg(P.sl_sinr + cell);
                    const float Num = __ldg(P.sl_cohesion + cell) + (gs * zs - Zw * P.sl_gammaw) * (cr * cr) * __ldg(P.sl_tanfa + cell);
                    const float Den = gs * zs * sr * cr;
                    if (Num / Den <= P.sl_fs) fail = true;
                }
                if (fail) {
                    P.sl_slid[cell] = 1;
                    atomicAdd(P.sl_acum + cell, 1);
                    int marks = 1;
                    if (P.sl_gullie == 1 && type <= 2) {  // slide_hill2gullie :1683-1707
                        int local = cell, ltype = type;
                        for (;;) {
                            const int d = P.drena[(size_t)local * P.drena_stride];
                            if (d == 0) break;
                            const int dir = N - d;
                            const int dtype = (__ldg(P.tw + dir) >> 4) & 3;
                            if (dtype > ltype) break;
                            atomicMin(P.sl_marktime + dir, t);
                            atomicAdd(P.sl_acum + dir, 1);
                            ++marks;
                            local = dir;
                            ltype = dtype;
                        }
                    }
                    atomicAdd(P.sl_ncelltime + t, marks);
                }
            }
        }
        // ---- records (:749-787) ----
        if (tw & TW_CTRLQ) {
            const int row = 1 + find_row(P.ctrl_cells, P.n_ctrl, cell);
            if (row < P.n_cont) {
                if (P.q) P.q[qoff + row] = (type > 1) ? h[3] * m3 / dt : 0.0f;
                if (type > 1) {
                    if (P.show_speed) {
                        if (P.speed) P.speed[qoff + row] = hs[3];
                        if (P.area) P.area[qoff + row] = sec_area;
                    }
                    if (SED && P.qsed)
#pragma unroll
                        for (int i = 0; i < 3; ++i) P.qsed[(((size_t)m * P.T + t) * 3 + i) * P.n_cont + row] = Vsal[i];
                }
            }
        }
        if (tw & TW_CTRLH) {
            const int row = find_row(P.ctrlh_cells, P.n_ctrlh, cell);
            if (row < P.n_conth) {
                const size_t k = ((size_t)m * P.T + t) * P.n_conth + row;
                if (P.hum) P.hum[k] = S[0] + S[2];
                if (P.st1o) P.st1o[k] = S[0];
                if (P.st3o) P.st3o[k] = S[2];
            }
        }
        // ---- publish ----
#pragma unroll
        for (int k = 0; k < 5; ++k) P.sto[k * NM + base] = S[k];
        if (P.sto_snap) {  // save_storage (:799-803)
            const int sl = __ldg(P.snap_slot + t);
            if (sl >= 0) {
#pragma unroll
                for (int k = 0; k < 5; ++k) P.sto_snap[((size_t)sl * N + cell) * 5 + k] = S[k];
            }
        }
        if (P.spd_snap)    // save_speed (:815-818)
            reinterpret_cast<float4 *>(P.spd_snap)[(size_t)t * N + cell] =
                make_float4(hs[0], hs[1], hs[2], (type > 1) ? hs[3] : P.hsp[3 * NM + base]);
        {
            float *ob = P.outb + (size_t)(t & 1) * 4 * NM;
            ob[0 * NM + base] = o0;
            ob[1 * NM + base] = o1;
            ob[2 * NM + base] = o2;
            ob[3 * NM + base] = o3;
        }
        if (SED) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                P.sed[(0 + k) * NM + base] = SL.Vs[k];
                P.sed[(3 + k) * NM + base] = SL.Vd[k];
                P.sed[(6 + k) * NM + base] = SL.Vsc[k];
                P.sed[(9 + k) * NM + base] = SL.Vdc[k];
            }
            float *so = P.sedout + (size_t)(t & 1) * 12 * NM;
#pragma unroll
            for (int k = 0; k < 12; ++k) so[k * NM + base] = SL.up[k];
#pragma unroll
            for (int i = 0; i < 3; ++i) { sed_ero[i] = SL.ero_add[i]; sed_dep[i] = SL.dep_add[i]; }
        }
    }
    // ---- per-step reductions (balance :846, mean_storage :828, mean_speed :833, mean_retorno :838) ----
    if ((P.want_red | P.want_vmean) && P.M == 1) {
        // steps are monotone along the cell index, so a warp (usually a whole block) shares one step
        const unsigned full = 0xffffffffu;
        const int t0 = __shfl_sync(full, t, 0);
        const bool uni = __all_sync(full, t == t0);
        const int sub = blockIdx.x & (RED_SUB - 1);
        const bool lead = (threadIdx.x & 31) == 0;
        auto flush = [&](double v, int q, int tt) {
            if (v != 0.0) atomicAdd(P.red + ((size_t)tt * RQ_COUNT + q) * RED_SUB + sub, v);
        };
        // sum(hspeed,dim=2): linear tanks hold their constant speed, the channel speed of a hillslope cell keeps its
        // initial value
        auto spd = [&](int k) -> double {
            if (!active) return 0.0;
            if (k < 3) return (double)hs[k];
            return (double)((type > 1) ? hs[3] : (P.hs_state ? P.hsp[3 * NM + base] : 0.0f));
        };
        if (uni) {
            if (t0 >= 0) {
                const double a = warp_sum_d(r_sal);
                if (lead) flush(a, RQ_SAL, t0);
                if (P.show_storage) {
#pragma unroll
                    for (int k = 0; k < 5; ++k) {
                        const double b = warp_sum_d((double)S[k]);
                        if (lead) flush(b, RQ_STO1 + k, t0);
                    }
                } else {
                    const double b = warp_sum_d(((double)S[0] + (double)S[1]) + ((double)S[2] + (double)S[3]) + (double)S[4]);
                    if (lead) flush(b, RQ_STO1, t0);
                }
                if (P.show_mean_speed) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const double c = warp_sum_d(spd(k));
                        if (lead) flush(c, RQ_SPD1 + k, t0);
                    }
                }
                if (P.want_retorned) {
                    const double c = warp_sum_d(r_ret);
                    if (lead) flush(c, RQ_RET, t0);
                }
                if (P.want_vmean) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const double c = warp_sum_d(r_vf[k]);
                        if (lead) flush(c, RQ_VFL1 + k, t0);
                    }
                }
            }
        } else if (t >= 0 && active) {
            flush(r_sal, RQ_SAL, t);
            if (P.show_storage) {
#pragma unroll
                for (int k = 0; k < 5; ++k) flush((double)S[k], RQ_STO1 + k, t);
            } else {
                flush(((double)S[0] + (double)S[1]) + ((double)S[2] + (double)S[3]) + (double)S[4], RQ_STO1, t);
            }
            if (P.show_mean_speed)
#pragma unroll
                for (int k = 0; k < 4; ++k) flush(spd(k), RQ_SPD1 + k, t);
            if (P.want_retorned) flush(r_ret, RQ_RET, t);
            if (P.want_vmean)
#pragma unroll
                for (int k = 0; k < 4; ++k) flush(r_vf[k], RQ_VFL1 + k, t);
        }
    }
    if (SED) {  // EROt / DEPt (:1403, :1420): global totals, fp64
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const double a = warp_sum_d((double)sed_ero[i]);
            const double b = warp_sum_d((double)sed_dep[i]);
            if ((threadIdx.x & 31) == 0) {
                if (a != 0.0) atomicAdd(P.sedtot + i, a);
                if (b != 0.0) atomicAdd(P.sedtot + 3 + i, b);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Time-blocked wavefront.  One thread owns a (cell, member) for K consecutive steps: its state stays in registers,
// the per-cell parameters are derived once per block of steps, and only the per-step traffic (children's outflow,
// rain, own outflow) touches memory -- produced one wave earlier, so it is served from L2.  Cell (level L, time
// block b) runs in wave w = b + (Lmax - L): its children finished block b one wave earlier, and so did its own
// block b-1.  Hillslope and channel cells run in different CTAs of the same launch (lists of cell ids per role),
// so no warp mixes the cheap hillslope update with the channel's kinematic-wave iterations.
// ---------------------------------------------------------------------------------------------------
#ifndef TB_THREADS_N
#define TB_THREADS_N 256
#endif
constexpr int TB_THREADS = TB_THREADS_N;
#ifndef TB_MIN_BLOCKS
#define TB_MIN_BLOCKS 4
#endif
#ifndef TB_MIN_BLOCKS_KIN   // kinematic hillslope tanks (calc_speed: five dependent pow per tank and step): 85 registers instead
#define TB_MIN_BLOCKS_KIN 3 // of 64 + 268 B of spills; C2 with speed_type = [2,2,1], 400 steps: 33.0 -> 31.3 ms (2 blocks: 37.6)
#endif
// Asynchronous staging of a time block's inputs (compile with -DWMF_TB_STAGE=1): everything a thread reads during its K
// steps -- the float4 its first two tributaries sent at each step and the K rain values -- was written by the previous
// wave, so it can be requested with cp.async right after griddepcontrol.wait and land in shared memory while the first
// step computes, holding no registers.  Blocks of up to 4 steps (36 KB per 256-thread CTA at K = 4).
// MEASURED AND REJECTED (round 2, profiles/README.md): C2 wave loop 28.47 -> 30.06 ms.  The staged kernel executes 12 %
// more warp instructions per CTA (copy issue, wait ladders, shared loads) and its long-scoreboard stalls per issued
// instruction do not fall (3.5 before and after): the tributary and rain loads were already hidden behind the surface
// tank's pow chain; what remains are the state / static-map loads at the head of a block.  Off by default.
#ifndef WMF_TB_STAGE
#define WMF_TB_STAGE 0
#endif
__device__ __forceinline__ void cp_async_16(void *smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_4(void *smem, const void *gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
template <int MAXN>
__device__ __forceinline__ void cp_async_wait_dyn(int n) {  // wait until at most n (0..MAXN) groups are pending
    if constexpr (MAXN > 0) {
        if (n >= MAXN) { cp_async_wait<MAXN>(); return; }
        cp_async_wait_dyn<MAXN - 1>(n);
    } else {
        cp_async_wait<0>();
    }
}
template <int K>
struct TbStage {
    static constexpr bool on = (WMF_TB_STAGE != 0) && (K <= 4);
};

#ifndef TB_MIN_BLOCKS_SED   // separate_rain variants (and, until the sediments got their own kernel, the SED ones)
#define TB_MIN_BLOCKS_SED 2
#endif
#ifndef SED_MIN_BLOCKS      // k_sed_tb
#define SED_MIN_BLOCKS 3
#endif

// x**y and log(x) as the reference's libm evaluates them in fp32 (glibc: correctly rounded in all but a vanishing share of
// cases): double precision, rounded once.  The flood loop compares against them, so a last-bit difference would change counts.
__device__ __forceinline__ float powf_cr(float x, float y) { return (float)pow((double)x, (double)y); }
__device__ __forceinline__ float logf_cr(float x) { return (float)log((double)x); }

// flood_params (:1730-1749) + flood_debris_flow2 (:1786-1822) for one type-3 channel cell at one step
__device__ __noinline__ void flood_dev(const ShiaP &P, int cell, int t, float Q, float speed) {
    const int N = P.N, S = P.fl_S;
    const float d50 = __ldg(P.fl_d50 + cell), cmax = P.fl_cmax;
    float h = Q / (speed * __ldg(P.fl_w + cell));
    if (h > P.fl_hmax) h = P.fl_hmax;
    const float ufr = speed / (5.75f * logf_cr(h / d50) + 6.25f);
    const float cr = cmax * powf_cr(0.06f * h, 0.2f / ufr);
    const float rdf = (1.0f / d50) * powf_cr((9.81f / 0.0128f) * (cr + (1.0f - cr) * (P.fl_dw / P.fl_dsed)), 0.5f) *
                      (powf_cr(cmax / cr, 1.0f / 3.0f) - 1.0f);
    const float Qsed = Q * (1.0f + (cr / (1.0f - cr)));
    const float *sec = P.fl_sections + (size_t)S * cell;
    const float fondo = sec[S / 2];
    const float slope = __ldg(P.fl_slope + cell);
    float Qp = 0.0f, dif = 0.0f, linea = 0.0f;
    int i = 1;
    bool ran = false;
    while (Qp < Qsed && (float)i * P.fl_step < 50.0f) {
        dif = (float)i * P.fl_step;
        linea = fondo + dif;
        float sum = 0.0f;
        for (int k = 0; k < S; ++k) {
            const float d = linea - sec[k];
            if (d > 0.0f) sum = sum + d;
        }
        const float area = sum * P.dxp;
        Qp = ((2.0f / 5.0f) * rdf * powf_cr((float)i * P.fl_step, 3.0f / 2.0f) * slope * (1.0f / 2.0f)) * area;
        i = i + 1;
        ran = true;
    }
    // flood_profundidad(:,celda) = diferencias of the last trip (zeros when the loop never ran); the flooded cells of the
    // section get 1, the channel cell itself 2; later (step, cell) events overwrite earlier ones
    const unsigned long long ev = ((unsigned long long)t * (unsigned long long)N + (unsigned long long)cell) << 2;
    float *prof = P.fl_prof + (size_t)S * cell;
    const float *cells = P.fl_cells + (size_t)S * cell;
    for (int k = 0; k < S; ++k) {
        const float d = ran ? linea - sec[k] : 0.0f;
        prof[k] = d;
        if (d > 0.0f) {
            const int pos = (int)cells[k];
            if (pos >= 1 && pos <= N) atomicMax(P.fl_key + (pos - 1), ev | 1ull);
        }
    }
    atomicMax(P.fl_key + cell, ev | 2ull);
    float *o = P.fl_out;
    o[cell] = h; o[(size_t)N + cell] = ufr; o[2 * (size_t)N + cell] = cr; o[3 * (size_t)N + cell] = Q;
    o[4 * (size_t)N + cell] = Qsed; o[5 * (size_t)N + cell] = speed; o[6 * (size_t)N + cell] = rdf; o[7 * (size_t)N + cell] = dif;
    P.fl_last[cell] = t;
}

template <int K, bool CHAN, bool KIN, bool SED, bool SLIDES, bool SNAP, bool SEPF, bool SEPR, bool FLOOD>
__device__ __forceinline__ void tb_role(const ShiaP &P, const int w, const int jlo, const int jhi, const int tile,
                                        float *__restrict__ s_acc, float4 *__restrict__ s_kid, int32_t *__restrict__ s_rain) {
    constexpr bool STAGE = TbStage<K>::on;
    const int N = P.N, M = P.M;
    const long long g = (long long)tile * 32 + (threadIdx.x & 31);
    int j, m;
    if (M == 1) { j = jlo + (int)g; m = 0; }
    else { j = jlo + (int)(g / M); m = (int)(g % M); }
    bool active = j < jhi;
    int cell = 0, b = -1;
    uint32_t tw = 0;
    if (active) {
        cell = __ldg((CHAN ? P.cidx : P.hidx) + j);
        tw = __ldg(P.tw + cell);
        b = w - (P.Lmax - (int)(tw >> 8));
        active = (b >= 0) && (b < P.nb);
    }
    // programmatic dependent launch: let the next wave's CTAs be scheduled now (they stop at their own wait below)
    asm volatile("griddepcontrol.launch_dependents;");
    const int t0 = b * K;
    const int nt = active ? min(K, P.T - t0) : 0;
    const float dt = P.dt;
    const size_t NM = (size_t)N * M;
    const size_t base = (size_t)cell * M + m;
    const int type = (tw >> 4) & 3, nch = tw & 15;
    // ---- per-thread constants of the time block (:285-304) ----
    float S0 = 0.f, S1 = 0.f, S2 = 0.f, S3 = 0.f, S4 = 0.f;
    float vs1 = 0.f, vs2 = 0.f, vs3 = 0.f, vs4 = 0.f, H1 = 1.f, H2 = 0.f, H3 = 0.f, Lh = 1.f, m3 = 1.f;
    float fl[3] = {0.f, 0.f, 0.f};      // linear reservoirs: 1 - L/(v dt + L), constant over the run
    float kv[3] = {0.f, 0.f, 0.f};      // speeds (state for kinematic tanks)
    float kc[3] = {0.f, 0.f, 0.f}, ke[3] = {0.f, 0.f, 0.f};
    float vch = 0.f, ccoef = 0.f, cexpo = 0.f, Ls = 1.f, ret_acc = 0.f;
    int c0 = 0, rowq = -1, rowh = -1;
    if (active) {
        const float *cal = P.calib + (size_t)m * 11;
        const float4 vc = __ldg(reinterpret_cast<const float4 *>(P.v_coef) + cell);
        const float4 hc = __ldg(reinterpret_cast<const float4 *>(P.h_coef) + cell);
        vs1 = vc.x * __ldg(cal + 0) * dt; vs2 = vc.y * __ldg(cal + 1) * dt;
        vs3 = vc.z * __ldg(cal + 2) * dt; vs4 = vc.w * __ldg(cal + 3) * dt;
        H1 = __ldg(P.max_capilar + cell) * __ldg(cal + 8);
        if (P.retorno_gr > 0.0f || SLIDES) H2 = __ldg(P.max_gravita + cell) * __ldg(cal + 9);
        if (P.retorno_aq > 0.0f) H3 = __ldg(P.max_aquifer + cell) * __ldg(cal + 10);
        Lh = __ldg(P.hill_long + cell);
        m3 = __ldg(P.elem_area + cell) / 1000.0f;
        float4 he = make_float4(0.f, 0.f, 0.f, 0.f);
        if (KIN || CHAN) he = __ldg(reinterpret_cast<const float4 *>(P.h_exp) + cell);
        const float hcoef[3] = {hc.x, hc.y, hc.z}, hexp[3] = {he.x, he.y, he.z};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (!KIN || P.st[k] == 1) {
                if (!P.hs_state) {
                    kv[k] = hcoef[k] * __ldg(cal + 4 + k);
                    fl[k] = 1.0f - Lh / (kv[k] * dt + Lh);
                }
            } else {
                kc[k] = hcoef[k] * __ldg(cal + 4 + k);
                ke[k] = hexp[k];
            }
        }
        if (CHAN) {
            ccoef = hc.w * __ldg(cal + 7);
            cexpo = he.w;
            Ls = __ldg(P.stream_long + cell);
        }
        if (nch) c0 = __ldg(P.cs + cell);
        if (tw & TW_CTRLQ) {
            rowq = 1 + find_row(P.ctrl_cells, P.n_ctrl, cell);
            if (rowq >= P.n_cont) rowq = -1;
        }
        if (tw & TW_CTRLH) {
            rowh = find_row(P.ctrlh_cells, P.n_ctrlh, cell);
            if (rowh >= P.n_conth) rowh = -1;
        }
    }
    // everything above reads run constants only; what follows was written by the previous wave
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (active) {
        const float4 s = P.st4[base];
        S0 = s.x; S1 = s.y; S2 = s.z; S3 = s.w;
        S4 = P.s5[base];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (!KIN || P.st[k] == 1) {
                if (P.hs_state) {
                    kv[k] = P.hsp[k * NM + base];
                    fl[k] = 1.0f - Lh / (kv[k] * dt + Lh);
                }
            } else {
                kv[k] = P.hsp[k * NM + base];
            }
        }
        if (CHAN) vch = P.hsp[3 * NM + base];
        if (P.retorned) ret_acc = P.retorned[cell];
    }
    // sediments (:1298-1608) run in their own kernel (k_sed_tb, launched after every wave of this one): this kernel only
    // stages what they read from the hydrology of each step
    float4 *stg = SED ? P.sedstg4 + (size_t)(b & 1) * K * 2 * NM + base : nullptr;
    // slides (:1649-1707): own failure flag and the step at which an upstream failure reached this cell.  Every mark
    // with a step of this block or earlier was made by a wave before this one, so one read per block is enough.
    int sl_slid = 0, sl_mark = INT_MAX;
    int sl_mt_in = INT_MAX, sl_cnt_in = 0, sl_fail_t = INT_MAX;   // walks arriving from upstream in this block; own failure step
    bool sl_risky = false;
    if (SLIDES && active) {
        sl_risky = __ldg(P.sl_risk + cell) == 2.0f;
        sl_slid = P.sl_slid[base];
        sl_mark = P.sl_marktime[base];
        // slide_hill2gullie (:1683-1707) without the walk: a failure marks every cell downstream of it while the unit type
        // does not grow.  Marks travel with the wavefront instead: every cell hands its parent the earliest failure step of
        // this time block among the walks that pass through it, and how many they are; the parent takes them when its own
        // type is not larger than the sender's.  (The tributaries processed this block in the previous wave.)
        if (P.sl_gullie == 1) {
            const int2 *rin = P.sl_ring + (size_t)(b & 1) * NM + (size_t)c0 * M + m;
            for (int c = 0; c < nch; ++c, rin += M) {
                const int2 e = *rin;
                if (type <= (int)((unsigned)e.y >> 30)) {
                    sl_mt_in = min(sl_mt_in, e.x);
                    sl_cnt_in += e.y & 0x3FFFFFFF;
                }
            }
            sl_mark = min(sl_mark, sl_mt_in);
        }
    }
    float rc1 = 0.f, rc2 = 0.f;           // save_rc accumulators
    if (SNAP && active && P.rc) { rc1 = P.rc[cell]; rc2 = P.rc[(size_t)N + cell]; }
    float C0 = 0.f, C1 = 0.f, C2 = 0.f, C3 = 0.f, C4 = 0.f;  // separate_rain: convective shadow of the five tanks
    float T0 = 0.f, T1 = 0.f, T2 = 0.f, T3 = 0.f, T4 = 0.f;  //                stratiform shadow
    if (SEPR && active) {
        C0 = P.shc[base]; C1 = P.shc[NM + base]; C2 = P.shc[2 * NM + base]; C3 = P.shc[3 * NM + base]; C4 = P.shc[4 * NM + base];
        T0 = P.shs[base]; T1 = P.shs[NM + base]; T2 = P.shs[2 * NM + base]; T3 = P.shs[3 * NM + base]; T4 = P.shs[4 * NM + base];
    }
    float F0 = 0.f, F1 = 0.f, F2 = 0.f;  // separate_fluxes: the channel tank's shares
    if (SEPF && CHAN && active) { F0 = P.flx[base]; F1 = P.flx[NM + base]; F2 = P.flx[2 * NM + base]; }
    const bool outlet = active && (cell == N - 1);
    const float4 *ob_in = P.outb4 + (size_t)(b & 1) * K * NM + (size_t)c0 * M + m;
    float4 *ob_out = P.outb4 + (size_t)(b & 1) * K * NM + base;
    if (STAGE) {  // one commit group per step, waited for at the step that consumes it
#pragma unroll
        for (int tt = 0; tt < K; ++tt) {
            if (tt < nt) {
                const int p = __ldg(P.pos + (size_t)m * P.pos_stride + t0 + tt);
                if (p != 1) cp_async_4(s_rain + tt * TB_THREADS + threadIdx.x, P.rain + (size_t)(p - 1) * N + cell);
            }
            cp_async_commit();      // group 2 tt: the rain value of step tt (needed first)
            if (tt < nt) {
                if (nch > 0) cp_async_16(s_kid + (2 * tt) * TB_THREADS + threadIdx.x, ob_in + (size_t)tt * NM);
                if (nch > 1) cp_async_16(s_kid + (2 * tt + 1) * TB_THREADS + threadIdx.x, ob_in + (size_t)tt * NM + M);
            }
            cp_async_commit();      // group 2 tt + 1: what the first two tributaries sent at step tt (needed after the surface tank)
        }
    }
    // ---- the K steps of the block ----
#pragma unroll 1
    for (int tt = 0; tt < K; ++tt) {
        float dq = 0.0f;
        float vx1 = 0.f, vx2 = 0.f, vx3 = 0.f, vx4 = 0.f;  // this step's vertical fluxes (dump variants only)
        if (tt < nt) {
            const int t = t0 + tt;
            const float So0 = S0, So1 = S1, So2 = S2, So3 = S3, So4 = S4;
            // Loads first: the outflow of the first two upstream neighbours and this step's rain are requested before
            // the surface-tank update (which needs only the rain) so that their latency hides behind its pow chain.
            const float4 *ob = ob_in;
            float4 oa = make_float4(0.f, 0.f, 0.f, 0.f), obb = oa;
            if (!STAGE) {
                if (nch > 0) oa = ob[0];
                if (nch > 1) obb = ob[M];
            }
            // rain (:444-449)
            const int p = __ldg(P.pos + (size_t)m * P.pos_stride + t);
            float R = 0.0f;
            if (STAGE) {
                cp_async_wait_dyn<2 * K - 1>(2 * (K - 1 - tt) + 1);
                if (p != 1) R = div_by_1000((float)s_rain[tt * TB_THREADS + threadIdx.x]);
            } else if (p != 1) R = div_by_1000((float)__ldg(P.rain + (size_t)(p - 1) * N + cell));
            // vertical cascade, surface tank (:478-486)
            const float vf1 = fmaxf(0.0f, R - H1 + S0);
            S0 = S0 + R - vf1;
            const float E = fminf(__ldg(P.evp + t) * vs1 * wmf_powf(S0 / H1, 0.6f), S0);
            float Co = 0.0f, St = 0.0f;
            if (SEPR) {
                // rain types in force (:452-456 reads them on wet steps only; :471-474 integer division by 1000)
                const int ec = __ldg(P.effc + t), es = __ldg(P.effs + t);
                if (ec > 0) Co = (float)(__ldg(P.conv + (size_t)(ec - 1) * N + cell) / 1000);
                if (es > 0) St = (float)(__ldg(P.stra + (size_t)(es - 1) * N + cell) / 1000);
                // :485-488, against the capillary tank before the evaporation leaves; the compiled MAX(0., NaN) is 0
                const float a = C0 - E * C0 / S0, bq = T0 - E * T0 / S0;
                C0 = (a > 0.0f) ? a : 0.0f;
                T0 = (bq > 0.0f) ? bq : 0.0f;
            }
            S0 = S0 - E;
            if (STAGE) {
                cp_async_wait_dyn<2 * K - 1>(2 * (K - 1 - tt));
                if (nch > 0) oa = s_kid[(2 * tt) * TB_THREADS + threadIdx.x];
                if (nch > 1) obb = s_kid[(2 * tt + 1) * TB_THREADS + threadIdx.x];
            }
            // what the upstream neighbours sent this step, added in ascending cell order (:601,:617,:672) -- the
            // reference's children deposit into StoOut(2:5) before the cell itself is processed
            if (nch > 0) {
                S1 = S1 + oa.x; S2 = S2 + oa.y; S3 = S3 + oa.z; S4 = S4 + oa.w;
                if (nch > 1) {
                    S1 = S1 + obb.x; S2 = S2 + obb.y; S3 = S3 + obb.z; S4 = S4 + obb.w;
                    ob += 2 * (size_t)M;
#pragma unroll 1
                    for (int c = 2; c < nch; ++c, ob += M) {
                        const float4 o = *ob;
                        S1 = S1 + o.x;
                        S2 = S2 + o.y;
                        S3 = S3 + o.z;
                        S4 = S4 + o.w;
                    }
                }
            }
            if (SEPF && CHAN) {  // what the upstream channel cells did to my shares this step, in cell order (:674-680)
                const float4 *fp = P.flxout4 + ((size_t)(b & 1) * K + tt) * NM + (size_t)c0 * M + m;
#pragma unroll 1
                for (int c = 0; c < nch; ++c, fp += M) {
                    const float4 f = *fp;
                    if (f.w == 1.0f) { F0 = F0 + f.x; F1 = F1 + f.y; F2 = F2 + f.z; }
                    else if (f.w == 0.0f) { F0 = 0.0f; F1 = 0.0f; F2 = 0.0f; }
                }
            }
            if (SEPR) {  // the shadows the upstream cells sent this step, in cell order (:605-608, :623-625, :683-686)
                const float4 *cp = P.cvout4 + ((size_t)(b & 1) * K + tt) * NM + (size_t)c0 * M + m;
                const float4 *sp = P.srout4 + ((size_t)(b & 1) * K + tt) * NM + (size_t)c0 * M + m;
#pragma unroll 1
                for (int c = 0; c < nch; ++c, cp += M, sp += M) {
                    const float4 a = *cp, bq = *sp;
                    C1 = C1 + a.x; C2 = C2 + a.y; C3 = C3 + a.z; C4 = C4 + a.w;
                    T1 = T1 + bq.x; T2 = T2 + bq.y; T3 = T3 + bq.z; T4 = T4 + bq.w;
                }
            }
            float4 stg_h = make_float4(0.f, 0.f, 0.f, 0.f);  // (no sed_hillslope unless tank 2 is kinematic, :569-575)
            const float vf2 = fminf(vf1, vs2);
            S1 = S1 + vf1 - vf2;
            const float vf3 = fminf(vf2, vs3);
            S2 = S2 + vf2 - vf3;
            const float vf4 = fminf(vf3, vs4);
            S3 = S3 + vf3 - vf4;
            if (P.retorno_aq > 0.0f) {
                const float Ret_aq = fmaxf(0.0f, S3 - H3);
                S2 = S2 + Ret_aq;
                S3 = S3 - Ret_aq;
            }
            float ret_step = 0.0f;
            if (P.retorno_gr > 0.0f) {
                const float Ret = fmaxf(0.0f, S2 - H2);
                S1 = S1 + Ret;
                S2 = S2 - Ret;
                ret_acc = ret_acc + Ret;
                ret_step = Ret;
            }
            if (SEPR) {  // :530-547: vflux(2:3) already carry the return flow (:510-511), and Ret is applied once more
                const bool rg = P.retorno_gr > 0.0f;
                const float w2 = rg ? vf2 + ret_step : vf2, w3 = rg ? vf3 - ret_step : vf3;
                C0 = C0 + (R - vf1) * Co; T0 = T0 + (R - vf1) * St;
                C1 = C1 + (vf1 - w2) * Co; T1 = T1 + (vf1 - w2) * St;
                C2 = C2 + (w2 - w3) * Co; T2 = T2 + (w2 - w3) * St;
                C3 = C3 + (w3 - vf4) * Co; T3 = T3 + (w3 - vf4) * St;
                if (rg) {
                    C1 = C1 + ret_step * Co; C2 = C2 - ret_step * Co;
                    T1 = T1 + ret_step * St; T2 = T2 - ret_step * St;
                }
            }
            if (SNAP) {  // the dumps that follow the vertical fluxes; vflux(2:3) carry the return flow (:509-510)
                vx1 = vf1; vx4 = vf4;
                vx2 = (P.retorno_gr > 0.0f) ? vf2 + ret_step : vf2;
                vx3 = (P.retorno_gr > 0.0f) ? vf3 - ret_step : vf3;
                if (P.rc) { rc1 = rc1 + vx2 - vx3; rc2 = rc2 + R; }               // :521-524
                if (P.ret_snap) P.ret_snap[(size_t)t * N + cell] = ret_step;     // :820-823 (Retorned is zeroed every step)
                if (P.vfl_snap) {
                    const int sl = __ldg(P.vsnap_slot + t);
                    if (sl >= 0) reinterpret_cast<float4 *>(P.vfl_snap)[(size_t)sl * N + cell] = make_float4(vx1, vx2, vx3, vx4);
                }
            }
            float sal = vf4 + E;
            // hillslope outflows of tanks 2..4 (:553-595)
            float h0, h1, h2;
            float hc[3] = {0.f, 0.f, 0.f}, hs[3] = {0.f, 0.f, 0.f};  // separate_rain: shadow outflows of tanks 2..4
            {
                float Sk[3] = {S1, S2, S3}, hk[3];
                float Ck[3] = {C1, C2, C3}, Tk[3] = {T1, T2, T3};
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    if (!KIN || P.st[k] == 1) {
                        hk[k] = fl[k] * Sk[k];
                    } else {
                        const float sec = calc_speed_dev(Sk[k] * m3, kc[k], ke[k], Lh, kv[k], dt, P.niter);
                        hk[k] = fminf(sec * kv[k] * dt / m3, Sk[k]);
                        if (SED && k == 0)   // the reference calls sed_hillslope from inside the kinematic case only (:569-575)
                            stg_h = make_float4(kv[0], Sk[0], hk[0] * m3 / (dt * P.dxp), 1.0f);
                    }
                    if (SEPR) {  // :580-592, in proportion to the tank before the outflow leaves
                        if (Sk[k] > 0.0f) {
                            hc[k] = hk[k] * Ck[k] / Sk[k];
                            hs[k] = hk[k] * Tk[k] / Sk[k];
                        }
                        Ck[k] = Ck[k] - hc[k];
                        Tk[k] = Tk[k] - hs[k];
                    }
                    Sk[k] = Sk[k] - hk[k];
                }
                if (SEPR) { C1 = Ck[0]; C2 = Ck[1]; C3 = Ck[2]; T1 = Tk[0]; T2 = Tk[1]; T3 = Tk[2]; }
                S1 = Sk[0]; S2 = Sk[1]; S3 = Sk[2];
                h0 = hk[0]; h1 = hk[1]; h2 = hk[2];
            }
            // routing (:599-719)
            float4 o;
            // row offset of step t in the (n_cont, n_reg) record tables: only the outlet and control cells need it
            auto qoff_of = [&]() { return ((size_t)m * P.T + t) * P.n_cont; };
            if (!CHAN) {
                o = make_float4(h0, h1, h2, 0.0f);
                if (outlet) {
                    const float s3 = (h0 + h1) + h2;
                    if (P.q) P.q[qoff_of()] = s3 * m3;
                    sal = sal + s3;
                }
                if (rowq >= 0 && P.q) P.q[qoff_of() + rowq] = 0.0f;
                if (SEPR) {
                    P.cvout4[((size_t)(b & 1) * K + tt) * NM + base] = make_float4(hc[0], hc[1], hc[2], 0.0f);
                    P.srout4[((size_t)(b & 1) * K + tt) * NM + base] = make_float4(hs[0], hs[1], hs[2], 0.0f);
                }
                if (SEPF) P.flxout4[((size_t)(b & 1) * K + tt) * NM + base] = make_float4(0.f, 0.f, 0.f, 2.0f);
            } else {
                S4 = S4 + (h0 + h1) + h2 * (float)(type - 2);
                if (SEPR) {  // :623-628
                    C4 = C4 + (hc[0] + hc[1]) + hc[2] * (float)(type - 2);
                    T4 = T4 + (hs[0] + hs[1]) + hs[2] * (float)(type - 2);
                }
                const float sec_area = calc_speed_dev(S4 * m3, ccoef, cexpo, Ls, vch, dt, P.niter);
                const float h3 = fminf(sec_area * vch * dt / m3, S4);
                if (FLOOD && type == 3 && vch * P.fl_av > P.fl_umbral)   // :728-743 (reads only this step's outflow and speed)
                    flood_dev(P, cell, t, h3 * m3 / dt, vch * P.fl_av);
                if (SEPR) {  // :645-655 (the channel tank before the outflow leaves), :683-686, :713-716, :767-770
                    float c3 = 0.0f, s3 = 0.0f;
                    if (S4 != 0.0f) { c3 = h3 * C4 / S4; s3 = h3 * T4 / S4; }
                    C4 = C4 - c3;
                    T4 = T4 - s3;
                    P.cvout4[((size_t)(b & 1) * K + tt) * NM + base] = make_float4(0.0f, 0.0f, hc[2] * (float)(3 - type), c3);
                    P.srout4[((size_t)(b & 1) * K + tt) * NM + base] = make_float4(0.0f, 0.0f, hs[2] * (float)(3 - type), s3);
                    if (P.qsepr && (outlet || rowq >= 0)) {
                        const size_t k2 = ((size_t)m * P.T + t) * 2 * P.n_cont + (outlet ? 0 : (size_t)rowq);
                        P.qsepr[k2] = c3 * m3 / dt;
                        P.qsepr[k2 + P.n_cont] = s3 * m3 / dt;
                    }
                }
                if (SED) stg[(size_t)(2 * tt + 1) * NM] = make_float4(S4, vch, h3 * m3 / dt, sec_area);   // what sed_channel reads (:1537-1608)
                if (SEPF) {  // :658-665 (the ratio uses the channel tank BEFORE the outflow is taken out)
                    F0 = F0 + h0; F1 = F1 + h1; F2 = F2 + h2;
                    if (S4 > 0.0f) {
                        const float r = h3 / S4;
                        F0 = F0 - F0 * r; F1 = F1 - F1 * r; F2 = F2 - F2 * r;
                    } else {
                        F0 = 0.0f; F1 = 0.0f; F2 = 0.0f;
                    }
                }
                S4 = S4 - h3;
                if (SEPF) {  // :674-680, :697-704, :752-759 (and here AFTER)
                    float4 f = make_float4(0.f, 0.f, 0.f, 0.0f);
                    if (S4 > 0.0f) {
                        const float r = h3 / S4;
                        f = make_float4(F0 * r, F1 * r, F2 * r, 1.0f);
                    }
                    P.flxout4[((size_t)(b & 1) * K + tt) * NM + base] = f;
                    if (P.qsep && (outlet || rowq >= 0)) {
                        const size_t row = outlet ? 0 : (size_t)rowq;
                        const size_t k3 = ((size_t)m * P.T + t) * 3 * P.n_cont + row;
                        P.qsep[k3] = f.x * m3 / dt;
                        P.qsep[k3 + P.n_cont] = f.y * m3 / dt;
                        P.qsep[k3 + 2 * (size_t)P.n_cont] = f.z * m3 / dt;
                    }
                }
                o = make_float4(0.0f, 0.0f, h2 * (float)(3 - type), h3);
                if (outlet) {
                    const size_t qoff = qoff_of();
                    if (P.q) P.q[qoff] = h3 * m3 / dt;
                    sal = sal + h3;
                    if (P.show_speed) {
                        if (P.speed) P.speed[qoff] = vch;
                        if (P.area) P.area[qoff] = sec_area;
                    }
                }
                if (rowq >= 0) {  // records (:749-787)
                    const size_t qoff = qoff_of();
                    if (P.q) P.q[qoff + rowq] = h3 * m3 / dt;
                    if (P.show_speed) {
                        if (P.speed) P.speed[qoff + rowq] = vch;
                        if (P.area) P.area[qoff + rowq] = sec_area;
                    }
                }
            }
            if (rowh >= 0) {
                const size_t k = ((size_t)m * P.T + t) * P.n_conth + rowh;
                if (P.hum) P.hum[k] = S0 + S2;
                if (P.st1o) P.st1o[k] = S0;
                if (P.st3o) P.st3o[k] = S2;
            }
            if (SLIDES && sl_risky && sl_slid == 0 && sl_mark > t) {  // slide_ocurrence (:1649-1682)
                const float zs = __ldg(P.sl_zs + cell);
                const float Zw = (S2 * zs) / H2;
                bool fail = false;
                if (Zw >= __ldg(P.sl_zcrit + cell)) {
                    fail = true;
                } else {
                    const float gs = __ldg(P.sl_gammas + cell);
                    const float cr = __ldg(P.sl_cosr + cell), sr = __ldg(P.sl_sinr + cell);
                    const float Num = __ldg(P.sl_cohesion + cell) + (gs * zs - Zw * P.sl_gammaw) * (cr * cr) * __ldg(P.sl_tanfa + cell);
                    const float Den = gs * zs * sr * cr;
                    if (Num / Den <= P.sl_fs) fail = true;
                }
                if (fail) {   // the cell itself + the cells its walk marks (a static count, k_slide_len)
                    sl_slid = 1;
                    sl_fail_t = t;
                    const int marks = 1 + ((P.sl_gullie == 1 && type <= 2) ? __ldg(P.sl_len + cell) : 0);
                    atomicAdd(P.sl_ncelltime + (size_t)m * P.T + t, marks);
                }
            }
            if (SED) stg[(size_t)(2 * tt) * NM] = stg_h;
            *ob_out = o;
            if (SNAP && P.sto_snap) {  // save_storage (:799-803)
                const int sl = __ldg(P.snap_slot + t);
                if (sl >= 0) {
                    float *d = P.sto_snap + ((size_t)sl * N + cell) * 5;
                    d[0] = S0; d[1] = S1; d[2] = S2; d[3] = S3; d[4] = S4;
                }
            }
            if (SNAP && P.spd_snap)  // save_speed (:815-818); a hillslope cell's channel speed keeps its initial value
                reinterpret_cast<float4 *>(P.spd_snap)[(size_t)t * N + cell] =
                    make_float4(kv[0], kv[1], kv[2], CHAN ? vch : P.hsp[3 * NM + base]);
            // storage change of the cell over the step + what left the basin through it: summed over the cells this
            // is sum(StoOut) - StoAtras + salidas of :846 (the transfers between cells cancel)
            dq = (((S0 - So0) + (S1 - So1)) + ((S2 - So2) + (S3 - So3))) + (S4 - So4) + sal;
        }
        if (P.want_red) s_acc[tt * TB_THREADS + threadIdx.x] = dq;
        if (SNAP && P.want_vmean) {  // mean_vfluxes (:807-809): fp64 sums per step; one atomic per warp when it shares the step
            const unsigned full = 0xffffffffu;
            const int ts = (tt < nt) ? t0 + tt : -1;
            const int tl = __shfl_sync(full, ts, 0);
            const bool same = __all_sync(full, ts == tl);
            const float vx[4] = {vx1, vx2, vx3, vx4};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (same) {
                    const double a = warp_sum_d((double)vx[k]);
                    if ((threadIdx.x & 31) == 0 && tl >= 0 && a != 0.0)
                        atomicAdd(P.red + ((size_t)tl * RQ_COUNT + RQ_VFL1 + k) * RED_SUB + (blockIdx.x & (RED_SUB - 1)), a);
                } else if (ts >= 0 && vx[k] != 0.0f) {
                    atomicAdd(P.red + ((size_t)ts * RQ_COUNT + RQ_VFL1 + k) * RED_SUB + (blockIdx.x & (RED_SUB - 1)), (double)vx[k]);
                }
            }
        }
        ob_in += NM;   // next step's slice of the ring
        ob_out += NM;
    }
    if (active) {
        P.st4[base] = make_float4(S0, S1, S2, S3);
        P.s5[base] = S4;
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (KIN && P.st[k] == 2) P.hsp[k * NM + base] = kv[k];
        if (CHAN) P.hsp[3 * NM + base] = vch;
        if (P.retorned) P.retorned[cell] = ret_acc;
        if (SLIDES) {
            const int own = (sl_fail_t != INT_MAX) ? 1 : 0;
            P.sl_slid[base] = sl_slid;
            if (sl_mt_in != INT_MAX) P.sl_marktime[base] = sl_mark;
            if (sl_cnt_in + own) P.sl_acum[base] += sl_cnt_in + own;     // only this thread writes the cell's counter
            if (P.sl_gullie == 1) {
                const int walks = sl_cnt_in + ((own && type <= 2) ? 1 : 0);
                P.sl_ring[(size_t)(b & 1) * NM + base] = make_int2(min(sl_mt_in, type <= 2 ? sl_fail_t : INT_MAX), walks | (type << 30));
            }
        }
        if (SEPF && CHAN) { P.flx[base] = F0; P.flx[NM + base] = F1; P.flx[2 * NM + base] = F2; }
        if (SEPR) {
            P.shc[base] = C0; P.shc[NM + base] = C1; P.shc[2 * NM + base] = C2; P.shc[3 * NM + base] = C3; P.shc[4 * NM + base] = C4;
            P.shs[base] = T0; P.shs[NM + base] = T1; P.shs[2 * NM + base] = T2; P.shs[3 * NM + base] = T3; P.shs[4 * NM + base] = T4;
        }
        if (SNAP && P.rc) { P.rc[cell] = rc1; P.rc[(size_t)N + cell] = rc2; }
    }
    // ---- per-step balance terms: one fp64 atomic per (warp, step) ----
    if (P.want_red) {
        const unsigned full = 0xffffffffu;
        const unsigned act = __ballot_sync(full, active);
        if (act) {
            const int lane = threadIdx.x & 31;
            const int b0 = __shfl_sync(full, b, __ffs(act) - 1);
            const bool uni = __all_sync(full, !active || b == b0);
            const int sub = blockIdx.x & (RED_SUB - 1);
            __syncwarp();  // the lanes read each other's s_acc slots below
            if (uni) {
                constexpr int G = 32 / K;  // lanes per step
                const int s = lane / G, part = lane % G;
                const float *src = s_acc + s * TB_THREADS + (threadIdx.x & ~31) + part * K;
                double a = 0.0;
#pragma unroll
                for (int i = 0; i < K; ++i) a += (double)src[i];
#pragma unroll
                for (int o = G / 2; o > 0; o >>= 1) a += __shfl_xor_sync(full, a, o);
                const int t = b0 * K + s;
                if (part == 0 && t < P.T && a != 0.0) atomicAdd(P.red + ((size_t)t * RQ_COUNT + RQ_SAL) * RED_SUB + sub, a);
            } else if (active) {
                for (int tt = 0; tt < nt; ++tt) {
                    const double a = (double)s_acc[tt * TB_THREADS + threadIdx.x];
                    if (a != 0.0) atomicAdd(P.red + ((size_t)(t0 + tt) * RQ_COUNT + RQ_SAL) * RED_SUB + sub, a);
                }
            }
        }
    }
}

// Every CTA carries the same mix of roles: its first `cpb` warps take 32-thread tiles of the channel list, the
// others tiles of the hillslope list, so the SMs stay balanced although a channel thread costs ~4x a hillslope one.
template <int K, bool KIN, bool SED, bool SLIDES, bool SNAP, bool SEPF, bool SEPR, bool FLOOD>
__global__ void __launch_bounds__(TB_THREADS, SEPR ? TB_MIN_BLOCKS_SED : (KIN ? TB_MIN_BLOCKS_KIN : TB_MIN_BLOCKS))
k_shia_tb(const __grid_constant__ ShiaP P, const int w, const int clo, const int chi, const int hlo, const int hhi,
          const int cpb, const int n_ctiles, const int n_htiles) {
    __shared__ float s_acc[K * TB_THREADS];
    __shared__ float4 s_kid[TbStage<K>::on ? 2 * K * TB_THREADS : 1];
    __shared__ int32_t s_rain[TbStage<K>::on ? K * TB_THREADS : 1];
    const int wi = threadIdx.x >> 5;
    if (wi < cpb) {
        const int tile = (int)blockIdx.x * cpb + wi;
        if (tile < n_ctiles) tb_role<K, true, KIN, SED, SLIDES, SNAP, SEPF, SEPR, FLOOD>(P, w, clo, chi, tile, s_acc, s_kid, s_rain);
    } else {
        const int tile = (int)blockIdx.x * (TB_THREADS / 32 - cpb) + (wi - cpb);
        if (tile < n_htiles) tb_role<K, false, KIN, SED, SLIDES, SNAP, SEPF, SEPR, FLOOD>(P, w, hlo, hhi, tile, s_acc, s_kid, s_rain);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// sediments of a time block (sed_hillslope :1321-1428, sed_channel :1537-1608) -- the second kernel of every wave
// ---------------------------------------------------------------------------------------------------------------------
// The SED transport reads the hydrology (speeds, storages, discharge of the step) but never writes it, so it does not have
// to live in the hydrology's registers: compiled into k_shia_tb it made a 128-register, 137 KB kernel that ran with a
// quarter of the warp slots and instruction-fetch stalls (profiles/README.md, "C5: where the time goes").  k_shia_tb now
// stages two float4 per cell-step and this kernel -- same waves, same role tiles, same time blocks -- carries the four
// 3-fraction stores in registers for the K steps of the block.  Same arithmetic, statement by statement, as before.
template <int K, bool CHAN>
__device__ __forceinline__ void sed_role(const ShiaP &P, const int w, const int jlo, const int jhi, const int tile) {
    const int N = P.N, M = P.M;
    const long long g = (long long)tile * 32 + (threadIdx.x & 31);
    int j, m;
    if (M == 1) { j = jlo + (int)g; m = 0; }
    else { j = jlo + (int)(g / M); m = (int)(g % M); }
    bool active = j < jhi;
    int cell = 0, b = -1;
    uint32_t tw = 0;
    if (active) {
        cell = __ldg((CHAN ? P.cidx : P.hidx) + j);
        tw = __ldg(P.tw + cell);
        b = w - (P.Lmax - (int)(tw >> 8));
        active = (b >= 0) && (b < P.nb);
    }
    asm volatile("griddepcontrol.launch_dependents;");
    const int t0 = b * K;
    const int nt = active ? min(K, P.T - t0) : 0;
    const size_t NM = (size_t)N * M;
    const size_t base = (size_t)cell * M + m;
    const int type = (tw >> 4) & 3, nch = tw & 15;
    int c0 = 0, rowq = -1;
    float aSo = 0.f, So_ch = 0.f;
    if (active) {
        if (nch) c0 = __ldg(P.cs + cell);
        if (tw & TW_CTRLQ) {
            rowq = 1 + find_row(P.ctrl_cells, P.n_ctrl, cell);
            if (rowq >= P.n_cont) rowq = -1;
        }
        if (P.st[0] == 2) aSo = __ldg(P.sed_aso + cell);
        if (CHAN) So_ch = __ldg(P.stream_slope + cell);
    }
    const bool outlet = active && (cell == N - 1);
    // the hydrology kernel of this wave (staging) and the sediment kernel of the previous wave (tributaries) are complete
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (!active) return;
    SedLocal SL;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        SL.Vs[k] = P.sed[(0 + k) * NM + base];
        SL.Vd[k] = P.sed[(3 + k) * NM + base];
        SL.Vsc[k] = P.sed[(6 + k) * NM + base];
        SL.Vdc[k] = P.sed[(9 + k) * NM + base];
    }
    float vdep = P.voldepo[base], vero = P.volero[base];
    double ero_acc[3] = {0.0, 0.0, 0.0}, dep_acc[3] = {0.0, 0.0, 0.0};
    const float4 *stg = P.sedstg4 + (size_t)(b & 1) * K * 2 * NM + base;
    const float4 *so_in = P.sedout4 + (size_t)(b & 1) * K * 3 * NM + (size_t)c0 * M + m;
    float4 *so_out = P.sedout4 + (size_t)(b & 1) * K * 3 * NM + base;
    // Everything a step reads was written before this kernel started (the staging by this wave's hydrology kernel, the
    // tributaries' sediment by the previous wave), so the inputs of step tt + 1 -- the staged values and what the first
    // tributary sent -- are requested while step tt computes.
    float4 n_sh = make_float4(0.f, 0.f, 0.f, 0.f), n_sc = n_sh, n_f0 = n_sh, n_f1 = n_sh, n_f2 = n_sh;
    if (nt > 0) {
        n_sh = stg[0];
        if (CHAN) n_sc = stg[NM];
        if (nch > 0) { n_f0 = so_in[0]; n_f1 = so_in[NM]; n_f2 = so_in[2 * NM]; }
    }
#pragma unroll 1
    for (int tt = 0; tt < nt; ++tt, stg += 2 * NM, so_in += 3 * NM, so_out += 3 * NM) {
        const int t = t0 + tt;
        const float4 sh = n_sh, sc = n_sc;
        float4 f0 = n_f0, f1 = n_f1, f2 = n_f2;
        if (tt + 1 < nt) {
            n_sh = stg[2 * NM];
            if (CHAN) n_sc = stg[3 * NM];
            if (nch > 0) { n_f0 = so_in[3 * NM]; n_f1 = so_in[4 * NM]; n_f2 = so_in[5 * NM]; }
        }
#pragma unroll
        for (int k = 0; k < 12; ++k) SL.up[k] = 0.f;
        SL.ero_add[0] = SL.ero_add[1] = SL.ero_add[2] = 0.f;
        SL.dep_add[0] = SL.dep_add[1] = SL.dep_add[2] = 0.f;
        SL.volero_add = SL.voldepo_add = 0.f;
        // what the upstream cells sent this step, in ascending cell order (:1373-1377, :1401, :1418, :1601)
        const float4 *so = so_in;
#pragma unroll 1
        for (int c = 0; c < nch; ++c) {
            if (c > 0) { so += M; f0 = so[0]; f1 = so[NM]; f2 = so[2 * NM]; }
            const float u[12] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w, f2.x, f2.y, f2.z, f2.w};
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                SL.Vs[i] = SL.Vs[i] + u[0 + i];
                SL.Vs[i] = SL.Vs[i] + u[3 + i];
                SL.Vs[i] = SL.Vs[i] + u[6 + i];
                SL.Vsc[i] = SL.Vsc[i] + u[9 + i];
            }
        }
        if (sh.w != 0.0f) sed_hillslope_dev(P, SL, aSo, sh.x, sh.y, sh.z, cell, type);
        if (!CHAN) {
            if (P.st[0] == 2) {  // hillslope cell: single VolDEPo / VolERO increment (:1426-1427)
                vdep = vdep + SL.voldepo_add;
                vero = vero + SL.volero_add;
            }
        } else {
            float Vsal[3] = {0.f, 0.f, 0.f};
            float vd2 = 0.f;
            sed_channel_dev(P, SL, sc.x, sc.y, sc.z, So_ch, sc.w, Vsal, &vd2);
            // VolDEPo(celda) is incremented twice in a step for a channel cell (:1427 then :1607)
            vdep = vdep + SL.voldepo_add;
            vdep = vdep + vd2;
            vero = vero + SL.volero_add;
            if (P.qsed) {
                if (outlet)
#pragma unroll
                    for (int i = 0; i < 3; ++i) P.qsed[(((size_t)m * P.T + t) * 3 + i) * P.n_cont] = Vsal[i];
                if (rowq >= 0)
#pragma unroll
                    for (int i = 0; i < 3; ++i) P.qsed[(((size_t)m * P.T + t) * 3 + i) * P.n_cont + rowq] = Vsal[i];
            }
        }
        so_out[0] = make_float4(SL.up[0], SL.up[1], SL.up[2], SL.up[3]);
        so_out[NM] = make_float4(SL.up[4], SL.up[5], SL.up[6], SL.up[7]);
        so_out[2 * NM] = make_float4(SL.up[8], SL.up[9], SL.up[10], SL.up[11]);
#pragma unroll
        for (int i = 0; i < 3; ++i) { ero_acc[i] += (double)SL.ero_add[i]; dep_acc[i] += (double)SL.dep_add[i]; }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        P.sed[(0 + k) * NM + base] = SL.Vs[k];
        P.sed[(3 + k) * NM + base] = SL.Vd[k];
        P.sed[(6 + k) * NM + base] = SL.Vsc[k];
        P.sed[(9 + k) * NM + base] = SL.Vdc[k];
    }
    P.voldepo[base] = vdep;
    P.volero[base] = vero;
    // EROt / DEPt (:1403, :1420): fp64 running totals per (cell, member), summed over the cells at the end
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (ero_acc[i] != 0.0) P.sedacc[(size_t)i * NM + base] += ero_acc[i];
        if (dep_acc[i] != 0.0) P.sedacc[(size_t)(3 + i) * NM + base] += dep_acc[i];
    }
}

template <int K>
__global__ void __launch_bounds__(TB_THREADS, SED_MIN_BLOCKS)
k_sed_tb(const __grid_constant__ ShiaP P, const int w, const int clo, const int chi, const int hlo, const int hhi,
         const int cpb, const int n_ctiles, const int n_htiles) {
    const int wi = threadIdx.x >> 5;
    if (wi < cpb) {
        const int tile = (int)blockIdx.x * cpb + wi;
        if (tile < n_ctiles) sed_role<K, true>(P, w, clo, chi, tile);
    } else {
        const int tile = (int)blockIdx.x * (TB_THREADS / 32 - cpb) + (wi - cpb);
        if (tile < n_htiles) sed_role<K, false>(P, w, hlo, hhi, tile);
    }
}

// sed_factor * So**1.664 (:1339) per cell, once per run
__global__ void k_sed_aso(const float *__restrict__ hill_slope, float sed_factor, int n, float *__restrict__ aso) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) aso[i] = sed_factor * wmf_powf_dd(hill_slope[i], 1.664f);
}

// role lists: is_h[i] = 1 for hillslope cells; hpos = exclusive scan of is_h
__global__ void k_role_flags(const uint32_t *__restrict__ tw, int n, int32_t *__restrict__ is_h) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) is_h[i] = (((tw[i] >> 4) & 3) == 1) ? 1 : 0;
}
__global__ void k_role_lists(const int32_t *__restrict__ is_h, const int32_t *__restrict__ hpos, int n,
                             int32_t *__restrict__ hidx, int32_t *__restrict__ cidx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (is_h[i]) hidx[hpos[i]] = i;
    else cidx[i - hpos[i]] = i;
}
// hfirst[L] = number of hillslope cells below the first cell of level L (list position of the level's first cell)
__global__ void k_role_level_first(const int32_t *__restrict__ level_first, const int32_t *__restrict__ hpos, int nlev,
                                   int32_t *__restrict__ hfirst) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l < nlev) hfirst[l] = hpos[level_first[l]];
}

// StoOut = StoIn | Storage + storage_constant (:271-277), (5,N) -> tanks 1-4 as one float4 per (cell, member) + tank 5
__global__ void k_state_init_tb(const float *__restrict__ src, float add, int n, int M, float4 *__restrict__ st4,
                                float *__restrict__ s5) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) v[k] = src[5 * (size_t)i + k] + add;
    for (int m = 0; m < M; ++m) {
        st4[(size_t)i * M + m] = make_float4(v[0], v[1], v[2], v[3]);
        s5[(size_t)i * M + m] = v[4];
    }
}

__global__ void k_state_final_tb(const ShiaP P, float *__restrict__ stoout, float *__restrict__ hspeed_out) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)P.N * P.M) return;
    const int i = (int)(g / P.M), m = (int)(g % P.M);
    const size_t NM = (size_t)P.N * P.M, base = (size_t)i * P.M + m;
    if (stoout) {
        const float4 s = P.st4[base];
        float *o = stoout + ((size_t)m * P.N + i) * 5;
        o[0] = s.x; o[1] = s.y; o[2] = s.z; o[3] = s.w; o[4] = P.s5[base];
    }
    if (hspeed_out) {
        const float4 hc = reinterpret_cast<const float4 *>(P.h_coef)[i];
        const float hcoef[3] = {hc.x, hc.y, hc.z};
        const float *cal = P.calib + (size_t)m * 11;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float v;
            if (k < 3 && P.st[k] == 1 && !P.hs_state) v = hcoef[k] * cal[4 + k];
            else v = P.hsp[k * NM + base];
            hspeed_out[((size_t)m * P.N + i) * 4 + k] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// plan: levels, child ranges, packed topology words
// ---------------------------------------------------------------------------------------------------
__global__ void k_plan_init(const int32_t *__restrict__ drena, int stride, int n, int32_t *__restrict__ lev,
                            int32_t *__restrict__ jmp, int32_t *__restrict__ nch, int32_t *__restrict__ cmin,
                            int32_t *__restrict__ cmax, int *__restrict__ err) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int d = drena[(size_t)i * stride];
    if (d == 0) {
        if (i != n - 1) *err = 2;  // a second outlet
        lev[i] = 0;
        jmp[i] = -1;
        return;
    }
    const int p = n - d;
    if (d < 0 || p <= i || p >= n) { *err = 1; lev[i] = 0; jmp[i] = -1; return; }
    lev[i] = 1;
    jmp[i] = p;
    atomicAdd(&nch[p], 1);
    atomicMin(&cmin[p], i);
    atomicMax(&cmax[p], i);
}

// pointer doubling: lev[i] = hops to the outlet
__global__ void k_plan_jump(const int32_t *__restrict__ lev_in, const int32_t *__restrict__ jmp_in, int n,
                            int32_t *__restrict__ lev_out, int32_t *__restrict__ jmp_out, int *__restrict__ more) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int j = jmp_in[i];
    int l = lev_in[i], nj = -1;
    if (j >= 0) {
        l += lev_in[j];
        nj = jmp_in[j];
        if (nj >= 0) *more = 1;
    }
    lev_out[i] = l;
    jmp_out[i] = nj;
}

__global__ void k_plan_pack(const int32_t *__restrict__ lev, const int32_t *__restrict__ nch,
                            const int32_t *__restrict__ cmin, const int32_t *__restrict__ cmax,
                            const int32_t *__restrict__ unit_type, const int32_t *__restrict__ control,
                            const int32_t *__restrict__ control_h, int n, uint32_t *__restrict__ tw,
                            int32_t *__restrict__ cs, int32_t *__restrict__ level_first, int32_t *__restrict__ fq,
                            int32_t *__restrict__ fh, int *__restrict__ err) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int l = lev[i];
    if (i + 1 < n && lev[i + 1] > l) *err = 3;             // not level-ordered
    if (l >= (1 << 24)) *err = 4;
    const int c = nch[i];
    if (c > 0 && cmax[i] - cmin[i] + 1 != c) *err = 5;     // children not contiguous
    if (c > 15) *err = 6;
    const int ut = unit_type[i];
    if (ut < 1 || ut > 3) *err = 7;
    if (i == 0 || lev[i - 1] != l) level_first[l] = i;
    const int cq = (control && control[i] != 0) ? 1 : 0, ch = (control_h && control_h[i] != 0) ? 1 : 0;
    fq[i] = cq;
    fh[i] = ch;
    tw[i] = ((uint32_t)l << 8) | (ch ? TW_CTRLH : 0u) | (cq ? TW_CTRLQ : 0u) | ((uint32_t)(ut & 3) << 4) | (uint32_t)(c & 15);
    cs[i] = (c > 0) ? cmin[i] : 0;
}

__global__ void k_compact(const int32_t *__restrict__ flag, const int32_t *__restrict__ pos, int n,
                          int32_t *__restrict__ list) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i]) list[pos[i]] = i;
}

// 64-bit order-independent checksum of the arrays a plan is derived from (drain ids, unit types, control points): the key
// of the plan cache.  Every element is mixed with its position, the mixes are summed.
__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}
__global__ void k_plan_hash(const int32_t *__restrict__ drena, int stride, const int32_t *__restrict__ unit_type,
                            const int32_t *__restrict__ control, const int32_t *__restrict__ control_h, int n,
                            unsigned long long *__restrict__ out) {
    unsigned long long h = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned long long a = (unsigned)drena[(size_t)i * stride], b = (unsigned)unit_type[i];
        const unsigned long long c = control ? (control[i] != 0) : 0, d = control_h ? (control_h[i] != 0) : 0;
        h += mix64(((unsigned long long)i << 32) ^ a) + mix64(((unsigned long long)i << 36) ^ (b << 2) ^ (c << 1) ^ d ^ 0x9e3779b97f4a7c15ull);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, h);
}

// ---------------------------------------------------------------------------------------------------
// init / finalize
// ---------------------------------------------------------------------------------------------------
// StoOut = StoIn | Storage + storage_constant (:271-277), AoS (5,N) -> SoA [5][N][M]; also the initial total
__global__ void k_state_init(const float *__restrict__ src, float add, int n, int M, float *__restrict__ sto,
                             double *__restrict__ sto0) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double s = 0.0;
    if (i < n) {
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const float v = src[5 * (size_t)i + k] + add;
            s += (double)v;
            for (int m = 0; m < M; ++m) sto[((size_t)k * n + i) * M + m] = v;
        }
    }
    s = warp_sum_d(s);
    if ((threadIdx.x & 31) == 0 && s != 0.0) atomicAdd(sto0, s);
}

// hspeed initial state (:285-299): HspeedIn, or h_coef*Calib for linear tanks / 0 for kinematic ones; tank 4 -> 0
__global__ void k_hspeed_init(const float *__restrict__ hspeedin, int n, int M, float *__restrict__ hsp) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)n * M) return;
    const int i = (int)(g / M), m = (int)(g % M);
#pragma unroll
    for (int k = 0; k < 4; ++k) hsp[((size_t)k * n + i) * M + m] = hspeedin ? hspeedin[4 * (size_t)i + k] : 0.0f;
}

__global__ void k_state_final(const ShiaP P, float *__restrict__ stoout, float *__restrict__ hspeed_out) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)P.N * P.M) return;
    const int i = (int)(g / P.M), m = (int)(g % P.M);
    const size_t NM = (size_t)P.N * P.M, base = (size_t)i * P.M + m;
    if (stoout)
#pragma unroll
        for (int k = 0; k < 5; ++k) stoout[((size_t)m * P.N + i) * 5 + k] = P.sto[k * NM + base];
    if (hspeed_out) {
        const float4 hc = reinterpret_cast<const float4 *>(P.h_coef)[i];
        const float hcoef[3] = {hc.x, hc.y, hc.z};
        const float *cal = P.calib + (size_t)m * 11;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float v;
            if (k < 3 && P.st[k] == 1 && !P.hs_state) v = hcoef[k] * cal[4 + k];
            else v = P.hsp[k * NM + base];
            hspeed_out[((size_t)m * P.N + i) * 4 + k] = v;
        }
    }
}

// per-record basin total of the rain field in mm (fp64), used for Mean_Rain (:797) and `entradas` (:469)
// (records no step refers to are skipped: a streamed table never received them; blockIdx.y + r0 = record, so that
// tables with more than 65535 records are covered by several launches)
__global__ void __launch_bounds__(256) k_rain_record_sum(const int32_t *__restrict__ rain, int n, int r0,
                                                         const int32_t *__restrict__ used, double *__restrict__ recsum) {
    const int r = r0 + blockIdx.y;
    if (!used[r]) return;
    const int32_t *rec = rain + (size_t)r * n;
    double s = 0.0;
    // 128-bit loads where the record is 16-byte aligned (n % 4 == 0 or r == 0), scalar tail / fallback otherwise
    const bool vec = ((((size_t)r * n) & 3) == 0) && ((reinterpret_cast<uintptr_t>(rain) & 15) == 0);
    const int n4 = vec ? (n >> 2) : 0;
    const int4 *rec4 = reinterpret_cast<const int4 *>(rec);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
        const int4 v = __ldg(rec4 + i);
        s += (double)((float)v.x / 1000.0f);
        s += (double)((float)v.y / 1000.0f);
        s += (double)((float)v.z / 1000.0f);
        s += (double)((float)v.w / 1000.0f);
    }
    for (int i = 4 * n4 + blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        s += (double)((float)rec[i] / 1000.0f);
    s = warp_sum_d(s);
    if ((threadIdx.x & 31) == 0 && s != 0.0) atomicAdd(recsum + r, s);
}

// Acum_rain (:451): per-cell fp32 running sum over the steps, in time order
__global__ void k_acum_rain(const int32_t *__restrict__ rain, const int32_t *__restrict__ pos, int n, int T,
                            float *__restrict__ acum) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float a = 0.0f;
    for (int t = 0; t < T; ++t) {
        const int p = pos[t];
        if (p != 1) a = a + (float)rain[(size_t)(p - 1) * n + i] / 1000.0f;
    }
    acum[i] = a;
}

__global__ void k_shia_epilogue(const double *__restrict__ red, const double *__restrict__ recsum,
                                const int32_t *__restrict__ pos, const double *__restrict__ sto0, int n, int T,
                                int show_storage, int bal_direct, float *__restrict__ balance,
                                float *__restrict__ mean_rain,
                                float *__restrict__ mean_storage, float *__restrict__ mean_speed,
                                float *__restrict__ mean_retorno, float *__restrict__ mean_vfluxes) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    auto q = [&](int tt, int k) {
        double s = 0.0;
        for (int j = 0; j < RED_SUB; ++j) s += red[((size_t)tt * RQ_COUNT + k) * RED_SUB + j];
        return s;
    };
    auto total = [&](int tt) {
        if (tt < 0) return *sto0;
        if (!show_storage) return q(tt, RQ_STO1);
        double s = 0.0;
        for (int k = 0; k < 5; ++k) s += q(tt, RQ_STO1 + k);
        return s;
    };
    const double entradas = recsum[pos[t] - 1];
    if (balance) {
        if (bal_direct) balance[t] = (float)(q(t, RQ_SAL) - entradas);  // k_shia_tb sums dS + outputs per cell
        else balance[t] = (float)(total(t) - total(t - 1) - entradas + q(t, RQ_SAL));
    }
    if (mean_rain) mean_rain[t] = (float)(entradas / (double)n);
    if (mean_storage)
        for (int k = 0; k < 5; ++k) mean_storage[5 * (size_t)t + k] = (float)(q(t, RQ_STO1 + k) / (double)n);
    if (mean_speed)
        for (int k = 0; k < 4; ++k) mean_speed[4 * (size_t)t + k] = (float)(q(t, RQ_SPD1 + k) / (double)n);
    if (mean_retorno) mean_retorno[t] = (float)(q(t, RQ_RET) / (double)n);
    if (mean_vfluxes)
        for (int k = 0; k < 4; ++k) mean_vfluxes[4 * (size_t)t + k] = (float)(q(t, RQ_VFL1 + k) / (double)n);
}

// floods epilogue: flood_flood from the last-writer keys; the cell of the very last evaluation (largest (step, cell))
__global__ void k_flood_final(const unsigned long long *__restrict__ key, const int32_t *__restrict__ last, int N,
                              int32_t *__restrict__ flood, unsigned long long *__restrict__ best) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    if (flood) flood[i] = (int32_t)(key[i] & 3ull);
    if (last[i] >= 0) atomicMax(best, ((unsigned long long)last[i] << 32) | (unsigned)i);
}
__global__ void k_flood_scalars(const unsigned long long *__restrict__ best, const float *__restrict__ o, const int32_t *__restrict__ last,
                                int N, float *__restrict__ sc) {
    if (threadIdx.x || blockIdx.x) return;
    sc[0] = sc[1] = sc[2] = 0.0f;
    const int c = (int)(*best & 0xffffffffull);
    if (last[c] >= 0) {                    // best == 0 is also what "no evaluation at all" leaves behind
        sc[0] = o[6 * (size_t)N + c];      // flood_rdf
        sc[2] = o[7 * (size_t)N + c];      // flood_diff (flood_area is never assigned by the Fortran)
    }
}

// Storage_conv / Storage_stra: SoA [5][N] state -> the (5,N) f2py layout
__global__ void k_shadow_final(const float *__restrict__ sh, int N, float *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
#pragma unroll
    for (int k = 0; k < 5; ++k) out[5 * (size_t)i + k] = sh[(size_t)k * N + i];
}

// save_rc epilogue (:863-867): rc(1) = min(rc(1), rc(2)), SoA accumulators -> the (2,N) f2py layout
__global__ void k_rc_final(const float *__restrict__ rc, int N, float *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float a = rc[i], b = rc[(size_t)N + i];
    reinterpret_cast<float2 *>(out)[i] = make_float2(a > b ? b : a, b);
}

// slide_allocate, modelosv2.f90:1613-1648
// The fp32 trig of the reference (libm, < 1 ulp) is reproduced by evaluating in fp64 and rounding once: the soil
// thresholds subtract nearly equal tangents, so a 2-4 ulp tanf would be amplified into visible differences.
__device__ __forceinline__ float tan_cr(float x) { return (float)tan((double)x); }
__device__ __forceinline__ float cos_cr(float x) { return (float)cos((double)x); }
__device__ __forceinline__ float sin_cr(float x) { return (float)sin((double)x); }
__device__ __forceinline__ float atan_cr(float x) { return (float)atan((double)x); }

__global__ void k_slide_allocate(const ShiaP P, float *__restrict__ zmin, float *__restrict__ zcrit,
                                 float *__restrict__ zmax, float *__restrict__ bo, float *__restrict__ risk,
                                 float *__restrict__ cosr, float *__restrict__ sinr, float *__restrict__ tanfa) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.N) return;
    const float gw = P.sl_gammaw;
    const float gs = P.sl_gammas[i], co = P.sl_cohesion[i], fa = P.sl_friction[i], rs = P.sl_radslope[i], zs = P.sl_zs[i];
    const float cr = cos_cr(rs), tfa = tan_cr(fa), trs = tan_cr(rs);
    cosr[i] = cr; sinr[i] = sin_cr(rs); tanfa[i] = tfa;
    const float c2 = cr * cr;
    const float zmn = co / ((gw * c2 * tfa) + (gs * c2 * (trs - tfa)));
    const float zcr = (gs / gw) * zs * (1.0f - (trs / tfa)) + (co / (gw * c2 * tfa));
    const float zmx = co / ((gs * c2) * (trs - tfa));
    const float b = atan_cr(-tan_cr(fa * (gw - gs) / gs));
    float rv = 2.0f;
    if (((P.tw[i] >> 4) & 3) == 3) rv = 0.0f;
    if (rs < b) rv = 0.0f;
    if (zs < zmn && rs < b) rv = 0.0f;
    if (zs > zmx && rs > b) rv = 1.0f;
    zmin[i] = zmn; zcrit[i] = zcr; zmax[i] = zmx; bo[i] = b; risk[i] = rv;
}

// outputs carry the member axis last in f2py terms: (1, N, M) -> [m][i]
__global__ void k_slide_final(const ShiaP P, float *__restrict__ occ, int32_t *__restrict__ acum_out) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)P.N * P.M) return;
    const int i = (int)(g / P.M), m = (int)(g % P.M);
    const size_t o = (size_t)m * P.N + i;
    if (occ) occ[o] = (P.sl_marktime[g] != INT_MAX) ? 3.0f : (P.sl_slid[g] ? 1.0f : 0.0f);
    if (acum_out) acum_out[o] = P.sl_acum[g];
}
// cells the marking walk from a cell visits (slide_hill2gullie :1683-1707): downstream while the unit type does not grow
__global__ void k_slide_len(const int32_t *__restrict__ drena, int stride, const uint32_t *__restrict__ tw, int n,
                            int32_t *__restrict__ len) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int local = i, ltype = (__ldg(tw + i) >> 4) & 3, cnt = 0;
    if (ltype <= 2) {
        for (;;) {
            const int dd = drena[(size_t)local * stride];
            if (dd == 0) break;
            const int dir = n - dd;
            const int dtype = (__ldg(tw + dir) >> 4) & 3;
            if (dtype > ltype) break;
            ++cnt;
            local = dir;
            ltype = dtype;
        }
    }
    len[i] = cnt;
}

__global__ void k_slide_state_init(const ShiaP P) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)P.N * P.M) return;
    P.sl_slid[g] = 0;
    P.sl_marktime[g] = INT_MAX;
    P.sl_acum[g] = 0;
}
// sum of sedacc over the cells: one block per (member, quantity); EROt / DEPt as (3, M) -> [m][3]
__global__ void __launch_bounds__(256) k_sed_totals(const double *__restrict__ sedacc, int n, int M, float *__restrict__ erot,
                                                    float *__restrict__ dept) {
    const int m = blockIdx.x, q = blockIdx.y;
    const size_t NM = (size_t)n * M;
    double a = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) a += sedacc[(size_t)q * NM + (size_t)i * M + m];
    a = warp_sum_d(a);
    __shared__ double sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < 8; ++k) t += sm[k];
        if (q < 3) { if (erot) erot[3 * m + q] = (float)t; }
        else if (dept) dept[3 * m + (q - 3)] = (float)t;
    }
}
// [N][M] -> [m][i]
__global__ void k_member_major_f32(const float *__restrict__ src, int n, int M, float *__restrict__ dst) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)n * M) return;
    const int i = (int)(g / M), m = (int)(g % M);
    dst[(size_t)m * n + i] = src[g];
}

__global__ void k_sed_init(const float *__restrict__ vd_init, int n, int M, float *__restrict__ sed) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)n * M) return;
    const int i = (int)(g / M), m = (int)(g % M);
    const size_t NM = (size_t)n * M, base = (size_t)i * M + m;
    for (int k = 0; k < 12; ++k) sed[k * NM + base] = 0.0f;
    if (vd_init)
        for (int k = 0; k < 3; ++k) sed[(3 + k) * NM + base] = vd_init[3 * (size_t)i + k];
}

// SoA [12][N] -> the four (3,N) arrays
// (3, N, M) outputs -> [m][i][k]
__global__ void k_sed_final(const float *__restrict__ sed, int n, int M, float *vs, float *vd, float *vsc, float *vdc) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)n * M) return;
    const int i = (int)(g / M), m = (int)(g % M);
    const size_t NM = (size_t)n * M, o = 3 * ((size_t)m * n + i);
    for (int k = 0; k < 3; ++k) {
        if (vs) vs[o + k] = sed[(size_t)(0 + k) * NM + g];
        if (vd) vd[o + k] = sed[(size_t)(3 + k) * NM + g];
        if (vsc) vsc[o + k] = sed[(size_t)(6 + k) * NM + g];
        if (vdc) vdc[o + k] = sed[(size_t)(9 + k) * NM + g];
    }
}

__global__ void k_d2f(const double *__restrict__ a, int n, float *__restrict__ b) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b[i] = (float)a[i];
}

// wmf.py:749-754 (__eval_nash__) and :775-781 (__eval_q_pico__), one block per member, fp64
__global__ void __launch_bounds__(256) k_eval_scores(const float *__restrict__ q, int n_cont, int T, int row,
                                                     const float *__restrict__ qobs, float *__restrict__ scores) {
    const int m = blockIdx.x;
    const float *qs = q + (size_t)m * T * n_cont + row;
    __shared__ double sh[3][8];
    __shared__ double s_mean;
    auto block_sum = [&](double v, int slot) {
        v = warp_sum_d(v);
        if ((threadIdx.x & 31) == 0) sh[slot][threadIdx.x >> 5] = v;
        __syncthreads();
        double r = 0.0;
        for (int k = 0; k < 8; ++k) r += sh[slot][k];
        __syncthreads();
        return r;
    };
    double so = 0.0, cnt = 0.0;
    for (int t = threadIdx.x; t < T; t += 256) {
        const float o = qobs[t];
        if (isfinite(o)) { so += (double)o; cnt += 1.0; }
        else if (!isnan(o)) { so += (double)o; }  // nansum keeps infinities
    }
    const double tot = block_sum(so, 0), n_ok = block_sum(cnt, 1);
    if (threadIdx.x == 0) s_mean = tot / n_ok;
    __syncthreads();
    const double med = s_mean;
    double num = 0.0, den = 0.0;
    // argmax over the non-NaN entries, first occurrence wins (np.argmax on a masked array)
    double bo = -INFINITY, bs = -INFINITY;
    int io = T, is = T;
    for (int t = threadIdx.x; t < T; t += 256) {
        const double o = (double)qobs[t], s = (double)qs[(size_t)t * n_cont];
        const double d1 = (o - s) * (o - s), d2 = (o - med) * (o - med);
        if (!isnan(d1)) num += d1;
        if (!isnan(d2)) den += d2;
        if (!isnan(o) && (o > bo)) { bo = o; io = t; }
        if (!isnan(s) && (s > bs)) { bs = s; is = t; }
    }
    num = block_sum(num, 0);
    den = block_sum(den, 1);
    __shared__ double mo[256], ms[256];
    __shared__ int ko[256], ks[256];
    mo[threadIdx.x] = bo; ko[threadIdx.x] = io; ms[threadIdx.x] = bs; ks[threadIdx.x] = is;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 256; ++k) {
            if (mo[k] > bo || (mo[k] == bo && ko[k] < io)) { bo = mo[k]; io = ko[k]; }
            if (ms[k] > bs || (ms[k] == bs && ks[k] < is)) { bs = ms[k]; is = ks[k]; }
        }
        scores[2 * (size_t)m + 0] = (float)(1.0 - num / den);
        scores[2 * (size_t)m + 1] = (float)(((bo - bs) / bo) * 100.0);
    }
}

__global__ void k_calc_speed(float sm, float coef, float expo, float elem_long, float dt, int niter, float *io) {
    float v = io[0];
    io[1] = calc_speed_dev(sm, coef, expo, elem_long, v, dt, niter);
    io[0] = v;
}

// ---------------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------------
static int read_flag(wmf_ctx *ctx, const int *d, int *h) {
    CU_TRY(ctx, cudaMemcpyAsync(h, d, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return WMF_OK;
}

// Host -> device streaming of the rain table next to the running wavefront.
// The skewed sweep reads the table along anti-diagonals: cell (level L) needs the record of step t at wave
// t/K + (Lmax - L), so the far-upstream cells want the last records long before the outlet wants the first ones.
// The table [record][cell] is therefore cut into tiles (record group x level-contiguous cell group), the tiles are
// copied in order of first use on the context's copy stream (2-D copies from pinned memory), and the compute stream
// waits, wave by wave, only for the tiles that wave can touch.  Records no step refers to are never copied.
// Pageable host memory cannot be copied asynchronously, and device memory needs no copy: both take the plain path.
struct RainStream {
    struct Tile { int w_first, r0, r1, c0, c1; };
    wmf_ctx *owner = nullptr;
    std::vector<Tile> tiles;
    const int32_t *host = nullptr;
    int32_t *dev = nullptr;
    int n_records = 0, N = 0, waited = 0, rg = 1, n_early = 0;
    std::vector<int> tmin;       // first step that reads each record
    std::vector<char> rg_done;   // record groups already on their way (early())
    bool active = false;
    ~RainStream() {  // an early error return must not free the buffer under a copy still in flight
        if (active && waited < (int)tiles.size() && owner && owner->copy_stream) cudaStreamSynchronize(owner->copy_stream);
    }
    int begin(wmf_ctx *ctx, Scope &sc, const int32_t *rain, int n_rec, int n, const int32_t **out) {
        cudaPointerAttributes a;
        bool pinned = false;
        if (cudaPointerGetAttributes(&a, rain) == cudaSuccess) pinned = (a.type == cudaMemoryTypeHost);
        else cudaGetLastError();
        const size_t total = (size_t)n_rec * n;
        size_t min_bytes = 32u << 20;  // below this a single copy costs less than the tile bookkeeping
        if (const char *e = getenv("WMF_RAIN_STREAM_MIN_MB")) min_bytes = (size_t)atol(e) << 20;
        if (!pinned || total * 4 < min_bytes) return sc.in(rain, total, out);
        WMF_TRY(sc.alloc(&dev, total));
        if (!ctx->copy_stream) CU_TRY(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        host = rain; n_records = n_rec; N = n; owner = ctx; active = true;
        *out = dev;
        return WMF_OK;
    }
    static void tile_grid(int *n_rg, int *n_cg) {
        *n_rg = 12; *n_cg = 12;  // measured best on C2 (tools/e2e_probe.py): 12x12 47.4 ms, 24x24 48.7, by record only 68.9
        if (const char *e = getenv("WMF_RAIN_TILES")) sscanf(e, "%d,%d", n_rg, n_cg);
        *n_rg = std::max(1, *n_rg); *n_cg = std::max(1, *n_cg);
    }
    // Before anything else of the run is queued: the record groups the sweep reads first go out whole (all cells), so the
    // link is busy while the other inputs are staged and the plan is built.  pos[t] = 1-based record of step t.
    int early(wmf_ctx *ctx, const std::vector<int32_t> &pos, int T) {   // pos: one or several sequences of T steps
        if (!active) return WMF_OK;
        tmin.assign((size_t)n_records, INT_MAX);
        for (size_t k = 0; k < pos.size(); ++k) tmin[pos[k] - 1] = std::min(tmin[pos[k] - 1], (int)(k % T));
        int n_rg, n_cg;
        tile_grid(&n_rg, &n_cg);
        rg = std::max(1, (n_records + n_rg - 1) / n_rg);
        const int groups = (n_records + rg - 1) / rg;
        rg_done.assign((size_t)groups, 0);
        // the allocation is ordered on the compute stream: the copy stream must not start before it
        CU_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        CU_TRY(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev1, 0));
        for (int k = 0; k < 2 && k < groups; ++k) {  // the two groups with the earliest first use
            int best = -1, best_t = INT_MAX;
            for (int g = 0; g < groups; ++g) {
                if (rg_done[g]) continue;
                int tm = INT_MAX;
                for (int r = g * rg; r < std::min(n_records, (g + 1) * rg); ++r) tm = std::min(tm, tmin[r]);
                if (tm < best_t) { best_t = tm; best = g; }
            }
            if (best < 0) break;
            rg_done[best] = 1;
            const int r0 = best * rg, r1 = std::min(n_records, r0 + rg);
            tiles.push_back({-1, r0, r1, 0, N});
            if (ctx->copy_events.size() < tiles.size()) {
                cudaEvent_t e;
                CU_TRY(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                ctx->copy_events.push_back(e);
            }
            CU_TRY(ctx, cudaMemcpyAsync(dev + (size_t)r0 * N, host + (size_t)r0 * N, (size_t)(r1 - r0) * N * 4, cudaMemcpyHostToDevice,
                                        ctx->copy_stream));
            CU_TRY(ctx, cudaEventRecord(ctx->copy_events[tiles.size() - 1], ctx->copy_stream));
        }
        n_early = (int)tiles.size();
        return WMF_OK;
    }
    // h_first[L] = first cell of level L (cells are ordered by decreasing level)
    int schedule(wmf_ctx *ctx, const std::vector<int32_t> &h_first, int Lmax, const std::vector<int32_t> &pos, int K) {
        if (!active) return WMF_OK;
        (void)pos;
        int n_rg, n_cg;
        tile_grid(&n_rg, &n_cg);
        std::vector<int> cb{0};
        {
            const int target = std::max(1, N / n_cg);
            for (int L = Lmax - 1; L >= 0; --L)
                if (h_first[L] - cb.back() >= target) cb.push_back(h_first[L]);
            if (cb.back() != N) cb.push_back(N);
        }
        // level of the first (deepest) cell of a group: levels decrease with the cell index
        auto level_of_cell = [&](int cell) {
            int lo = 0, hi = Lmax;  // largest L with h_first[L] <= cell  (h_first is decreasing in L)
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (h_first[mid] <= cell) hi = mid; else lo = mid + 1;
            }
            return lo;
        };
        for (int r0 = 0; r0 < n_records; r0 += rg) {
            if (rg_done[r0 / rg]) continue;  // went out whole in early()
            const int r1 = std::min(n_records, r0 + rg);
            int tm = INT_MAX;
            for (int r = r0; r < r1; ++r) tm = std::min(tm, tmin[r]);
            if (tm == INT_MAX) continue;  // nobody reads these records
            for (size_t g = 0; g + 1 < cb.size(); ++g)
                tiles.push_back({tm / K + (Lmax - level_of_cell(cb[g])), r0, r1, cb[g], cb[g + 1]});
        }
        std::stable_sort(tiles.begin() + n_early, tiles.end(), [](const Tile &x, const Tile &y) { return x.w_first < y.w_first; });
        while (ctx->copy_events.size() < tiles.size()) {
            cudaEvent_t e;
            CU_TRY(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ctx->copy_events.push_back(e);
        }
        for (size_t i = (size_t)n_early; i < tiles.size(); ++i) {
            const Tile &t = tiles[i];
            const size_t o = (size_t)t.r0 * N + t.c0;
            CU_TRY(ctx, cudaMemcpy2DAsync(dev + o, (size_t)N * 4, host + o, (size_t)N * 4, (size_t)(t.c1 - t.c0) * 4,
                                          (size_t)(t.r1 - t.r0), cudaMemcpyHostToDevice, ctx->copy_stream));
            CU_TRY(ctx, cudaEventRecord(ctx->copy_events[i], ctx->copy_stream));
        }
        return WMF_OK;
    }
    // make the compute stream wait for every tile wave `w` (or an earlier one) can read
    int wait_wave(wmf_ctx *ctx, int w) {
        if (!active) return WMF_OK;
        while (waited < (int)tiles.size() && tiles[waited].w_first <= w)
            CU_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->copy_events[waited++], 0));
        return WMF_OK;
    }
    int wait_all(wmf_ctx *ctx) { return wait_wave(ctx, INT_MAX); }
};

template <bool SED, bool SLIDES>
static void launch_wave(wmf_ctx *ctx, const ShiaP &P, int w, int lo, int hi) {
    const long long work = (long long)(hi - lo) * P.M;
    const int grid = (int)((work + WAVE_THREADS - 1) / WAVE_THREADS);
    k_shia_wave<SED, SLIDES><<<grid, WAVE_THREADS, 0, ctx->stream>>>(P, w, lo, hi);
    ctx->launches++;
}

template <int K, bool SED = false, bool SLIDES = false, bool SNAP = false, bool SEPF = false, bool SEPR = false, bool FLOOD = false>
static void launch_tb(wmf_ctx *ctx, const ShiaP &P, int w, int clo, int chi, int hlo, int hhi, bool pdl, bool pdl_sed = true) {
    constexpr int WPB = TB_THREADS / 32;
    const int wc = (int)(((long long)(chi - clo) * P.M + 31) / 32), wh = (int)(((long long)(hhi - hlo) * P.M + 31) / 32);
    if (wc + wh == 0) return;
    int cpb = 0, nblk = 0;
    if (wc == 0) { cpb = 0; nblk = (wh + WPB - 1) / WPB; }
    else if (wh == 0) { cpb = WPB; nblk = (wc + WPB - 1) / WPB; }
    else {
        for (int c = 1; c < WPB; ++c) {
            const int nb = std::max((wc + c - 1) / c, (wh + (WPB - c) - 1) / (WPB - c));
            if (nblk == 0 || nb < nblk) { nblk = nb; cpb = c; }
        }
    }
    // programmatic dependent launch: the CTAs of this wave may start while the previous wave drains; they run their
    // constant-only prologue and block in griddepcontrol.wait until the previous wave has completed
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)nblk);
    cfg.blockDim = dim3(TB_THREADS);
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    // Two streams (ensembles): the sediment kernel of wave w runs next to the hydrology kernel of wave w + 1 -- one is bound
    // by memory latency, the other by instruction issue.  The staging and sediment rings are two time blocks deep, so the
    // hydrology of wave w + 2 waits for the sediments of wave w.
    const bool two = SED && ctx->sed_two_streams;
    if (two && ctx->sed_seq >= 2) cudaStreamWaitEvent(ctx->stream, ctx->sed_ev_s[(ctx->sed_seq - 2) & 3], 0);
    const bool kin = P.st[0] == 2 || P.st[1] == 2 || P.st[2] == 2;
    if (kin) cudaLaunchKernelEx(&cfg, k_shia_tb<K, true, SED, SLIDES, SNAP, SEPF, SEPR, FLOOD>, P, w, clo, chi, hlo, hhi, cpb, wc, wh);
    else cudaLaunchKernelEx(&cfg, k_shia_tb<K, false, SED, SLIDES, SNAP, SEPF, SEPR, FLOOD>, P, w, clo, chi, hlo, hhi, cpb, wc, wh);
    ctx->launches++;
    if constexpr (SED) {  // the sediments of the same wave: same tiles, same time blocks, right behind the hydrology
        cfg.numAttrs = pdl_sed ? 1 : 0;
        if (two) {
            const int k = (int)(ctx->sed_seq & 3);
            cudaEventRecord(ctx->sed_ev_h[k], ctx->stream);
            cudaStreamWaitEvent(ctx->sed_stream, ctx->sed_ev_h[k], 0);
            cfg.stream = ctx->sed_stream;
            cfg.numAttrs = 0;
        }
        cudaLaunchKernelEx(&cfg, k_sed_tb<K>, P, w, clo, chi, hlo, hhi, cpb, wc, wh);
        ctx->launches++;
        if (two) cudaEventRecord(ctx->sed_ev_s[ctx->sed_seq & 3], ctx->sed_stream);
        ctx->sed_seq++;
    }
}

// Steps per time block of k_shia_tb; 0 = use the one-step-per-launch kernel k_shia_wave.
// A wave holds N * M * min(1, (T/K) / Lmax) threads and the sweep's critical path grows with K * Lmax, so K is the
// largest of {8, 4, 2, 1} that still fills about two thirds of the GPU's resident thread slots per wave (measured on the
// C2 basin, Lmax = 1022: T = 1000 -> K = 4 (29 ms; K = 2: 35, K = 8: 34), T = 400 -> K = 2, T = 200 -> K = 1..2).
static int tb_steps(int T, bool eligible, long long cells_x_members, int Lmax, int num_sms, bool single) {
    if (!eligible) return 0;
    if (const char *e = getenv("WMF_SHIA_K")) {
        const int k = atoi(e);
        if (k <= 0) return 0;
        int p = 1;
        while (p * 2 <= k && p * 2 <= 32) p *= 2;
        while (p > 1 && p > T) p /= 2;
        return p;
    }
    const double want = 0.66 * 2048.0 * num_sms;
    // ensembles have threads to spare: they take the longest block (measured, 260 100 cells x 512 members x 200 steps:
    // K = 4 58 G, 8 68 G, 16 72 G member-cell-steps/s); wmf_shia_run halves it again if the outflow ring would not fit
    for (int k = single ? 8 : 16; k > 1; k /= 2) {
        if (k > T) continue;
        const double in_flight = std::min(1.0, ((double)T / k) / std::max(1, Lmax));
        if ((double)cells_x_members * in_flight >= want) return k;
    }
    return 1;
}

// The plan of a sweep: hop level of every cell (pointer doubling), child ranges, packed topology words, the sorted lists
// of control cells and the hillslope / channel role lists.  Depends on the drain ids, unit types and control points only;
// wmf_shia_run keeps it in the context and rebuilds it when their checksum changes.
static int shia_build_plan(wmf_ctx *ctx, Scope &sc, const int32_t *drena, int stride, int N, const int32_t *d_unit,
                           const int32_t *d_control, const int32_t *d_controlh, wmf_shia_plan &pl) {
    int32_t *lev_a, *lev_b, *jmp_a, *jmp_b, *nch, *cmin, *cmax, *cs, *level_first, *fq, *fh, *posq, *posh;
    uint32_t *tw;
    int *d_err;
    WMF_TRY(sc.alloc(&lev_a, (size_t)N)); WMF_TRY(sc.alloc(&lev_b, (size_t)N));
    WMF_TRY(sc.alloc(&jmp_a, (size_t)N)); WMF_TRY(sc.alloc(&jmp_b, (size_t)N));
    WMF_TRY(sc.alloc(&nch, (size_t)N)); WMF_TRY(sc.alloc(&cmin, (size_t)N)); WMF_TRY(sc.alloc(&cmax, (size_t)N));
    WMF_TRY(sc.alloc(&cs, (size_t)N)); WMF_TRY(sc.alloc(&tw, (size_t)N));
    WMF_TRY(sc.alloc(&fq, (size_t)N)); WMF_TRY(sc.alloc(&fh, (size_t)N));
    WMF_TRY(sc.alloc(&posq, (size_t)N)); WMF_TRY(sc.alloc(&posh, (size_t)N));
    WMF_TRY(sc.alloc(&d_err, 2));
    CU_TRY(ctx, cudaMemsetAsync(d_err, 0, 2 * sizeof(int), ctx->stream));
    CU_TRY(ctx, cudaMemsetAsync(nch, 0, sizeof(int32_t) * (size_t)N, ctx->stream));
    CU_TRY(ctx, cudaMemsetAsync(cmin, 0x7f, sizeof(int32_t) * (size_t)N, ctx->stream));
    CU_TRY(ctx, cudaMemsetAsync(cmax, 0, sizeof(int32_t) * (size_t)N, ctx->stream));
    const int B = 256, G = grid_for(N, B);
    LAUNCH(ctx, k_plan_init, G, B, 0, drena, stride, N, lev_a, jmp_a, nch, cmin, cmax, d_err);
    int herr = 0;
    WMF_TRY(read_flag(ctx, d_err, &herr));
    if (herr == 1) WMF_FAIL(ctx, WMF_ERR_TOPOLOGY, "shia: drena is not upstream-first (need 0 <= id < n and parent index > cell index)");
    if (herr == 2) WMF_FAIL(ctx, WMF_ERR_TOPOLOGY, "shia: drena has an outlet (id 0) that is not the last cell");
    {
        int32_t dlast = 1;
        CU_TRY(ctx, cudaMemcpyAsync(&dlast, drena + (size_t)(N - 1) * stride, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        if (dlast != 0) WMF_FAIL(ctx, WMF_ERR_TOPOLOGY, "shia: the last cell must be the outlet (drena == 0)");
    }
    for (int round = 0; round < 40; ++round) {
        CU_TRY(ctx, cudaMemsetAsync(d_err + 1, 0, sizeof(int), ctx->stream));
        LAUNCH(ctx, k_plan_jump, G, B, 0, lev_a, jmp_a, N, lev_b, jmp_b, d_err + 1);
        std::swap(lev_a, lev_b);
        std::swap(jmp_a, jmp_b);
        int more = 0;
        WMF_TRY(read_flag(ctx, d_err + 1, &more));
        if (!more) break;
    }
    int32_t Lmax = 0;
    CU_TRY(ctx, cudaMemcpyAsync(&Lmax, lev_a, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (Lmax < 0 || Lmax >= (1 << 24)) WMF_FAIL(ctx, WMF_ERR_TOPOLOGY, "shia: tree depth %d unsupported", Lmax);
    WMF_TRY(sc.alloc(&level_first, (size_t)Lmax + 2));
    CU_TRY(ctx, cudaMemsetAsync(level_first, 0xff, sizeof(int32_t) * ((size_t)Lmax + 2), ctx->stream));
    LAUNCH(ctx, k_plan_pack, G, B, 0, lev_a, nch, cmin, cmax, d_unit, d_control, d_controlh, N, tw, cs, level_first, fq, fh, d_err);
    WMF_TRY(read_flag(ctx, d_err, &herr));
    if (herr == 3 || herr == 5)
        WMF_FAIL(ctx, WMF_ERR_TOPOLOGY,
                 "shia: cells are not in basin_find order (levels must be non-increasing and the cells draining into one cell "
                 "contiguous); pass the structure produced by basin_find/basin_cut");
    if (herr == 4 || herr == 6) WMF_FAIL(ctx, WMF_ERR_TOPOLOGY, "shia: more than 15 cells drain into one cell, or depth >= 2^24");
    if (herr == 7) WMF_FAIL(ctx, WMF_ERR_ARG, "shia: unit_type must be 1, 2 or 3");
    std::vector<int32_t> h_first((size_t)Lmax + 2);
    CU_TRY(ctx, cudaMemcpyAsync(h_first.data(), level_first, sizeof(int32_t) * ((size_t)Lmax + 1), cudaMemcpyDeviceToHost, ctx->stream));
    int64_t nq = 0, nh = 0;
    WMF_TRY(wmf_device_exclusive_scan_i32(ctx, sc, fq, posq, N, &nq));
    WMF_TRY(wmf_device_exclusive_scan_i32(ctx, sc, fh, posh, N, &nh));
    int32_t *ctrl_cells, *ctrlh_cells;
    WMF_TRY(sc.alloc(&ctrl_cells, (size_t)nq + 1));
    WMF_TRY(sc.alloc(&ctrlh_cells, (size_t)nh + 1));
    LAUNCH(ctx, k_compact, G, B, 0, fq, posq, N, ctrl_cells);
    LAUNCH(ctx, k_compact, G, B, 0, fh, posh, N, ctrlh_cells);
    // role lists of the time-blocked kernel
    int32_t *is_h, *hpos, *d_hfirst;
    int64_t n_hill = 0;
    WMF_TRY(sc.alloc(&is_h, (size_t)N)); WMF_TRY(sc.alloc(&hpos, (size_t)N));
    WMF_TRY(sc.alloc(&d_hfirst, (size_t)Lmax + 2));
    LAUNCH(ctx, k_role_flags, G, B, 0, tw, N, is_h);
    WMF_TRY(wmf_device_exclusive_scan_i32(ctx, sc, is_h, hpos, N, &n_hill));
    CU_TRY(ctx, cudaMalloc(&pl.hidx, sizeof(int32_t) * (size_t)N));
    CU_TRY(ctx, cudaMalloc(&pl.cidx, sizeof(int32_t) * (size_t)N));
    LAUNCH(ctx, k_role_lists, G, B, 0, is_h, hpos, N, pl.hidx, pl.cidx);
    LAUNCH(ctx, k_role_level_first, grid_for(Lmax + 1, B), B, 0, level_first, hpos, Lmax + 1, d_hfirst);
    pl.h_hfirst.assign((size_t)Lmax + 2, 0);
    CU_TRY(ctx, cudaMemcpyAsync(pl.h_hfirst.data(), d_hfirst, sizeof(int32_t) * ((size_t)Lmax + 1), cudaMemcpyDeviceToHost, ctx->stream));
    pl.h_first.swap(h_first);
    // the plan outlives this call: its arrays move from the call's arena to buffers the context owns
    CU_TRY(ctx, cudaMalloc(&pl.tw, sizeof(uint32_t) * (size_t)N));
    CU_TRY(ctx, cudaMalloc(&pl.cs, sizeof(int32_t) * (size_t)N));
    CU_TRY(ctx, cudaMalloc(&pl.ctrl_cells, sizeof(int32_t) * ((size_t)nq + 1)));
    CU_TRY(ctx, cudaMalloc(&pl.ctrlh_cells, sizeof(int32_t) * ((size_t)nh + 1)));
    CU_TRY(ctx, cudaMemcpyAsync(pl.tw, tw, sizeof(uint32_t) * (size_t)N, cudaMemcpyDeviceToDevice, ctx->stream));
    CU_TRY(ctx, cudaMemcpyAsync(pl.cs, cs, sizeof(int32_t) * (size_t)N, cudaMemcpyDeviceToDevice, ctx->stream));
    CU_TRY(ctx, cudaMemcpyAsync(pl.ctrl_cells, ctrl_cells, sizeof(int32_t) * ((size_t)nq + 1), cudaMemcpyDeviceToDevice, ctx->stream));
    CU_TRY(ctx, cudaMemcpyAsync(pl.ctrlh_cells, ctrlh_cells, sizeof(int32_t) * ((size_t)nh + 1), cudaMemcpyDeviceToDevice, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));  // h_first / h_hfirst are complete
    pl.Lmax = Lmax; pl.nq = nq; pl.nh = nh; pl.n_hill = n_hill;
    return WMF_OK;
}

extern "C" {

int wmf_shia_run(wmf_ctx *ctx, const wmf_shia_cfg *cfg, const wmf_shia_inputs *in, wmf_shia_outputs *out) {
    if (!ctx || !cfg || !in || !out) return WMF_ERR_ARG;
    const int N = cfg->n_cel, T = cfg->n_reg, M = cfg->n_members > 0 ? cfg->n_members : 1;
    if (N <= 0 || T <= 0 || cfg->n_cont < 1 || cfg->n_conth < 1) WMF_FAIL(ctx, WMF_ERR_ARG, "shia: bad sizes");
    if (cfg->calc_niter < 1) WMF_FAIL(ctx, WMF_ERR_ARG, "shia: calc_niter must be >= 1");
    for (int k = 0; k < 3; ++k)
        if (cfg->speed_type[k] != 1 && cfg->speed_type[k] != 2)
            WMF_FAIL(ctx, WMF_ERR_ARG, "shia: speed_type(%d)=%d, only 1 (linear) and 2 (kinematic) exist", k + 1, cfg->speed_type[k]);
    if (!in->drena || !in->unit_type || !in->hill_long || !in->elem_area || !in->v_coef || !in->h_coef || !in->h_exp ||
        !in->max_capilar || !in->max_gravita || !in->max_aquifer || !in->evpserie || !in->calib || !in->rain || !in->pos_evento ||
        !in->stream_long || (!in->storage && !in->stoin))
        WMF_FAIL(ctx, WMF_ERR_ARG, "shia: a required input array is NULL");
    const bool sed = cfg->sim_sediments == 1, slides = cfg->sim_slides == 1;
    if (sed && (!in->krus || !in->crus || !in->prus || !in->parliac || !in->hill_slope || !in->stream_slope))
        WMF_FAIL(ctx, WMF_ERR_ARG, "shia: sim_sediments=1 needs krus, crus, prus, parliac, hill_slope, stream_slope");
    if (slides && (!in->sl_zs || !in->sl_gammas || !in->sl_cohesion || !in->sl_frictionangle || !in->sl_radslope))
        WMF_FAIL(ctx, WMF_ERR_ARG, "shia: sim_slides=1 needs the sl_* soil maps");
    const bool pos_pm = cfg->pos_per_member == 1 && M > 1;   // storm scenarios: one record sequence per member
    const int stride = in->drena_stride > 0 ? in->drena_stride : 1;
    if ((long long)N * M >= (1LL << 40)) WMF_FAIL(ctx, WMF_ERR_ARG, "shia: ensemble too large");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    Scope sc(ctx);
    ShiaP P;
    memset(&P, 0, sizeof P);
    // WMF_TIMING=1: host wall clock per phase on stderr (each mark synchronises the stream, so it perturbs overlap)
    const bool timing = getenv("WMF_TIMING") && atoi(getenv("WMF_TIMING")) != 0;
    auto t_prev = std::chrono::steady_clock::now();
    auto mark = [&](const char *what) {
        if (!timing) return;
        cudaStreamSynchronize(ctx->stream);
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[wmf_shia_run] %-10s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
        t_prev = now;
    };
    P.N = N; P.T = T; P.M = M; P.n_cont = cfg->n_cont; P.n_conth = cfg->n_conth;
    P.dt = cfg->dt; P.dxp = cfg->dxp; P.retorno_gr = cfg->retorno_gr; P.retorno_aq = cfg->retorno_aq;
    for (int k = 0; k < 3; ++k) P.st[k] = cfg->speed_type[k];
    P.niter = cfg->calc_niter;
    P.show_speed = cfg->show_speed; P.show_storage = cfg->show_storage && out->mean_storage;
    P.show_mean_speed = cfg->show_mean_speed && out->mean_speed;
    P.sed_factor = cfg->sed_factor;
    for (int k = 0; k < 3; ++k) { P.wi[k] = cfg->wi[k]; P.diam[k] = cfg->diametro[k]; }
    P.sl_gullie = cfg->sl_gullienogullie; P.sl_fs = cfg->sl_fs; P.sl_gammaw = cfg->sl_gammaw;
    P.drena_stride = stride;

    // ---- rainfall: the long pole of a run fed from host memory ----
    if (in->n_records < 1) WMF_FAIL(ctx, WMF_ERR_ARG, "shia: n_records must be >= 1");
    // Rainfall is the one big input (C2: 1.7 GB).  From pinned host memory it is streamed in chunks of whole records on a
    // second stream while the wavefront already runs: wave w only needs the records of the steps it has reached.
    RainStream rs;
    WMF_TRY(rs.begin(ctx, sc, in->rain, in->n_records, N, &P.rain));
    // pos_evento is needed on both sides: validate on the host
    const size_t n_pos = pos_pm ? (size_t)M * T : (size_t)T;
    std::vector<int32_t> h_pos(n_pos), h_used;
    if (Scope::on_device(in->pos_evento)) {
        CU_TRY(ctx, cudaMemcpyAsync(h_pos.data(), in->pos_evento, sizeof(int32_t) * n_pos, cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    } else {
        memcpy(h_pos.data(), in->pos_evento, sizeof(int32_t) * n_pos);
    }
    for (size_t k = 0; k < n_pos; ++k)
        if (h_pos[k] < 1 || h_pos[k] > in->n_records)
            WMF_FAIL(ctx, WMF_ERR_ARG, "shia: pos_evento(%d)=%d outside 1..n_records=%d", (int)(k % T) + 1, h_pos[k], in->n_records);
    P.pos_stride = pos_pm ? T : 0;
    // ---- stage inputs ----
    const int32_t *d_unit, *d_control, *d_controlh;
    WMF_TRY(sc.in(in->drena, (size_t)N * stride - (stride - 1), &P.drena));
    WMF_TRY(sc.in(in->unit_type, (size_t)N, &d_unit));
    WMF_TRY(sc.in(in->control, (size_t)N, &d_control));
    WMF_TRY(sc.in(in->control_h, (size_t)N, &d_controlh));
    WMF_TRY(sc.in(in->v_coef, (size_t)4 * N, &P.v_coef));
    WMF_TRY(sc.in(in->h_coef, (size_t)4 * N, &P.h_coef));
    WMF_TRY(sc.in(in->h_exp, (size_t)4 * N, &P.h_exp));
    if (((uintptr_t)P.v_coef | (uintptr_t)P.h_coef | (uintptr_t)P.h_exp) & 15)
        WMF_FAIL(ctx, WMF_ERR_ARG, "shia: v_coef/h_coef/h_exp device pointers must be 16-byte aligned");
    WMF_TRY(sc.in(in->max_capilar, (size_t)N, &P.max_capilar));
    WMF_TRY(sc.in(in->max_gravita, (size_t)N, &P.max_gravita));
    WMF_TRY(sc.in(in->max_aquifer, (size_t)N, &P.max_aquifer));
    WMF_TRY(sc.in(in->hill_long, (size_t)N, &P.hill_long));
    WMF_TRY(sc.in(in->hill_slope, (size_t)N, &P.hill_slope));
    WMF_TRY(sc.in(in->stream_long, (size_t)N, &P.stream_long));
    WMF_TRY(sc.in(in->stream_slope, (size_t)N, &P.stream_slope));
    WMF_TRY(sc.in(in->elem_area, (size_t)N, &P.elem_area));
    WMF_TRY(sc.in(in->evpserie, (size_t)T, &P.evp));
    WMF_TRY(sc.in(in->calib, (size_t)11 * M, &P.calib));
    WMF_TRY(sc.in(in->pos_evento, n_pos, &P.pos));
    const float *d_storage = nullptr, *d_stoin = nullptr, *d_hspeedin = nullptr;
    // StoIn / HspeedIn are honoured when their first element is > 0 (:271, :285)
    auto first_positive = [&](const float *p, bool *res) -> int {
        float v = 0.f;
        if (!p) { *res = false; return WMF_OK; }
        if (Scope::on_device(p)) {
            CU_TRY(ctx, cudaMemcpyAsync(&v, p, 4, cudaMemcpyDeviceToHost, ctx->stream));
            CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        } else v = p[0];
        *res = v > 0.0f;
        return WMF_OK;
    };
    bool use_stoin = false, use_hsin = false;
    WMF_TRY(first_positive(in->stoin, &use_stoin));
    WMF_TRY(first_positive(in->hspeedin, &use_hsin));
    if (!use_stoin && !in->storage) WMF_FAIL(ctx, WMF_ERR_ARG, "shia: StoIn(1,1) <= 0 and no module Storage given");
    if (use_stoin) WMF_TRY(sc.in(in->stoin, (size_t)5 * N, &d_stoin));
    else WMF_TRY(sc.in(in->storage, (size_t)5 * N, &d_storage));
    if (use_hsin) WMF_TRY(sc.in(in->hspeedin, (size_t)4 * N, &d_hspeedin));
    P.hs_state = use_hsin ? 1 : 0;

    // the small inputs are queued; now the first rain records start to travel while the sweep is planned
    WMF_TRY(rs.early(ctx, h_pos, T));
    mark("stage");
    // ---- plan: levels, child ranges, role lists -- cached on a checksum of the arrays it derives from ----
    // (a calibration loop or an ensemble re-runs the same basin thousands of times: nothing is re-planned, no allocation
    // and a single 8-byte read-back per run)
    wmf_shia_plan &pl = ctx->plan;
    {
        unsigned long long *d_hash, h_hash = 0;
        WMF_TRY(sc.alloc(&d_hash, 1));
        CU_TRY(ctx, cudaMemsetAsync(d_hash, 0, sizeof(unsigned long long), ctx->stream));
        LAUNCH(ctx, k_plan_hash, std::min(grid_for(N, 256), ctx->num_sms * 8), 256, 0, P.drena, stride, d_unit, d_control, d_controlh, N, d_hash);
        CU_TRY(ctx, cudaMemcpyAsync(&h_hash, d_hash, sizeof h_hash, cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        const bool hit = pl.valid && pl.N == N && pl.hash == h_hash && !(getenv("WMF_SHIA_PLAN_CACHE") && atoi(getenv("WMF_SHIA_PLAN_CACHE")) == 0);
        if (!hit) {
            pl.release();
            WMF_TRY(shia_build_plan(ctx, sc, P.drena, stride, N, d_unit, d_control, d_controlh, pl));
            pl.N = N;
            pl.hash = h_hash;
            pl.valid = true;
        }
    }
    const int32_t Lmax = pl.Lmax;
    const int64_t nq = pl.nq, nh = pl.nh, n_hill = pl.n_hill;
    const std::vector<int32_t> &h_first = pl.h_first, &h_hfirst = pl.h_hfirst;
    if (nq + 1 > cfg->n_cont) WMF_FAIL(ctx, WMF_ERR_ARG, "shia: %lld control cells need n_cont >= %lld, got %d", (long long)nq, (long long)nq + 1, cfg->n_cont);
    if (nh > cfg->n_conth) WMF_FAIL(ctx, WMF_ERR_ARG, "shia: %lld humidity control cells but n_conth = %d", (long long)nh, cfg->n_conth);
    const int B = 256, G = grid_for(N, B);
    P.tw = pl.tw; P.cs = pl.cs; P.Lmax = Lmax;
    P.ctrl_cells = pl.ctrl_cells; P.ctrlh_cells = pl.ctrlh_cells; P.n_ctrl = (int)nq; P.n_ctrlh = (int)nh;

    // ---- which kernel: the time-blocked wavefront unless an output only the one-step kernel produces is wanted ----
    const bool per_run = (M == 1);
    const bool tb_ok = !(cfg->show_storage && out->mean_storage) && !(cfg->show_mean_speed && out->mean_speed) &&
                       !(cfg->retorno_gr == 1.0f && cfg->show_mean_retorno && out->mean_retorno);
    int KT = tb_steps(T, tb_ok, (long long)N * M, Lmax, ctx->num_sms, M == 1);
    if (KT > 4 && !getenv("WMF_SHIA_K")) {  // the outflow ring holds 2 * K float4 per (cell, member): keep it under a third of the free memory
        size_t free_b = 0, total_b = 0;
        CU_TRY(ctx, cudaMemGetInfo(&free_b, &total_b));
        uint64_t reserved = 0, used = 0;  // blocks the stream-ordered pool keeps cached count as available
        cudaMemPoolGetAttribute(ctx->pool, cudaMemPoolAttrReservedMemCurrent, &reserved);
        cudaMemPoolGetAttribute(ctx->pool, cudaMemPoolAttrUsedMemCurrent, &used);
        if (reserved > used) free_b += (size_t)(reserved - used);
        while (KT > 4 && (double)2 * KT * N * M * sizeof(float4) > 0.33 * (double)free_b) KT /= 2;
    }
    const bool vdump = cfg->save_vfluxes == 1 && (out->vfl_snap || out->mean_vfluxes);
    const bool rcdump = cfg->save_rc == 1 && out->rc_coef;
    const bool retdump = cfg->save_retorno == 1 && out->ret_snap;
    const bool snap = (cfg->save_storage == 1 && out->sto_snap) || (cfg->save_speed == 1 && out->speed_snap) || vdump || rcdump || retdump;
    if (KT && (sed || slides || snap)) {  // sediment / slide variants are built for K = 4 and 8, dump variants for K = 4
        const char *e = getenv("WMF_SHIA_K");
        const bool k8 = !snap && T >= 8 && ((e && atoi(e) == 8) || (!e && M >= 4));
        KT = k8 ? 8 : ((T >= 4) ? 4 : 0);
    }
    if (M > 1 && (sed || slides) && KT == 0)
        WMF_FAIL(ctx, WMF_ERR_ARG, "shia: ensembles with sediments / slides run in the time-blocked kernel: n_reg >= 4 and no per-step means");
    const bool sepf = cfg->separate_fluxes == 1;
    const bool sepr = cfg->separate_rain == 1;
    const bool floods = cfg->sim_floods == 1;
    if (floods) {
        if (sed || slides || snap || !tb_ok || M != 1 || sepf || sepr)
            WMF_FAIL(ctx, WMF_ERR_ARG, "shia: sim_floods cannot be combined with sediments, slides, state dumps, show_* means, separate_* or ensembles");
        if (!in->flood_w || !in->flood_d50 || !in->flood_slope || !in->flood_sections || !in->flood_sec_cells || cfg->flood_sec_tam < 1)
            WMF_FAIL(ctx, WMF_ERR_ARG, "shia: sim_floods = 1 needs flood_w, flood_d50, flood_slope, flood_sections, flood_sec_cells");
        KT = (T >= 4) ? 4 : 1;
    }
    if (sepf || sepr) {
        if (sed || slides || snap || !tb_ok || M != 1 || (sepf && sepr))
            WMF_FAIL(ctx, WMF_ERR_ARG, "shia: separate_fluxes / separate_rain cannot be combined with each other, sediments, slides, state dumps, show_* means or ensembles");
        KT = (T >= 4) ? 4 : 1;
    }
    if (sepr && (!in->conv || !in->stra || !in->pos_conv || !in->pos_stra || in->n_conv_records < 1 || in->n_stra_records < 1))
        WMF_FAIL(ctx, WMF_ERR_ARG, "shia: separate_rain = 1 needs the conv / stra tables and pos_conv / pos_stra");
    if (KT) {
        P.hidx = pl.hidx; P.cidx = pl.cidx;
        P.nb = (T + KT - 1) / KT;
        P.bal_direct = 1;
    }
    mark("plan");
    // ---- state ----
    const size_t NM = (size_t)N * M;
    WMF_TRY(sc.alloc(&P.sto, 5 * NM));
    WMF_TRY(sc.alloc(&P.hsp, 4 * NM));
    double *d_sto0, *d_recsum;
    WMF_TRY(sc.alloc(&d_sto0, 1));
    WMF_TRY(sc.alloc(&d_recsum, (size_t)in->n_records));
    CU_TRY(ctx, cudaMemsetAsync(d_sto0, 0, sizeof(double), ctx->stream));
    CU_TRY(ctx, cudaMemsetAsync(d_recsum, 0, sizeof(double) * (size_t)in->n_records, ctx->stream));
    if (KT) {
        P.st4 = reinterpret_cast<float4 *>(P.sto);
        P.s5 = P.sto + 4 * NM;
        WMF_TRY(sc.alloc(&P.outb4, (size_t)2 * KT * NM));
        if (sepf) {
            WMF_TRY(sc.alloc(&P.flx, 3 * NM));
            WMF_TRY(sc.alloc(&P.flxout4, (size_t)2 * KT * NM));
            CU_TRY(ctx, cudaMemsetAsync(P.flx, 0, sizeof(float) * 3 * NM, ctx->stream));  // Fluxes = 0 (:343)
        }
        if (sepr) {
            WMF_TRY(sc.alloc(&P.shc, 5 * NM));
            WMF_TRY(sc.alloc(&P.shs, 5 * NM));
            WMF_TRY(sc.alloc(&P.cvout4, (size_t)2 * KT * NM));
            WMF_TRY(sc.alloc(&P.srout4, (size_t)2 * KT * NM));
            CU_TRY(ctx, cudaMemsetAsync(P.shc, 0, sizeof(float) * 5 * NM, ctx->stream));  // Storage_conv = Storage_stra = 0 (:351-352)
            CU_TRY(ctx, cudaMemsetAsync(P.shs, 0, sizeof(float) * 5 * NM, ctx->stream));
            WMF_TRY(sc.in(in->conv, (size_t)in->n_conv_records * N, &P.conv));
            WMF_TRY(sc.in(in->stra, (size_t)in->n_stra_records * N, &P.stra));
            // record in force at every step: Conv / Stra are read on wet steps only (:452-456); a record outside the file
            // fails the read and leaves the vector as it was
            std::vector<int32_t> pe((size_t)T), pc((size_t)T), ps((size_t)T), eff((size_t)2 * T);
            const int32_t *src[3] = {in->pos_evento, in->pos_conv, in->pos_stra};
            int32_t *dst[3] = {pe.data(), pc.data(), ps.data()};
            for (int k = 0; k < 3; ++k) {
                if (Scope::on_device(src[k])) CU_TRY(ctx, cudaMemcpy(dst[k], src[k], sizeof(int32_t) * T, cudaMemcpyDeviceToHost));
                else memcpy(dst[k], src[k], sizeof(int32_t) * T);
            }
            int32_t ec = 0, es = 0;
            for (int t = 0; t < T; ++t) {
                if (pe[t] != 1) {
                    if (pc[t] >= 1 && pc[t] <= in->n_conv_records) ec = pc[t];
                    if (ps[t] >= 1 && ps[t] <= in->n_stra_records) es = ps[t];
                }
                eff[t] = ec;
                eff[(size_t)T + t] = es;
            }
            int32_t *d_eff;
            WMF_TRY(sc.alloc(&d_eff, (size_t)2 * T));
            CU_TRY(ctx, cudaMemcpyAsync(d_eff, eff.data(), sizeof(int32_t) * 2 * T, cudaMemcpyHostToDevice, ctx->stream));
            CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));  // `eff` goes out of scope
            P.effc = d_eff;
            P.effs = d_eff + T;
        }
        if (floods) {
            const int S = cfg->flood_sec_tam;
            P.fl_dw = cfg->flood_dw; P.fl_dsed = cfg->flood_dsed; P.fl_av = cfg->flood_av; P.fl_cmax = cfg->flood_cmax;
            P.fl_umbral = cfg->flood_umbral; P.fl_step = cfg->flood_step; P.fl_hmax = cfg->flood_hmax; P.fl_S = S;
            WMF_TRY(sc.in(in->flood_w, (size_t)N, &P.fl_w)); WMF_TRY(sc.in(in->flood_d50, (size_t)N, &P.fl_d50));
            WMF_TRY(sc.in(in->flood_slope, (size_t)N, &P.fl_slope));
            WMF_TRY(sc.in(in->flood_sections, (size_t)S * N, &P.fl_sections));
            WMF_TRY(sc.in(in->flood_sec_cells, (size_t)S * N, &P.fl_cells));
            WMF_TRY(sc.alloc(&P.fl_out, (size_t)8 * N)); WMF_TRY(sc.alloc(&P.fl_last, (size_t)N));
            WMF_TRY(sc.alloc(&P.fl_key, (size_t)N));
            WMF_TRY(sc.out(out->flood_profundidad, (size_t)S * N, &P.fl_prof));
            if (!P.fl_prof) WMF_TRY(sc.alloc(&P.fl_prof, (size_t)S * N));
            CU_TRY(ctx, cudaMemsetAsync(P.fl_out, 0, sizeof(float) * 8 * (size_t)N, ctx->stream));
            CU_TRY(ctx, cudaMemsetAsync(P.fl_last, 0xff, sizeof(int32_t) * (size_t)N, ctx->stream));
            CU_TRY(ctx, cudaMemsetAsync(P.fl_key, 0, sizeof(unsigned long long) * (size_t)N, ctx->stream));
            CU_TRY(ctx, cudaMemsetAsync(P.fl_prof, 0, sizeof(float) * (size_t)S * N, ctx->stream));
        }
        LAUNCH(ctx, k_state_init_tb, G, B, 0, use_stoin ? d_stoin : d_storage, use_stoin ? 0.0f : cfg->storage_constant, N, M, P.st4, P.s5);
    } else {
        WMF_TRY(sc.alloc(&P.outb, 8 * NM));
        LAUNCH(ctx, k_state_init, G, B, 0, use_stoin ? d_stoin : d_storage, use_stoin ? 0.0f : cfg->storage_constant, N, M, P.sto, d_sto0);
    }
    LAUNCH(ctx, k_hspeed_init, grid_for((long long)NM, B), B, 0, d_hspeedin, N, M, P.hsp);
    // ---- outputs ----
    const size_t QN = (size_t)M * cfg->n_cont * T, HN = (size_t)M * cfg->n_conth * T;
    WMF_TRY(sc.out(out->q, QN, &P.q));
    WMF_TRY(sc.out(out->speed, QN, &P.speed));
    WMF_TRY(sc.out(out->areacontrol, QN, &P.area));
    WMF_TRY(sc.out(out->hum, HN, &P.hum));
    WMF_TRY(sc.out(out->st1, HN, &P.st1o));
    WMF_TRY(sc.out(out->st3, HN, &P.st3o));
    if (sepf) {
        WMF_TRY(sc.out(out->qseparated, QN * 3, &P.qsep));
        if (P.qsep) CU_TRY(ctx, cudaMemsetAsync(P.qsep, 0, sizeof(float) * QN * 3, ctx->stream));
    }
    if (sepr) {
        WMF_TRY(sc.out(out->qsep_byrain, QN * 2, &P.qsepr));
        if (P.qsepr) CU_TRY(ctx, cudaMemsetAsync(P.qsepr, 0, sizeof(float) * QN * 2, ctx->stream));
    }
    if (P.q) CU_TRY(ctx, cudaMemsetAsync(P.q, 0, sizeof(float) * QN, ctx->stream));
    if (P.speed) CU_TRY(ctx, cudaMemsetAsync(P.speed, 0, sizeof(float) * QN, ctx->stream));
    if (P.area) CU_TRY(ctx, cudaMemsetAsync(P.area, 0, sizeof(float) * QN, ctx->stream));
    if (P.hum) CU_TRY(ctx, cudaMemsetAsync(P.hum, 0, sizeof(float) * HN, ctx->stream));
    if (P.st1o) CU_TRY(ctx, cudaMemsetAsync(P.st1o, 0, sizeof(float) * HN, ctx->stream));
    if (P.st3o) CU_TRY(ctx, cudaMemsetAsync(P.st3o, 0, sizeof(float) * HN, ctx->stream));
    float *d_stoout, *d_hsout, *d_balance = nullptr, *d_mean_rain = nullptr, *d_acum_rain = nullptr, *d_mean_storage = nullptr,
          *d_mean_speed = nullptr, *d_mean_retorno = nullptr;
    WMF_TRY(sc.out(out->stoout, 5 * NM, &d_stoout));
    WMF_TRY(sc.out(out->hspeed_out, 4 * NM, &d_hsout));
    if (per_run) {
        WMF_TRY(sc.out(out->balance, (size_t)T, &d_balance));
        WMF_TRY(sc.out(out->mean_rain, (size_t)T, &d_mean_rain));
        WMF_TRY(sc.out(out->acum_rain, (size_t)N, &d_acum_rain));
        if (cfg->show_storage) WMF_TRY(sc.out(out->mean_storage, (size_t)5 * T, &d_mean_storage));
        if (cfg->show_mean_speed) WMF_TRY(sc.out(out->mean_speed, (size_t)4 * T, &d_mean_speed));
        if (cfg->retorno_gr == 1.0f && cfg->show_mean_retorno) WMF_TRY(sc.out(out->mean_retorno, (size_t)T, &d_mean_retorno));
        if (cfg->retorno_gr == 1.0f && (out->retorned || d_mean_retorno)) {
            P.want_retorned = 1;
            if (out->retorned && !(cfg->show_mean_retorno == 1 || cfg->save_retorno == 1)) {
                WMF_TRY(sc.out(out->retorned, (size_t)N, &P.retorned));
                CU_TRY(ctx, cudaMemsetAsync(P.retorned, 0, sizeof(float) * (size_t)N, ctx->stream));
            }
        }
        P.want_red = (d_balance || d_mean_storage || d_mean_speed || d_mean_retorno) ? 1 : 0;
    }
    float *d_mean_vfluxes = nullptr;
    if (per_run && vdump && out->mean_vfluxes) {
        WMF_TRY(sc.out(out->mean_vfluxes, (size_t)4 * T, &d_mean_vfluxes));
        P.want_vmean = 1;
    }
    if (P.want_red || P.want_vmean) {
        WMF_TRY(sc.alloc(&P.red, (size_t)T * RQ_COUNT * RED_SUB));
        CU_TRY(ctx, cudaMemsetAsync(P.red, 0, sizeof(double) * (size_t)T * RQ_COUNT * RED_SUB, ctx->stream));
    }
    // ---- state dumps ----
    float *d_rc_out = nullptr;
    // guarda_* (n_reg): row of the snapshot table that receives each flagged step, -1 elsewhere
    auto make_slots = [&](const int32_t *guarda, const int32_t **d_slot_out, int *n_out) -> int {
        std::vector<int32_t> g((size_t)T, 0), slot((size_t)T);
        if (guarda) {
            if (Scope::on_device(guarda)) {
                CU_TRY(ctx, cudaMemcpyAsync(g.data(), guarda, sizeof(int32_t) * T, cudaMemcpyDeviceToHost, ctx->stream));
                CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
            } else {
                memcpy(g.data(), guarda, sizeof(int32_t) * T);
            }
        }
        int n = 0;
        for (int t = 0; t < T; ++t) slot[t] = (g[t] > 0) ? n++ : -1;
        int32_t *d_slot;
        WMF_TRY(sc.alloc(&d_slot, (size_t)T));
        CU_TRY(ctx, cudaMemcpyAsync(d_slot, slot.data(), sizeof(int32_t) * T, cudaMemcpyHostToDevice, ctx->stream));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));  // `slot` goes out of scope
        *d_slot_out = d_slot;
        *n_out = n;
        return WMF_OK;
    };
    if (vdump || rcdump || retdump) {
        if (M != 1) WMF_FAIL(ctx, WMF_ERR_ARG, "shia: save_vfluxes / save_rc / save_retorno need a single run (n_members = 1)");
        if (vdump && out->vfl_snap) {
            int n_vsnap = 0;
            WMF_TRY(make_slots(in->guarda_vfluxes, &P.vsnap_slot, &n_vsnap));
            WMF_TRY(sc.out(out->vfl_snap, (size_t)std::max(n_vsnap, 1) * 4 * N, &P.vfl_snap));
            if (reinterpret_cast<uintptr_t>(P.vfl_snap) & 15) WMF_FAIL(ctx, WMF_ERR_ARG, "shia: vfl_snap must be 16-byte aligned");
        }
        if (rcdump) {
            WMF_TRY(sc.alloc(&P.rc, (size_t)2 * N));
            CU_TRY(ctx, cudaMemsetAsync(P.rc, 0, sizeof(float) * 2 * (size_t)N, ctx->stream));  // rc_coef = 0 (:425)
            WMF_TRY(sc.out(out->rc_coef, (size_t)2 * N, &d_rc_out));
        }
        if (retdump) WMF_TRY(sc.out(out->ret_snap, (size_t)T * N, &P.ret_snap));
    }
    if (cfg->save_storage == 1 || cfg->save_speed == 1) {
        if (M != 1) WMF_FAIL(ctx, WMF_ERR_ARG, "shia: save_storage / save_speed need a single run (n_members = 1)");
        if (cfg->save_storage == 1 && out->sto_snap) {
            if (!in->guarda_cond) WMF_FAIL(ctx, WMF_ERR_ARG, "shia: save_storage = 1 needs guarda_cond");
            int n_snap = 0;
            WMF_TRY(make_slots(in->guarda_cond, &P.snap_slot, &n_snap));
            WMF_TRY(sc.out(out->sto_snap, (size_t)std::max(n_snap, 1) * 5 * N, &P.sto_snap));
        }
        if (cfg->save_speed == 1 && out->speed_snap) {
            WMF_TRY(sc.out(out->speed_snap, (size_t)T * 4 * N, &P.spd_snap));
            if (reinterpret_cast<uintptr_t>(P.spd_snap) & 15) WMF_FAIL(ctx, WMF_ERR_ARG, "shia: speed_snap must be 16-byte aligned");
        }
    }
    // ---- sediments / slides ----
    float *d_zmin = nullptr, *d_zcrit = nullptr, *d_zmax = nullptr, *d_bo = nullptr, *d_risk = nullptr;
    if (sed) {
        WMF_TRY(sc.in(in->krus, (size_t)N, &P.krus));
        WMF_TRY(sc.in(in->crus, (size_t)N, &P.crus));
        WMF_TRY(sc.in(in->prus, (size_t)N, &P.prus));
        WMF_TRY(sc.in(in->parliac, (size_t)3 * N, &P.parliac));
        const float *d_vd0;
        WMF_TRY(sc.in(in->vd_init, (size_t)3 * N, &d_vd0));
        WMF_TRY(sc.alloc(&P.sed, 12 * NM));
        if (KT) {
            WMF_TRY(sc.alloc(&P.sedout4, (size_t)2 * KT * 3 * NM));
            WMF_TRY(sc.alloc(&P.sedstg4, (size_t)2 * KT * 2 * NM));
            float *aso;
            WMF_TRY(sc.alloc(&aso, (size_t)N));
            if (cfg->speed_type[0] == 2) {
                if (!P.hill_slope) WMF_FAIL(ctx, WMF_ERR_ARG, "shia: sediments with a kinematic tank 2 need hill_slope");
                LAUNCH(ctx, k_sed_aso, G, B, 0, P.hill_slope, P.sed_factor, N, aso);
            }
            P.sed_aso = aso;
        } else WMF_TRY(sc.alloc(&P.sedout, 24 * NM));
        WMF_TRY(sc.alloc(&P.volero, NM));
        WMF_TRY(sc.alloc(&P.voldepo, NM));
        WMF_TRY(sc.alloc(&P.sedtot, 6));
        CU_TRY(ctx, cudaMemsetAsync(P.volero, 0, sizeof(float) * NM, ctx->stream));
        CU_TRY(ctx, cudaMemsetAsync(P.voldepo, 0, sizeof(float) * NM, ctx->stream));
        CU_TRY(ctx, cudaMemsetAsync(P.sedtot, 0, sizeof(double) * 6, ctx->stream));
        if (KT) {
            WMF_TRY(sc.alloc(&P.sedacc, 6 * NM));
            CU_TRY(ctx, cudaMemsetAsync(P.sedacc, 0, sizeof(double) * 6 * NM, ctx->stream));
        }
        if (!KT) CU_TRY(ctx, cudaMemsetAsync(P.sedout, 0, sizeof(float) * 24 * NM, ctx->stream));
        LAUNCH(ctx, k_sed_init, grid_for((long long)NM, B), B, 0, d_vd0, N, M, P.sed);
        WMF_TRY(sc.out(out->qsed, QN * 3, &P.qsed));
        if (P.qsed) CU_TRY(ctx, cudaMemsetAsync(P.qsed, 0, sizeof(float) * QN * 3, ctx->stream));
    }
    if (slides) {
        WMF_TRY(sc.in(in->sl_zs, (size_t)N, &P.sl_zs));
        WMF_TRY(sc.in(in->sl_gammas, (size_t)N, &P.sl_gammas));
        WMF_TRY(sc.in(in->sl_cohesion, (size_t)N, &P.sl_cohesion));
        WMF_TRY(sc.in(in->sl_frictionangle, (size_t)N, &P.sl_friction));
        WMF_TRY(sc.in(in->sl_radslope, (size_t)N, &P.sl_radslope));
        WMF_TRY(sc.alloc(&d_zmin, (size_t)N)); WMF_TRY(sc.alloc(&d_zcrit, (size_t)N)); WMF_TRY(sc.alloc(&d_zmax, (size_t)N));
        WMF_TRY(sc.alloc(&d_bo, (size_t)N)); WMF_TRY(sc.alloc(&d_risk, (size_t)N));
        WMF_TRY(sc.alloc(&P.sl_slid, NM)); WMF_TRY(sc.alloc(&P.sl_marktime, NM)); WMF_TRY(sc.alloc(&P.sl_acum, NM));
        WMF_TRY(sc.out(out->sl_slidencelltime, (size_t)M * T, &P.sl_ncelltime));
        if (!P.sl_ncelltime) WMF_TRY(sc.alloc(&P.sl_ncelltime, (size_t)M * T));
        CU_TRY(ctx, cudaMemsetAsync(P.sl_ncelltime, 0, sizeof(int32_t) * (size_t)M * T, ctx->stream));
        LAUNCH(ctx, k_slide_state_init, grid_for((long long)NM, B), B, 0, P);
        if (KT && cfg->sl_gullienogullie == 1) {
            int32_t *len;
            WMF_TRY(sc.alloc(&len, (size_t)N));
            WMF_TRY(sc.alloc(&P.sl_ring, 2 * NM));
            LAUNCH(ctx, k_slide_len, G, B, 0, P.drena, stride, P.tw, N, len);
            P.sl_len = len;
        }
        float *d_cosr, *d_sinr, *d_tanfa;
        WMF_TRY(sc.alloc(&d_cosr, (size_t)N)); WMF_TRY(sc.alloc(&d_sinr, (size_t)N)); WMF_TRY(sc.alloc(&d_tanfa, (size_t)N));
        LAUNCH(ctx, k_slide_allocate, G, B, 0, P, d_zmin, d_zcrit, d_zmax, d_bo, d_risk, d_cosr, d_sinr, d_tanfa);
        P.sl_zcrit = d_zcrit;
        P.sl_risk = d_risk;
        P.sl_cosr = d_cosr; P.sl_sinr = d_sinr; P.sl_tanfa = d_tanfa;
    }

    mark("setup");
    // ---- the wavefront ----
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));  // h_first is complete
    auto level_end = [&](int L) { return L == 0 ? N : h_first[L - 1]; };
    const int n_waves = (KT ? P.nb : T) + Lmax;
    CU_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    WMF_TRY(rs.schedule(ctx, h_first, Lmax, h_pos, KT ? KT : 1));
    ctx->sed_seq = 0;
    ctx->sed_two_streams = false;
    // whatever way this call ends, the compute stream waits for the sediment stream before the scope's buffers are released
    struct SedJoin {
        wmf_ctx *c;
        ~SedJoin() {
            if (c->sed_two_streams && c->sed_seq > 0) cudaStreamWaitEvent(c->stream, c->sed_ev_s[(c->sed_seq - 1) & 3], 0);
        }
    } sed_join{ctx};
    if (KT && sed) {
        const char *e = getenv("WMF_SED_STREAMS");
        ctx->sed_two_streams = e ? atoi(e) == 2 : true;   // C5: one scenario 122 -> 106 ms, eight 8.07 -> 8.46 G cell-steps/s
        if (ctx->sed_two_streams && !ctx->sed_stream) {
            CU_TRY(ctx, cudaStreamCreateWithFlags(&ctx->sed_stream, cudaStreamNonBlocking));
            for (int i = 0; i < 4; ++i) {
                CU_TRY(ctx, cudaEventCreateWithFlags(&ctx->sed_ev_h[i], cudaEventDisableTiming));
                CU_TRY(ctx, cudaEventCreateWithFlags(&ctx->sed_ev_s[i], cudaEventDisableTiming));
            }
        }
    }
    if (KT) {
        const int nb = P.nb;
        bool launched_any = false, use_pdl = true;
        if (const char *e = getenv("WMF_SHIA_PDL")) use_pdl = atoi(e) != 0;
        for (int w = 0; w < n_waves; ++w) {
            const int Lhi = std::min(Lmax, Lmax - w + nb - 1), Llo = std::max(0, Lmax - w);
            const int lo = h_first[Lhi], hi = level_end(Llo);
            if (hi <= lo) continue;
            WMF_TRY(rs.wait_wave(ctx, w));
            const int hlo = h_hfirst[Lhi], hhi = (Llo == 0) ? (int)n_hill : h_hfirst[Llo - 1];
            const int clo = lo - hlo, chi = hi - hhi;
            const bool pdl = use_pdl && launched_any;  // the first wave follows memsets and plan kernels: plain launch
            launched_any = true;
            if (floods) {  // HydroFlash: K = 4 (or 1 for runs shorter than 4 steps)
                if (KT == 4) launch_tb<4, false, false, false, false, false, true>(ctx, P, w, clo, chi, hlo, hhi, pdl);
                else launch_tb<1, false, false, false, false, false, true>(ctx, P, w, clo, chi, hlo, hhi, pdl);
                continue;
            }
            if (sepr) {  // separate_rain: K = 4 (or 1 for runs shorter than 4 steps)
                if (KT == 4) launch_tb<4, false, false, false, false, true>(ctx, P, w, clo, chi, hlo, hhi, pdl);
                else launch_tb<1, false, false, false, false, true>(ctx, P, w, clo, chi, hlo, hhi, pdl);
                continue;
            }
            if (sepf) {  // separate_fluxes: K = 4 (or 1 for runs shorter than 4 steps)
                if (KT == 4) launch_tb<4, false, false, false, true>(ctx, P, w, clo, chi, hlo, hhi, pdl);
                else launch_tb<1, false, false, false, true>(ctx, P, w, clo, chi, hlo, hhi, pdl);
                continue;
            }
            if (snap) {  // state dumps: K = 4 variants with the snapshot stores compiled in
                if (sed && slides) launch_tb<4, true, true, true>(ctx, P, w, clo, chi, hlo, hhi, pdl, use_pdl);
                else if (sed) launch_tb<4, true, false, true>(ctx, P, w, clo, chi, hlo, hhi, pdl, use_pdl);
                else if (slides) launch_tb<4, false, true, true>(ctx, P, w, clo, chi, hlo, hhi, pdl);
                else launch_tb<4, false, false, true>(ctx, P, w, clo, chi, hlo, hhi, pdl);
                continue;
            }
            if (sed || slides) {
                if (KT == 8) {
                    if (sed && slides) launch_tb<8, true, true>(ctx, P, w, clo, chi, hlo, hhi, pdl, use_pdl);
                    else if (sed) launch_tb<8, true, false>(ctx, P, w, clo, chi, hlo, hhi, pdl, use_pdl);
                    else launch_tb<8, false, true>(ctx, P, w, clo, chi, hlo, hhi, pdl);
                    continue;
                }
                if (sed && slides) launch_tb<4, true, true>(ctx, P, w, clo, chi, hlo, hhi, pdl, use_pdl);
                else if (sed) launch_tb<4, true, false>(ctx, P, w, clo, chi, hlo, hhi, pdl, use_pdl);
                else launch_tb<4, false, true>(ctx, P, w, clo, chi, hlo, hhi, pdl);
                continue;
            }
            switch (KT) {
                case 1: launch_tb<1>(ctx, P, w, clo, chi, hlo, hhi, pdl); break;
                case 2: launch_tb<2>(ctx, P, w, clo, chi, hlo, hhi, pdl); break;
                case 4: launch_tb<4>(ctx, P, w, clo, chi, hlo, hhi, pdl); break;
                case 8: launch_tb<8>(ctx, P, w, clo, chi, hlo, hhi, pdl); break;
                case 16: launch_tb<16>(ctx, P, w, clo, chi, hlo, hhi, pdl); break;
                default: launch_tb<32>(ctx, P, w, clo, chi, hlo, hhi, pdl); break;
            }
        }
    } else {
        for (int w = 0; w < n_waves; ++w) {
            const int Lhi = std::min(Lmax, Lmax - w + T - 1), Llo = std::max(0, Lmax - w);
            const int lo = h_first[Lhi], hi = level_end(Llo);
            if (hi <= lo) continue;
            WMF_TRY(rs.wait_wave(ctx, w));
            if (sed && slides) launch_wave<true, true>(ctx, P, w, lo, hi);
            else if (sed) launch_wave<true, false>(ctx, P, w, lo, hi);
            else if (slides) launch_wave<false, true>(ctx, P, w, lo, hi);
            else launch_wave<false, false>(ctx, P, w, lo, hi);
        }
    }
    if (ctx->sed_two_streams && ctx->sed_seq > 0)   // the epilogue reads what the last sediment kernel wrote
        CU_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->sed_ev_s[(ctx->sed_seq - 1) & 3], 0));
    CU_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    CU_TRY(ctx, cudaGetLastError());
    WMF_TRY(rs.wait_all(ctx));
    // rain statistics that do not depend on the model state
    if (per_run) {
        h_used.assign((size_t)in->n_records, 0);       // (function scope: alive until the final synchronisation)
        for (int t = 0; t < T; ++t) h_used[h_pos[t] - 1] = 1;
        int32_t *d_used;
        WMF_TRY(sc.alloc(&d_used, (size_t)in->n_records));
        CU_TRY(ctx, cudaMemcpyAsync(d_used, h_used.data(), sizeof(int32_t) * (size_t)in->n_records, cudaMemcpyHostToDevice, ctx->stream));
        for (int r0 = 0; r0 < in->n_records; r0 += 65535) {
            dim3 rg((unsigned)std::min(grid_for(N, 1024), 128), (unsigned)std::min(65535, in->n_records - r0));
            LAUNCH(ctx, k_rain_record_sum, rg, 256, 0, P.rain, N, r0, d_used, d_recsum);
        }
        if (d_acum_rain) LAUNCH(ctx, k_acum_rain, G, B, 0, P.rain, P.pos, N, T, d_acum_rain);
    }
    ctx->last_levels = Lmax + 1;
    ctx->last_waves = n_waves;

    mark("waves");
    // ---- epilogue ----
    if (d_stoout || d_hsout) {
        if (KT) LAUNCH(ctx, k_state_final_tb, grid_for((long long)NM, B), B, 0, P, d_stoout, d_hsout);
        else LAUNCH(ctx, k_state_final, grid_for((long long)NM, B), B, 0, P, d_stoout, d_hsout);
    }
    if (per_run && (d_balance || d_mean_rain || d_mean_storage || d_mean_speed || d_mean_retorno || d_mean_vfluxes)) {
        double *red = P.red;
        if (!red) {  // only mean_rain wanted
            WMF_TRY(sc.alloc(&red, (size_t)T * RQ_COUNT * RED_SUB));
            CU_TRY(ctx, cudaMemsetAsync(red, 0, sizeof(double) * (size_t)T * RQ_COUNT * RED_SUB, ctx->stream));
        }
        LAUNCH(ctx, k_shia_epilogue, grid_for(T, 128), 128, 0, red, d_recsum, P.pos, d_sto0, N, T, P.show_storage, P.bal_direct,
               d_balance, d_mean_rain, d_mean_storage, d_mean_speed, d_mean_retorno, d_mean_vfluxes);
    }
    if (d_rc_out) LAUNCH(ctx, k_rc_final, G, B, 0, P.rc, N, d_rc_out);
    if (floods) {
        int32_t *d_ff = nullptr;
        float *d_sc = nullptr;
        unsigned long long *best;
        WMF_TRY(sc.out(out->flood_flood, (size_t)N, &d_ff));
        WMF_TRY(sc.alloc(&best, 1));
        CU_TRY(ctx, cudaMemsetAsync(best, 0, sizeof(unsigned long long), ctx->stream));
        LAUNCH(ctx, k_flood_final, G, B, 0, P.fl_key, P.fl_last, N, d_ff, best);
        float *outs[6] = {out->flood_h, out->flood_ufr, out->flood_cr, out->flood_q, out->flood_qsed, out->flood_speed};
        for (int k = 0; k < 6; ++k) {
            float *d;
            WMF_TRY(sc.out(outs[k], (size_t)N, &d));
            if (d) CU_TRY(ctx, cudaMemcpyAsync(d, P.fl_out + (size_t)k * N, sizeof(float) * (size_t)N, cudaMemcpyDeviceToDevice, ctx->stream));
        }
        WMF_TRY(sc.out(out->flood_scalars, 3, &d_sc));
        if (d_sc) LAUNCH(ctx, k_flood_scalars, 1, 32, 0, best, P.fl_out, P.fl_last, N, d_sc);
    }
    if (sepr) {
        float *oc = nullptr, *os = nullptr;
        WMF_TRY(sc.out(out->storage_conv, 5 * NM, &oc));
        WMF_TRY(sc.out(out->storage_stra, 5 * NM, &os));
        if (oc) LAUNCH(ctx, k_shadow_final, G, B, 0, P.shc, N, oc);
        if (os) LAUNCH(ctx, k_shadow_final, G, B, 0, P.shs, N, os);
    }
    if (sed) {
        float *vs, *vd, *vsc, *vdc, *ve, *vdp, *erot, *dept;
        WMF_TRY(sc.out(out->vs, 3 * NM, &vs)); WMF_TRY(sc.out(out->vd, 3 * NM, &vd));
        WMF_TRY(sc.out(out->vsc, 3 * NM, &vsc)); WMF_TRY(sc.out(out->vdc, 3 * NM, &vdc));
        LAUNCH(ctx, k_sed_final, grid_for((long long)NM, B), B, 0, P.sed, N, M, vs, vd, vsc, vdc);
        WMF_TRY(sc.out(out->volero, NM, &ve)); WMF_TRY(sc.out(out->voldepo, NM, &vdp));
        if (ve) LAUNCH(ctx, k_member_major_f32, grid_for((long long)NM, B), B, 0, P.volero, N, M, ve);
        if (vdp) LAUNCH(ctx, k_member_major_f32, grid_for((long long)NM, B), B, 0, P.voldepo, N, M, vdp);
        WMF_TRY(sc.out(out->erot, (size_t)3 * M, &erot)); WMF_TRY(sc.out(out->dept, (size_t)3 * M, &dept));
        if (KT) {
            if (erot || dept) LAUNCH(ctx, k_sed_totals, dim3((unsigned)M, 6), 256, 0, P.sedacc, N, M, erot, dept);
        } else {
            if (erot) LAUNCH(ctx, k_d2f, 1, 32, 0, P.sedtot, 3, erot);
            if (dept) LAUNCH(ctx, k_d2f, 1, 32, 0, P.sedtot + 3, 3, dept);
        }
    }
    if (slides) {
        float *o;
        auto copy_out = [&](float *dst, const float *src) -> int {
            if (!dst) return WMF_OK;
            float *d;
            WMF_TRY(sc.out(dst, (size_t)N, &d));
            CU_TRY(ctx, cudaMemcpyAsync(d, src, sizeof(float) * (size_t)N, cudaMemcpyDeviceToDevice, ctx->stream));
            return WMF_OK;
        };
        WMF_TRY(copy_out(out->sl_zmin, d_zmin)); WMF_TRY(copy_out(out->sl_zcrit, d_zcrit));
        WMF_TRY(copy_out(out->sl_zmax, d_zmax)); WMF_TRY(copy_out(out->sl_bo, d_bo));
        WMF_TRY(copy_out(out->sl_riskvector, d_risk));
        int32_t *oa;
        WMF_TRY(sc.out(out->sl_slideocurrence, NM, &o));
        WMF_TRY(sc.out(out->sl_slideacumulate, NM, &oa));
        if (o || oa) LAUNCH(ctx, k_slide_final, grid_for((long long)NM, B), B, 0, P, o, oa);
    }
    if (per_run && out->retorned && cfg->retorno_gr == 1.0f && (cfg->show_mean_retorno == 1 || cfg->save_retorno == 1)) {
        // the Fortran zeroes Retorned after every step in this mode (:841-843)
        float *r;
        WMF_TRY(sc.out(out->retorned, (size_t)N, &r));
        CU_TRY(ctx, cudaMemsetAsync(r, 0, sizeof(float) * (size_t)N, ctx->stream));
    }
    mark("epilogue");
    WMF_TRY(sc.finish());
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    mark("d2h");
    cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1);
    return WMF_OK;
}

int wmf_eval_scores(wmf_ctx *ctx, const float *q, int32_t n_members, int32_t n_cont, int32_t n_reg, int32_t row,
                    const float *qobs, float *scores) {
    if (!ctx || !q || !qobs || !scores || n_members <= 0 || n_cont <= 0 || n_reg <= 0 || row < 0 || row >= n_cont)
        return WMF_ERR_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    Scope sc(ctx);
    const float *d_q, *d_o;
    float *d_s;
    WMF_TRY(sc.in(q, (size_t)n_members * n_cont * n_reg, &d_q));
    WMF_TRY(sc.in(qobs, (size_t)n_reg, &d_o));
    WMF_TRY(sc.out(scores, (size_t)2 * n_members, &d_s));
    LAUNCH(ctx, k_eval_scores, n_members, 256, 0, d_q, n_cont, n_reg, row, d_o, d_s);
    return sc.finish();
}

int wmf_calc_speed(wmf_ctx *ctx, float sm, float coef, float expo, float elem_long, float dt, int32_t calc_niter,
                   float *speed_inout, float *area_out) {
    if (!ctx || !speed_inout || !area_out || calc_niter < 1) return WMF_ERR_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    Scope sc(ctx);
    float *d;
    WMF_TRY(sc.alloc(&d, 2));
    CU_TRY(ctx, cudaMemcpyAsync(d, speed_inout, 4, cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(ctx, k_calc_speed, 1, 1, 0, sm, coef, expo, elem_long, dt, calc_niter, d);
    float h[2];
    CU_TRY(ctx, cudaMemcpyAsync(h, d, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    *speed_inout = h[0];
    *area_out = h[1];
    return WMF_OK;
}

}  // extern "C"
