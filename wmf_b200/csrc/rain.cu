// rain.cu -- rainfall interpolation on the GPU: rain_idw (modelosv2.f90:1195-1271) and rain_mit (:1104-1194), cells
// model.  The reference interpolates one field per time step on the host, writes the wet ones to a direct-access
// file and shia_v1 reads them back record by record; here the record table is produced in HBM in the layout
// wmf_shia_run consumes (record 1 = zeros, record k = k-th wet step, int32 mm*1000 truncated), so a run can go from
// station series to hydrograph without the table ever crossing PCIe.
//
//   k_idw_weights   W(i,cell) = 1 / sqrt(dx^2+dy^2)**pp, station-major so that a warp reads 32 consecutive cells
//   k_rain_field    one thread per cell, RAIN_TCH time steps per thread in registers; the station loop is the outer
//                   loop, so each weight is loaded once per RAIN_TCH steps and every (cell, step) sum runs over the
//                   stations in the reference's order (bit-identical fields).  Only the steps with a positive reading
//                   somewhere are interpolated, in ONE pass: the record is written where it will most likely live while
//                   the per-step sum / counts that decide "wet" (:1244) are reduced; the rare dry candidate is squeezed out.
// Work: N * steps * stations * ~7 fp32 instructions (no FMA: the reference has none) against 4 B written per wet cell-step;
// the time chunk is the fastest block index so that a tile's weights are read from DRAM once.
#include "common.cuh"

namespace {

constexpr int RAIN_TCH = 16;      // time steps per thread
constexpr int RAIN_THREADS = 128;

// x**y as the reference's libm evaluates it in fp32: glibc's powf is correctly rounded in all but a vanishing share of
// cases, which a double-precision pow rounded once to float reproduces
__device__ __forceinline__ float powf_cr(float x, float y) { return (float)pow((double)x, (double)y); }

__global__ void k_idw_weights(const float2 *__restrict__ xy, const float2 *__restrict__ coord, float pp, int N, int ncoord,
                              float *__restrict__ W) {
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= N) return;
    const float2 p = xy[cell];
    for (int i = 0; i < ncoord; ++i) {
        const float2 c = __ldg(coord + i);
        const float dx = c.x - p.x, dy = c.y - p.y;
        W[(size_t)i * N + cell] = 1.0f / powf_cr(sqrtf(dx * dx + dy * dy), pp);   // :1227-1229
    }
}

struct RainP {
    int N, ncoord, T;      // T = steps this launch works on (the candidate steps), padded row length Tp of rainT
    int Tp;
    float umbral;
    const float *rainT;    // [ncoord][Tp]: the candidate steps' readings, station-major (16 consecutive steps = 4 x float4)
    const float *W;        // IDW: [ncoord][N]
    const float2 *xy, *coord;
    const int32_t *tin;    // MIT: (3,ntin) 1-based station ids
    const int32_t *perte;  // MIT: (N) 1-based triangle of each cell
    double *sum;           // [T] field sum of every candidate step
    int32_t *n_umbral, *n_pos;
    int32_t *table;        // [1 + T][N]: candidate j goes to row j + 1 (row 0 = the all-zero record 1); NULL = statistics only
    int as_float;          // hills model: keep the fp32 field (its bits) in `table`; the hills kernel sums and converts
};

template <int MODE>
__global__ void __launch_bounds__(RAIN_THREADS) k_rain_field(const RainP P) {
    // time chunk fastest: the blocks that share a tile of cells (and its weights) run next to each other, so the weights
    // are read from DRAM once instead of once per time chunk
    const int nchunks = P.Tp / RAIN_TCH;
    const int cell = (int)(blockIdx.x / nchunks) * RAIN_THREADS + threadIdx.x;
    const int t0 = (int)(blockIdx.x % nchunks) * RAIN_TCH;
    const int nt = min(RAIN_TCH, P.T - t0);
    const bool active = cell < P.N;
    float campo[RAIN_TCH];
#pragma unroll
    for (int k = 0; k < RAIN_TCH; ++k) campo[k] = 0.0f;
    if (active) {
        if (MODE == 0) {   // IDW :1234-1242
            float Wr[RAIN_TCH], den[RAIN_TCH];
#pragma unroll
            for (int k = 0; k < RAIN_TCH; ++k) { Wr[k] = 0.0f; den[k] = 0.0f; }
            const float4 *rp = reinterpret_cast<const float4 *>(P.rainT + t0);
            const size_t row4 = (size_t)P.Tp / 4;
            constexpr int SU = 4;   // stations per trip: their weights are requested together, the sums stay in station order
            for (int i0 = 0; i0 < P.ncoord; i0 += SU) {
                float w[SU];
#pragma unroll
                for (int u = 0; u < SU; ++u) w[u] = (i0 + u < P.ncoord) ? P.W[(size_t)(i0 + u) * P.N + cell] : 0.0f;
#pragma unroll
                for (int u = 0; u < SU; ++u) {
                    if (i0 + u < P.ncoord) {
                        float r[RAIN_TCH];   // the same addresses for the whole warp: four 128-bit broadcast loads
#pragma unroll
                        for (int q = 0; q < RAIN_TCH / 4; ++q) {
                            const float4 v = __ldg(rp + (size_t)(i0 + u) * row4 + q);
                            r[4 * q] = v.x; r[4 * q + 1] = v.y; r[4 * q + 2] = v.z; r[4 * q + 3] = v.w;
                        }
#pragma unroll
                        for (int k = 0; k < RAIN_TCH; ++k) {   // padding steps hold -1: they add nothing
                            if (r[k] > 0.0f) Wr[k] = Wr[k] + w[u] * r[k];
                            if (r[k] >= 0.0f) den[k] = den[k] + w[u];
                        }
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < RAIN_TCH; ++k) {
                const float q = Wr[k] / den[k];
                const float valor = (q > 0.0f) ? q : 0.0f;          // the compiled MAX(x, 0.) gives 0 for NaN
                campo[k] = (valor == valor - 1.0f) ? 0.0f : valor;  // :1237-1241 (infinity guard)
            }
        } else {           // MIT :1145-1165
            const int tri = max(__ldg(P.perte + cell) - 1, 0);   // an unassigned cell (0) is the caller's error (wmf.py:3329)
            const int ia = __ldg(P.tin + 3 * tri) - 1, ib = __ldg(P.tin + 3 * tri + 1) - 1, ic = __ldg(P.tin + 3 * tri + 2) - 1;
            const float2 a = __ldg(P.coord + ia), b = __ldg(P.coord + ib), c = __ldg(P.coord + ic), p = P.xy[cell];
            const float coef1 = (b.x - a.x) * (c.y - a.y) - (c.x - a.x) * (b.y - a.y);
            const float det1 = (c.x - a.x) * (p.y - a.y), det2 = (b.y - a.y) * (p.x - a.x);
            const float det3 = (c.y - a.y) * (p.x - a.x), det4 = (p.y - a.y) * (b.x - a.x);
            const float *ra = P.rainT + (size_t)ia * P.Tp + t0, *rb = P.rainT + (size_t)ib * P.Tp + t0, *rc = P.rainT + (size_t)ic * P.Tp + t0;
#pragma unroll
            for (int q4 = 0; q4 < RAIN_TCH / 4; ++q4) {
                const float4 va = __ldg(reinterpret_cast<const float4 *>(ra) + q4), vb = __ldg(reinterpret_cast<const float4 *>(rb) + q4),
                             vc = __ldg(reinterpret_cast<const float4 *>(rc) + q4);
                const float xa[4] = {va.x, va.y, va.z, va.w}, xb[4] = {vb.x, vb.y, vb.z, vb.w}, xc[4] = {vc.x, vc.y, vc.z, vc.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float az = fmaxf(xa[e], 0.0f), bz = fmaxf(xb[e], 0.0f), cz = fmaxf(xc[e], 0.0f);
                    const float coef2 = det1 * (bz - az) + det2 * (cz - az) - det3 * (bz - az) - det4 * (cz - az);
                    const float q = az - coef2 / coef1;
                    campo[4 * q4 + e] = (q > 0.0f) ? q : 0.0f;
                }
            }
        }
    }
    // per-step sum and counts (:1244-1246): fp64 atomics, one per warp and step
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < RAIN_TCH; ++k) {
        if (k < nt) {
            const float v = active ? campo[k] : 0.0f;
            const double s = warp_sum_d((double)v);
            const int nu = warp_sum_i((active && v > P.umbral) ? 1 : 0);
            const int np = warp_sum_i((active && v > 0.0f) ? 1 : 0);
            if (lane == 0) {
                if (s != 0.0) atomicAdd(P.sum + t0 + k, s);
                if (nu) atomicAdd(P.n_umbral + t0 + k, nu);
                if (np) atomicAdd(P.n_pos + t0 + k, np);
            }
        }
    }
    if (active && P.table) {   // campoInt = campo*1000 (truncation), one record per candidate step (:1251-1252)
#pragma unroll
        for (int k = 0; k < RAIN_TCH; ++k)
            if (k < nt) P.table[(size_t)(t0 + k + 1) * P.N + cell] = P.as_float ? __float_as_int(campo[k]) : (int32_t)(campo[k] * 1000.0f);
    }
}

// hills model (maskVector = hill of every cell): module models' basin_subbasin_map2subbasin is called with sum_mean = 2
// (modelosv2.f90:1841-1877): the hill's value is the SUM of its cells' rainfall in ascending cell order; column i of the
// record is hill n_hills - i + 1.  One thread per (hill, record).
__global__ void k_rain_hills(const int32_t *__restrict__ field, const int32_t *__restrict__ order, const int32_t *__restrict__ hstart,
                             const int32_t *__restrict__ hist, int N, int nhills, int nrec, int32_t *__restrict__ table) {
    const int h = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int j = blockIdx.y + 1;                      // record row (row 0 = zeros)
    if (h > nhills || j >= nrec) return;
    const int m = hist[h];
    const int32_t *l = order + hstart[h];
    const int32_t *f = field + (size_t)j * N;
    float s = 0.0f;
    for (int q = 0; q < m; ++q) s = s + __int_as_float(f[l[q]]);
    table[(size_t)j * nhills + (nhills - h)] = (int32_t)((m > 0 ? s : 0.0f) * 1000.0f);
}

// One pass: every step with a positive reading somewhere (a "candidate"; all other fields are identically zero, hence
// dry) is interpolated straight into the record it will most likely own.  A candidate whose field fails the wet test
// (:1244, only possible with umbral > 0) is squeezed out afterwards by moving the later records up.
template <int MODE>
int rain_run(wmf_ctx *ctx, RainP P, Scope &sc, const float *rain_dev, float pp, int32_t *table, int32_t capacity, float *mean_rain,
             int32_t *pos_ids, int32_t *n_records, const int32_t *mask = nullptr, int nhills = 0) {
    const int N = P.N, T_all = P.T, ns = P.ncoord;
    const int W = nhills > 0 ? nhills : N;             // values per record
    std::vector<float> h_rain((size_t)ns * T_all);
    CU_TRY(ctx, cudaMemcpyAsync(h_rain.data(), rain_dev, sizeof(float) * h_rain.size(), cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    std::vector<int32_t> cand;
    for (int t = 0; t < T_all; ++t) {
        bool any = P.umbral < 0.0f;   // a negative threshold makes even an all-zero field "wet"
        for (int i = 0; i < ns && !any; ++i) any = h_rain[(size_t)t * ns + i] > 0.0f;
        if (any) cand.push_back(t);
    }
    const int C = (int)cand.size();
    const int Tp = ((C + RAIN_TCH - 1) / RAIN_TCH) * RAIN_TCH;
    if (table && capacity < C + 1)
        WMF_FAIL(ctx, WMF_ERR_ARG, "rain interpolation: the table holds %d records, %d are needed while interpolating "
                 "(1 + the steps with a positive reading; nreg + 1 always suffices)", capacity, C + 1);
    std::vector<int32_t> h_pos((size_t)T_all, 1);
    std::vector<float> h_mean((size_t)T_all, 0.0f);
    int32_t cont = 2;
    int32_t *d_table = nullptr, *d_field = nullptr;
    bool table_on_host = false;
    if (table) {
        table_on_host = !Scope::on_device(table);
        if (table_on_host) WMF_TRY(sc.alloc(&d_table, (size_t)(C + 1) * W));
        else d_table = table;
        CU_TRY(ctx, cudaMemsetAsync(d_table, 0, sizeof(int32_t) * (size_t)W, ctx->stream));   // record 1 = zeros (:1224)
        if (nhills > 0) WMF_TRY(sc.alloc(&d_field, (size_t)(C + 1) * N));                     // the cell fields, fp32
    }
    if (C > 0) {
        std::vector<float> h_rT((size_t)ns * Tp, -1.0f);
        for (int j = 0; j < C; ++j)
            for (int i = 0; i < ns; ++i) h_rT[(size_t)i * Tp + j] = h_rain[(size_t)cand[j] * ns + i];
        float *d_rT;
        WMF_TRY(sc.alloc(&d_rT, h_rT.size()));
        CU_TRY(ctx, cudaMemcpyAsync(d_rT, h_rT.data(), sizeof(float) * h_rT.size(), cudaMemcpyHostToDevice, ctx->stream));
        P.rainT = d_rT; P.T = C; P.Tp = Tp;
        if (MODE == 0) {
            float *W;
            WMF_TRY(sc.alloc(&W, (size_t)ns * N));
            LAUNCH(ctx, k_idw_weights, grid_for(N, 256), 256, 0, P.xy, P.coord, pp, N, ns, W);
            P.W = W;
        }
        WMF_TRY(sc.alloc(&P.sum, (size_t)C));
        WMF_TRY(sc.alloc(&P.n_umbral, (size_t)2 * C));
        P.n_pos = P.n_umbral + C;
        CU_TRY(ctx, cudaMemsetAsync(P.sum, 0, sizeof(double) * C, ctx->stream));
        CU_TRY(ctx, cudaMemsetAsync(P.n_umbral, 0, sizeof(int32_t) * 2 * C, ctx->stream));
        P.table = nhills > 0 ? d_field : d_table;
        P.as_float = nhills > 0 ? 1 : 0;
        const long long nblocks = (long long)grid_for(N, RAIN_THREADS) * (Tp / RAIN_TCH);
        if (nblocks > 0x7fffffffLL) WMF_FAIL(ctx, WMF_ERR_ARG, "rain interpolation: %lld blocks exceed the grid limit", nblocks);
        const dim3 grid((unsigned)nblocks);
        LAUNCH(ctx, k_rain_field<MODE>, grid, RAIN_THREADS, 0, P);
        if (nhills > 0 && d_table) {
            const int32_t *d_mask;
            int32_t *order, *hstart, *hist;
            WMF_TRY(sc.in(mask, (size_t)N, &d_mask));
            WMF_TRY(wmf_hill_groups(ctx, sc, d_mask, nullptr, N, nhills, &order, &hstart, &hist));
            const dim3 hg((unsigned)grid_for(nhills, 128), (unsigned)C);
            LAUNCH(ctx, k_rain_hills, hg, 128, 0, d_field, order, hstart, hist, N, nhills, C + 1, d_table);
        }
        std::vector<double> h_sum((size_t)C);
        std::vector<int32_t> h_cnt((size_t)2 * C);
        CU_TRY(ctx, cudaMemcpyAsync(h_sum.data(), P.sum, sizeof(double) * C, cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(ctx, cudaMemcpyAsync(h_cnt.data(), P.n_umbral, sizeof(int32_t) * 2 * C, cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));   // also keeps h_rT alive until its copy is done
        for (int j = 0; j < C; ++j) {   // :1244-1268
            if ((float)h_sum[j] > P.umbral && h_cnt[j] > 0) {
                h_mean[cand[j]] = (float)(h_sum[j] / (double)h_cnt[(size_t)C + j]);
                if (d_table && cont != j + 2)   // a dry candidate came before: move this record up to its final row
                    CU_TRY(ctx, cudaMemcpyAsync(d_table + (size_t)(cont - 1) * W, d_table + (size_t)(j + 1) * W, sizeof(int32_t) * (size_t)W,
                                                cudaMemcpyDeviceToDevice, ctx->stream));
                h_pos[cand[j]] = cont++;
            }
        }
    }
    *n_records = cont - 1;
    float *d_mean;
    int32_t *d_pos;
    WMF_TRY(sc.out(mean_rain, (size_t)T_all, &d_mean));
    WMF_TRY(sc.out(pos_ids, (size_t)T_all, &d_pos));
    if (d_mean) CU_TRY(ctx, cudaMemcpyAsync(d_mean, h_mean.data(), sizeof(float) * T_all, cudaMemcpyHostToDevice, ctx->stream));
    if (d_pos) CU_TRY(ctx, cudaMemcpyAsync(d_pos, h_pos.data(), sizeof(int32_t) * T_all, cudaMemcpyHostToDevice, ctx->stream));
    if (table_on_host)
        CU_TRY(ctx, cudaMemcpyAsync(table, d_table, sizeof(int32_t) * (size_t)*n_records * W, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));   // the host vectors above go out of scope
    return sc.finish();
}

int rain_stage(wmf_ctx *ctx, Scope &sc, RainP &P, const float **rain_dev, const float *xy_basin, const float *coord, const float *rain,
               int32_t nceldas, int32_t ncoord, int32_t nreg, float umbral) {
    if (!xy_basin || !coord || !rain || nceldas <= 0 || ncoord <= 0 || nreg <= 0)
        WMF_FAIL(ctx, WMF_ERR_ARG, "rain interpolation: bad sizes or null inputs");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    P.N = nceldas; P.ncoord = ncoord; P.T = nreg; P.umbral = umbral;
    const float *d_xy, *d_co;
    WMF_TRY(sc.in(xy_basin, (size_t)2 * nceldas, &d_xy));
    WMF_TRY(sc.in(coord, (size_t)2 * ncoord, &d_co));
    WMF_TRY(sc.in(rain, (size_t)ncoord * nreg, rain_dev));
    if ((reinterpret_cast<uintptr_t>(d_xy) | reinterpret_cast<uintptr_t>(d_co)) & 7)
        WMF_FAIL(ctx, WMF_ERR_ARG, "rain interpolation: xy_basin / coord must be 8-byte aligned");
    P.xy = reinterpret_cast<const float2 *>(d_xy);
    P.coord = reinterpret_cast<const float2 *>(d_co);
    return WMF_OK;
}

// rain_pre_mit, modelosv2.f90:1068-1101: one thread per cell scans the triangles; the last one with s > 0 and t > 0 wins
__global__ void k_pre_mit(const float2 *__restrict__ xy, const int32_t *__restrict__ tin, const float2 *__restrict__ coord, int N,
                          int ntin, int32_t *__restrict__ perte) {
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= N) return;
    const float2 p = xy[cell];
    int own = 0;
    for (int k = 0; k < ntin; ++k) {
        const float2 p0 = __ldg(coord + __ldg(tin + 3 * k) - 1), p1 = __ldg(coord + __ldg(tin + 3 * k + 1) - 1),
                     p2 = __ldg(coord + __ldg(tin + 3 * k + 2) - 1);
        const float Area = fabsf(1.0f / 2.0f * (-p1.y * p2.x + p0.y * (-p1.x + p2.x) + p0.x * (p1.y - p2.y) + p1.x * p2.y));
        const float inv = 1.0f / (2.0f * Area);
        const float s = inv * (p0.y * p2.x - p0.x * p2.y + (p2.y - p0.y) * p.x + (p0.x - p2.x) * p.y);
        const float t = inv * (p0.x * p1.y - p0.y * p1.x + (p0.y - p1.y) * p.x + (p1.x - p0.x) * p.y);
        if (s > 0.0f && t > 0.0f) own = k + 1;
    }
    perte[cell] = own;
}

}  // namespace

extern "C" {

int wmf_rain_pre_mit(wmf_ctx *ctx, const float *xy_basin, const int32_t *tin, const float *coord, int32_t nceldas, int32_t ntin,
                     int32_t ncoord, int32_t *tin_perte) {
    if (!ctx || !xy_basin || !tin || !coord || !tin_perte || nceldas <= 0 || ntin <= 0 || ncoord <= 0) return WMF_ERR_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    Scope sc(ctx);
    const float *d_xy, *d_co;
    const int32_t *d_t;
    int32_t *d_p;
    WMF_TRY(sc.in(xy_basin, (size_t)2 * nceldas, &d_xy));
    WMF_TRY(sc.in(coord, (size_t)2 * ncoord, &d_co));
    WMF_TRY(sc.in(tin, (size_t)3 * ntin, &d_t));
    WMF_TRY(sc.out(tin_perte, (size_t)nceldas, &d_p));
    LAUNCH(ctx, k_pre_mit, grid_for(nceldas, 256), 256, 0, reinterpret_cast<const float2 *>(d_xy), d_t,
           reinterpret_cast<const float2 *>(d_co), nceldas, ntin, d_p);
    return sc.finish();
}


int wmf_rain_idw(wmf_ctx *ctx, const float *xy_basin, const float *coord, const float *rain, float pp, int32_t nceldas,
                 int32_t ncoord, int32_t nreg, float umbral, int32_t *table, int32_t capacity, float *mean_rain,
                 int32_t *pos_ids, int32_t *n_records, const int32_t *mask_vector, int32_t nhills) {
    if (!ctx || !n_records || nhills < 0 || (nhills > 0 && !mask_vector)) return WMF_ERR_ARG;
    Scope sc(ctx);
    RainP P{};
    const float *d_rain;
    WMF_TRY(rain_stage(ctx, sc, P, &d_rain, xy_basin, coord, rain, nceldas, ncoord, nreg, umbral));
    return rain_run<0>(ctx, P, sc, d_rain, pp, table, capacity, mean_rain, pos_ids, n_records, mask_vector, nhills);
}

int wmf_rain_mit(wmf_ctx *ctx, const float *xy_basin, const float *coord, const float *rain, const int32_t *tin,
                 const int32_t *tin_perte, int32_t nceldas, int32_t ncoord, int32_t ntin, int32_t nreg, float umbral,
                 int32_t *table, int32_t capacity, float *mean_rain, int32_t *pos_ids, int32_t *n_records) {
    if (!ctx || !n_records || !tin || !tin_perte || ntin <= 0) return WMF_ERR_ARG;
    Scope sc(ctx);
    RainP P{};
    const float *d_rain;
    WMF_TRY(rain_stage(ctx, sc, P, &d_rain, xy_basin, coord, rain, nceldas, ncoord, nreg, umbral));
    WMF_TRY(sc.in(tin, (size_t)3 * ntin, &P.tin));
    WMF_TRY(sc.in(tin_perte, (size_t)nceldas, &P.perte));
    return rain_run<1>(ctx, P, sc, d_rain, 0.0f, table, capacity, mean_rain, pos_ids, n_records);
}

}  // extern "C"
