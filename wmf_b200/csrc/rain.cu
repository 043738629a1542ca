// rain.cu -- rainfall interpolation on the GPU: rain_idw (modelosv2.f90:1195-1271) and rain_mit (:1104-1194), cells
// model.  The reference interpolates one field per time step on the host, writes the wet ones to a direct-access
// file and shia_v1 reads them back record by record; here the record table is produced in HBM in the layout
// wmf_shia_run consumes (record 1 = zeros, record k = k-th wet step, int32 mm*1000 truncated), so a run can go from
// station series to hydrograph without the table ever crossing PCIe.
//
//   k_idw_weights   W(i,cell) = 1 / sqrt(dx^2+dy^2)**pp, station-major so that a warp reads 32 consecutive cells
//   k_rain_field    one thread per cell, RAIN_TCH time steps per thread in registers; the station loop is the outer
//                   loop, so each weight is loaded once per RAIN_TCH steps and every (cell, step) sum runs over the
//                   stations in the reference's order (bit-identical fields).  Pass 0 reduces sum / counts per step
//                   (which steps are "wet" depends on them, :1244), pass 1 writes the wet records.
// HBM roofline: pass 1 writes 4 B per wet cell-step and re-reads 4*ncoord B of weights per cell per RAIN_TCH steps.
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
    int N, ncoord, T;
    float umbral;
    const float *rain;     // [T][ncoord]
    const float *W;        // IDW: [ncoord][N]
    const float2 *xy, *coord;
    const int32_t *tin;    // MIT: (3,ntin) 1-based station ids
    const int32_t *perte;  // MIT: (N) 1-based triangle of each cell
    double *sum;           // [T] pass 0
    int32_t *n_umbral, *n_pos;
    const int32_t *slot;   // [T] pass 1: record (1-based) that receives step t, 1 = dry
    int32_t *table;        // [n_records][N]
};

template <int MODE, int PASS>
__global__ void __launch_bounds__(RAIN_THREADS) k_rain_field(const RainP P) {
    const int cell = blockIdx.x * RAIN_THREADS + threadIdx.x;
    const int t0 = blockIdx.y * RAIN_TCH;
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
            const float *rp = P.rain + (size_t)t0 * P.ncoord;
            for (int i = 0; i < P.ncoord; ++i) {
                const float w = P.W[(size_t)i * P.N + cell];
#pragma unroll
                for (int k = 0; k < RAIN_TCH; ++k) {
                    if (k < nt) {
                        const float r = __ldg(rp + (size_t)k * P.ncoord + i);   // same address for the whole warp
                        if (r > 0.0f) Wr[k] = Wr[k] + w * r;
                        if (r >= 0.0f) den[k] = den[k] + w;
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
#pragma unroll
            for (int k = 0; k < RAIN_TCH; ++k) {
                if (k < nt) {
                    const float *r = P.rain + (size_t)(t0 + k) * P.ncoord;
                    const float az = fmaxf(__ldg(r + ia), 0.0f), bz = fmaxf(__ldg(r + ib), 0.0f), cz = fmaxf(__ldg(r + ic), 0.0f);
                    const float coef2 = det1 * (bz - az) + det2 * (cz - az) - det3 * (bz - az) - det4 * (cz - az);
                    const float q = az - coef2 / coef1;
                    campo[k] = (q > 0.0f) ? q : 0.0f;
                }
            }
        }
    }
    if (PASS == 0) {   // per-step sum and counts (:1244-1246): fp64 atomics, one per warp and step
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
    } else if (active) {   // campoInt = campo*1000 (truncation), record `cont` of the file (:1251-1252)
#pragma unroll
        for (int k = 0; k < RAIN_TCH; ++k) {
            if (k < nt) {
                const int s = __ldg(P.slot + t0 + k);
                if (s > 1) P.table[(size_t)(s - 1) * P.N + cell] = (int32_t)(campo[k] * 1000.0f);
            }
        }
    }
}

template <int MODE>
int rain_run(wmf_ctx *ctx, RainP P, Scope &sc, float pp, int32_t *table, int32_t capacity, float *mean_rain, int32_t *pos_ids,
             int32_t *n_records) {
    const int N = P.N, T = P.T;
    if (MODE == 0) {
        float *W;
        WMF_TRY(sc.alloc(&W, (size_t)P.ncoord * N));
        LAUNCH(ctx, k_idw_weights, grid_for(N, 256), 256, 0, P.xy, P.coord, pp, N, P.ncoord, W);
        P.W = W;
    }
    WMF_TRY(sc.alloc(&P.sum, (size_t)T));
    WMF_TRY(sc.alloc(&P.n_umbral, (size_t)2 * T));
    P.n_pos = P.n_umbral + T;
    CU_TRY(ctx, cudaMemsetAsync(P.sum, 0, sizeof(double) * T, ctx->stream));
    CU_TRY(ctx, cudaMemsetAsync(P.n_umbral, 0, sizeof(int32_t) * 2 * T, ctx->stream));
    const dim3 grid((unsigned)grid_for(N, RAIN_THREADS), (unsigned)grid_for(T, RAIN_TCH));
    auto k_pass0 = k_rain_field<MODE, 0>;
    auto k_pass1 = k_rain_field<MODE, 1>;
    LAUNCH(ctx, k_pass0, grid, RAIN_THREADS, 0, P);
    std::vector<double> h_sum((size_t)T);
    std::vector<int32_t> h_cnt((size_t)2 * T), h_slot((size_t)T);
    CU_TRY(ctx, cudaMemcpyAsync(h_sum.data(), P.sum, sizeof(double) * T, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaMemcpyAsync(h_cnt.data(), P.n_umbral, sizeof(int32_t) * 2 * T, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    int32_t cont = 2;
    std::vector<float> h_mean((size_t)T);
    for (int t = 0; t < T; ++t) {   // :1244-1268
        if ((float)h_sum[t] > P.umbral && h_cnt[t] > 0) {
            h_mean[t] = (float)(h_sum[t] / (double)h_cnt[(size_t)T + t]);
            h_slot[t] = cont++;
        } else {
            h_mean[t] = 0.0f;
            h_slot[t] = 1;
        }
    }
    *n_records = cont - 1;
    float *d_mean;
    int32_t *d_pos, *d_slot;
    WMF_TRY(sc.alloc(&d_slot, (size_t)T));
    CU_TRY(ctx, cudaMemcpyAsync(d_slot, h_slot.data(), sizeof(int32_t) * T, cudaMemcpyHostToDevice, ctx->stream));
    WMF_TRY(sc.out(mean_rain, (size_t)T, &d_mean));
    WMF_TRY(sc.out(pos_ids, (size_t)T, &d_pos));
    if (d_mean) CU_TRY(ctx, cudaMemcpyAsync(d_mean, h_mean.data(), sizeof(float) * T, cudaMemcpyHostToDevice, ctx->stream));
    if (d_pos) CU_TRY(ctx, cudaMemcpyAsync(d_pos, d_slot, sizeof(int32_t) * T, cudaMemcpyDeviceToDevice, ctx->stream));
    if (table) {
        if (capacity < *n_records)
            WMF_FAIL(ctx, WMF_ERR_ARG, "rain interpolation: the table holds %d records, %d are needed", capacity, *n_records);
        WMF_TRY(sc.out(table, (size_t)*n_records * N, &P.table));
        CU_TRY(ctx, cudaMemsetAsync(P.table, 0, sizeof(int32_t) * (size_t)N, ctx->stream));   // record 1 = zeros (:1224)
        P.slot = d_slot;
        LAUNCH(ctx, k_pass1, grid, RAIN_THREADS, 0, P);
    }
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));   // the host vectors above go out of scope
    return sc.finish();
}

int rain_stage(wmf_ctx *ctx, Scope &sc, RainP &P, const float *xy_basin, const float *coord, const float *rain, int32_t nceldas,
               int32_t ncoord, int32_t nreg, float umbral) {
    if (!xy_basin || !coord || !rain || nceldas <= 0 || ncoord <= 0 || nreg <= 0)
        WMF_FAIL(ctx, WMF_ERR_ARG, "rain interpolation: bad sizes or null inputs");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    P.N = nceldas; P.ncoord = ncoord; P.T = nreg; P.umbral = umbral;
    const float *d_xy, *d_co;
    WMF_TRY(sc.in(xy_basin, (size_t)2 * nceldas, &d_xy));
    WMF_TRY(sc.in(coord, (size_t)2 * ncoord, &d_co));
    WMF_TRY(sc.in(rain, (size_t)ncoord * nreg, &P.rain));
    if ((reinterpret_cast<uintptr_t>(d_xy) | reinterpret_cast<uintptr_t>(d_co)) & 7)
        WMF_FAIL(ctx, WMF_ERR_ARG, "rain interpolation: xy_basin / coord must be 8-byte aligned");
    P.xy = reinterpret_cast<const float2 *>(d_xy);
    P.coord = reinterpret_cast<const float2 *>(d_co);
    return WMF_OK;
}

}  // namespace

extern "C" {

int wmf_rain_idw(wmf_ctx *ctx, const float *xy_basin, const float *coord, const float *rain, float pp, int32_t nceldas,
                 int32_t ncoord, int32_t nreg, float umbral, int32_t *table, int32_t capacity, float *mean_rain,
                 int32_t *pos_ids, int32_t *n_records) {
    if (!ctx || !n_records) return WMF_ERR_ARG;
    Scope sc(ctx);
    RainP P{};
    WMF_TRY(rain_stage(ctx, sc, P, xy_basin, coord, rain, nceldas, ncoord, nreg, umbral));
    return rain_run<0>(ctx, P, sc, pp, table, capacity, mean_rain, pos_ids, n_records);
}

int wmf_rain_mit(wmf_ctx *ctx, const float *xy_basin, const float *coord, const float *rain, const int32_t *tin,
                 const int32_t *tin_perte, int32_t nceldas, int32_t ncoord, int32_t ntin, int32_t nreg, float umbral,
                 int32_t *table, int32_t capacity, float *mean_rain, int32_t *pos_ids, int32_t *n_records) {
    if (!ctx || !n_records || !tin || !tin_perte || ntin <= 0) return WMF_ERR_ARG;
    Scope sc(ctx);
    RainP P{};
    WMF_TRY(rain_stage(ctx, sc, P, xy_basin, coord, rain, nceldas, ncoord, nreg, umbral));
    WMF_TRY(sc.in(tin, (size_t)3 * ntin, &P.tin));
    WMF_TRY(sc.in(tin_perte, (size_t)nceldas, &P.perte));
    return rain_run<1>(ctx, P, sc, 0.0f, table, capacity, mean_rain, pos_ids, n_records);
}

}  // extern "C"
