// ctx.cu -- context lifetime, module-`cu` globals and small host-side entry points.
#include <math.h>

#include "common.cuh"

extern "C" {

int wmf_abi_version(void) { return WMF_ABI_VERSION; }

int wmf_ctx_create(int device, wmf_ctx **out) {
    if (!out) return WMF_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return WMF_ERR_CUDA;  // no CPU fallback
    if (device < 0 || device >= ndev) return WMF_ERR_ARG;
    wmf_ctx *c = new wmf_ctx();
    c->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete c; return WMF_ERR_CUDA; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete c; return WMF_ERR_CUDA; }
    c->num_sms = prop.multiProcessorCount;
    // a *blocking* stream: it is implicitly ordered against the legacy default stream, which is where torch puts
    // the tensors that callers hand us as device pointers
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamDefault) != cudaSuccess) { delete c; return WMF_ERR_CUDA; }
    c->own_stream = true;
    if (cudaDeviceGetDefaultMemPool(&c->pool, device) != cudaSuccess) { delete c; return WMF_ERR_CUDA; }
    uint64_t keep = UINT64_MAX;  // keep freed blocks cached: steady-state calls never hit the driver
    cudaMemPoolSetAttribute(c->pool, cudaMemPoolAttrReleaseThreshold, &keep);
    cudaEventCreate(&c->ev0);
    cudaEventCreate(&c->ev1);
    *out = c;
    return WMF_OK;
}

int wmf_ctx_destroy(wmf_ctx *ctx) {
    if (!ctx) return WMF_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->q_lin) cudaFree(ctx->q_lin);
    if (ctx->q_par) cudaFree(ctx->q_par);
    if (ctx->q_acum) cudaFree(ctx->q_acum);
    ctx->plan.release();
    if (ctx->sub_basins) cudaFree(ctx->sub_basins);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    for (cudaEvent_t e : ctx->copy_events) cudaEventDestroy(e);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->sed_stream) cudaStreamDestroy(ctx->sed_stream);
    for (int i = 0; i < 4; ++i) {
        if (ctx->sed_ev_h[i]) cudaEventDestroy(ctx->sed_ev_h[i]);
        if (ctx->sed_ev_s[i]) cudaEventDestroy(ctx->sed_ev_s[i]);
    }
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return WMF_OK;
}

const char *wmf_last_error(const wmf_ctx *ctx) { return ctx ? ctx->err : "null context"; }

int wmf_set_stream(wmf_ctx *ctx, void *cuda_stream) {
    if (!ctx) return WMF_ERR_ARG;
    cudaStreamSynchronize(ctx->stream);
    if (cuda_stream == nullptr) {
        if (!ctx->own_stream) {
            CU_TRY(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamDefault));
            ctx->own_stream = true;
        }
        return WMF_OK;
    }
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->own_stream = false;
    return WMF_OK;
}

int wmf_synchronize(wmf_ctx *ctx) {
    if (!ctx) return WMF_ERR_ARG;
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return WMF_OK;
}

int64_t wmf_launch_count(const wmf_ctx *ctx) { return ctx ? ctx->launches : 0; }

int wmf_set_geo(wmf_ctx *ctx, int32_t ncols, int32_t nrows, float xll, float yll, float dx, float dy, float dxp,
                float nodata) {
    if (!ctx) return WMF_ERR_ARG;
    ctx->geo.ncols = ncols;
    ctx->geo.nrows = nrows;
    ctx->geo.xll = xll;
    ctx->geo.yll = yll;
    ctx->geo.dx = dx;
    ctx->geo.dy = dy;
    ctx->geo.dxp = dxp;
    ctx->geo.nodata = nodata;
    return WMF_OK;
}

int wmf_get_basin_scalars(wmf_ctx *ctx, float out5[5]) {
    if (!ctx || !out5) return WMF_ERR_ARG;
    memcpy(out5, ctx->scalars, sizeof(float) * 5);
    return WMF_OK;
}

// cuencas.f90:77-90 -- two fp32 divisions on the host; `volatile` keeps gcc from contracting/widening.
int wmf_coord2fil_col(wmf_ctx *ctx, float x, float y, int32_t *col, int32_t *fil) {
    if (!ctx || !col || !fil) return WMF_ERR_ARG;
    const wmf_geo &g = ctx->geo;
    volatile float qx = (x - g.xll) / g.dx;
    volatile float qy = (y - g.yll) / g.dx;
    int32_t c = (int32_t)ceilf(qx);
    if (c < 0) c = -c;
    int32_t f = g.nrows - (int32_t)floorf(qy);
    if (c > g.ncols || c <= 0) c = -999;
    if (f > g.nrows || f <= 0) f = -999;
    *col = c;
    *fil = f;
    return WMF_OK;
}

int wmf_last_kernel_ms(wmf_ctx *ctx, float *ms) {
    if (!ctx || !ms) return WMF_ERR_ARG;
    *ms = ctx->last_ms;
    return WMF_OK;
}

int wmf_last_shia_stats(wmf_ctx *ctx, int32_t *n_levels, int32_t *n_waves) {
    if (!ctx) return WMF_ERR_ARG;
    if (n_levels) *n_levels = ctx->last_levels;
    if (n_waves) *n_waves = ctx->last_waves;
    return WMF_OK;
}

}  // extern "C"
