// common.cuh -- internal helpers of libwmf_b200.so: context, host<->device staging, error plumbing,
// small device-wide primitives.  Compiled for sm_100a only; no CPU fallback anywhere.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <vector>

#include "../../include/wmf_b200.h"

struct wmf_geo {
    int32_t ncols = 0, nrows = 0;
    float xll = 0.f, yll = 0.f, dx = 1.f, dy = 1.f, dxp = 1.f, nodata = -9999.f;
};

// The plan of a SHIA sweep (shia.cu: shia_build_plan), kept between runs of the same basin
struct wmf_shia_plan {
    bool valid = false;
    int N = 0;
    unsigned long long hash = 0;   // checksum of drain ids, unit types, control points
    int32_t Lmax = 0;
    int64_t nq = 0, nh = 0, n_hill = 0;
    uint32_t *tw = nullptr;
    int32_t *cs = nullptr, *hidx = nullptr, *cidx = nullptr, *ctrl_cells = nullptr, *ctrlh_cells = nullptr;
    std::vector<int32_t> h_first, h_hfirst;
    void release() {
        void *p[] = {tw, cs, hidx, cidx, ctrl_cells, ctrlh_cells};
        for (void *q : p)
            if (q) cudaFree(q);
        tw = nullptr; cs = hidx = cidx = ctrl_cells = ctrlh_cells = nullptr;
        valid = false;
    }
};

struct wmf_ctx {
    int device = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    cudaStream_t copy_stream = nullptr;          // host->device streaming of big inputs next to the compute stream
    std::vector<cudaEvent_t> copy_events;
    // sediment kernels next to the hydrology kernels of the following wave (shia.cu, launch_tb): second stream + event rings
    cudaStream_t sed_stream = nullptr;
    cudaEvent_t sed_ev_h[4] = {nullptr, nullptr, nullptr, nullptr}, sed_ev_s[4] = {nullptr, nullptr, nullptr, nullptr};
    bool sed_two_streams = false;   // this run
    long long sed_seq = 0;          // waves with a sediment kernel launched so far in this run
    cudaMemPool_t pool = nullptr;
    char err[512] = {0};
    int64_t launches = 0;
    wmf_geo geo;
    float scalars[5] = {0, 0, 0, 0, 0};  // area, pend_media, elevacion, centrox, centroy
    // basin_find state (the reference's module global basin_temp)
    int32_t *q_lin = nullptr;  // BFS queue: linear raster index (row*nc+col, 0-based) per queue position
    int32_t *q_par = nullptr;  // queue position (1-based) of the cell it drains to; 0 for the outlet
    int32_t *q_acum = nullptr; // cells draining through each queue position (tile path only; what basin_acum returns for this table)
    bool q_acum_valid = false;
    int64_t q_cap = 0;
    int32_t find_n = 0, find_nc = 0, find_nr = 0, find_depth = 0;
    bool find_outside = false;   // the last basin_find's outlet lay outside the map (one-cell pseudo basin, cuencas.f90:547)
    // basin_subbasin_nod state (the reference's module global sub_basins_temp(2,n_nodos), returned by basin_subbasin_cut)
    int32_t *sub_basins = nullptr;
    int32_t sub_n = 0;
    wmf_shia_plan plan;
    // timing of the last shia run
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float last_ms = 0.f;
    int32_t last_levels = 0, last_waves = 0;
};

#define WMF_FAIL(ctx, code, ...)                                  \
    do {                                                          \
        snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__);    \
        return (code);                                            \
    } while (0)

#define CU_TRY(ctx, expr)                                                                           \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,  \
                     cudaGetErrorString(e__));                                                      \
            return WMF_ERR_CUDA;                                                                    \
        }                                                                                           \
    } while (0)

#define WMF_TRY(expr)              \
    do {                           \
        int rc__ = (expr);         \
        if (rc__ != WMF_OK) return rc__; \
    } while (0)

// Arena of temporaries + deferred host copies for one API call.  Device buffers come from the context's
// stream-ordered pool, so after warm-up a call allocates without touching the driver.
struct Scope {
    wmf_ctx *ctx;
    std::vector<void *> temps;
    struct Out { void *host; void *dev; size_t bytes; };
    std::vector<Out> outs;
    explicit Scope(wmf_ctx *c) : ctx(c) {}
    ~Scope() {
        for (void *p : temps) cudaFreeAsync(p, ctx->stream);
    }
    int alloc(void **p, size_t bytes) {
        if (bytes == 0) bytes = 4;
        cudaError_t e = cudaMallocAsync(p, bytes, ctx->pool, ctx->stream);
        if (e != cudaSuccess) {
            snprintf(ctx->err, sizeof(ctx->err), "cudaMallocAsync(%zu bytes): %s", bytes, cudaGetErrorString(e));
            return e == cudaErrorMemoryAllocation ? WMF_ERR_NOMEM : WMF_ERR_CUDA;
        }
        temps.push_back(*p);
        return WMF_OK;
    }
    template <typename T>
    int alloc(T **p, size_t count) { return alloc((void **)p, count * sizeof(T)); }
    static bool on_device(const void *p) {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
    }
    // input: device pointer to `count` Ts, copying from the host if needed
    template <typename T>
    int in(const T *src, size_t count, const T **dev) {
        if (!src) { *dev = nullptr; return WMF_OK; }
        if (on_device(src)) { *dev = src; return WMF_OK; }
        T *d;
        WMF_TRY(alloc(&d, count));
        CU_TRY(ctx, cudaMemcpyAsync(d, src, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
        *dev = d;
        return WMF_OK;
    }
    // output: device pointer to write `count` Ts to; host destinations are copied back in finish()
    template <typename T>
    int out(T *dst, size_t count, T **dev) {
        if (!dst) { *dev = nullptr; return WMF_OK; }
        if (on_device(dst)) { *dev = dst; return WMF_OK; }
        T *d;
        WMF_TRY(alloc(&d, count));
        outs.push_back({(void *)dst, (void *)d, count * sizeof(T)});
        *dev = d;
        return WMF_OK;
    }
    // copy host outputs back and wait: when a call returns, its results are complete wherever they live
    int finish() {
        for (auto &o : outs)
            CU_TRY(ctx, cudaMemcpyAsync(o.host, o.dev, o.bytes, cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(ctx, cudaGetLastError());
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        outs.clear();
        return WMF_OK;
    }
};

static inline int grid_for(int64_t n, int block) { return (int)((n + block - 1) / block); }

#define LAUNCH(ctx, kern, grid, block, smem, ...)                        \
    do {                                                                 \
        kern<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);   \
        (ctx)->launches++;                                               \
    } while (0)

// ---- device helpers -------------------------------------------------------------------------------
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// inclusive warp scan
__device__ __forceinline__ int warp_scan_incl(int v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}
// exclusive block scan of one int per thread (blockDim.x <= 1024, multiple of 32); returns exclusive prefix and
// writes the block total to *total.  `sm` must hold 33 ints.  Contains two __syncthreads().
__device__ __forceinline__ int block_scan_excl(int v, int *sm, int *total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int incl = warp_scan_incl(v);
    if (lane == 31) sm[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int w = (lane < nw) ? sm[lane] : 0;
        int wi = warp_scan_incl(w);
        sm[lane] = wi - w;
        if (lane == 31) sm[32] = wi;
    }
    __syncthreads();
    int res = sm[wid] + incl - v;
    *total = sm[32];
    __syncthreads();
    return res;
}

// entry points implemented in the other translation units
int wmf_device_exclusive_scan_i32(wmf_ctx *ctx, Scope &sc, const int32_t *in, int32_t *out, int64_t n,
                                  int64_t *total_host /* may be null */);

// basin_find without a level-by-level search (pathfind.cu): fills ctx->q_lin / q_par / q_acum
int wmf_pathfind(wmf_ctx *ctx, Scope &sc, const int32_t *d_dir, int nc, int nr, int lin0, int32_t *n_out, int32_t *depth_out);
// basin_acum of the table the last basin_find produced (verified drain id by drain id); *used = 1 when it applied
int wmf_pathfind_acum_cached(wmf_ctx *ctx, Scope &sc, const int32_t *d_basin_f, int n, int32_t *d_acum, int *used);

// cells gathered per hill in ascending cell order (subbasin.cu): order[hstart[h] .. +hist[h]) = the cells whose sub_pert is h
// (and whose cauce is 1 when cauce is given), h = 1..n_hills
int wmf_hill_groups(wmf_ctx *ctx, Scope &sc, const int32_t *sub_pert, const int32_t *cauce /* may be null */, int n, int n_hills,
                    int32_t **order, int32_t **hstart, int32_t **hist);
