// topo.cu -- Path A: D8 flow-direction topology (module `cu`, wmf/cuencas.f90) on sm_100a.
//
//  basin_find   BFS in ONE persistent cooperative kernel.  Queue order == the Fortran's FIFO order
//               (cuencas.f90:552-589): children of queue entry q are appended in the 3x3 scan order kf=1..3,
//               kc=1..3.  Each CTA owns a contiguous share of the frontier and expands it for 32 levels on its
//               own (no communication); a [level][CTA] table of child counts then gives every segment its queue
//               position.  Two grid-wide exchanges per 32 levels instead of barriers every level.
//  basin_acum   leaf-to-root "last arriver climbs" sweep: one 64-bit atomicAdd per tree edge carries both the
//               partial cell count and the arrival counter, no global barriers, exact (integer sums commute).
//  acum_var /   same sweep, but the last arriver *gathers* its children in ascending cell order, which
//  propagate    reproduces the Fortran's fp32 addition order bit for bit.
//  the rest     one thread per cell / raster cell.
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

namespace cg = cooperative_groups;

// ---------------------------------------------------------------------------------------------------
// basin_find
// ---------------------------------------------------------------------------------------------------
struct FindState {
    long long n;   // cells found
    int depth;     // BFS levels (tree depth + 1)
    int err;       // 1 = queue overflow (cyclic DIR)
};

#ifndef FIND_THREADS_N
#define FIND_THREADS_N 256
#endif
constexpr int FIND_THREADS = FIND_THREADS_N;
constexpr int FIND_CAP = 16384;       // cells of a CTA's share of a level kept in shared memory (lin, next lin, mask)
constexpr size_t FIND_SMEM = (size_t)FIND_CAP * 9;
constexpr int FIND_MAX_CTAS = 256;

// 8-bit mask of the neighbours of raster cell `lin` that drain into it, bit order = Fortran scan order
// (kf outer, kc inner, centre skipped): cuencas.f90:555-571.
__device__ __forceinline__ unsigned probe_children(const int32_t *__restrict__ dir, int nc, int nr, float rnc, int lin) {
    // row = lin / nc without the integer-division sequence: float estimate, one correction step (|error| <= 1)
    int f0 = (int)((float)lin * rnc);
    int c0 = lin - f0 * nc;
    if (c0 < 0) { c0 += nc; --f0; }
    else if (c0 >= nc) { c0 -= nc; ++f0; }
    unsigned m = 0;
    int bit = 0;
#pragma unroll
    for (int kf = 1; kf <= 3; ++kf) {
#pragma unroll
        for (int kc = 1; kc <= 3; ++kc) {
            if (kf == 2 && kc == 2) continue;
            const int c = c0 + kc - 2, f = f0 + kf - 2;
            if (c >= 0 && c < nc && f >= 0 && f < nr) {
                if (__ldg(dir + (size_t)f * nc + c) == 3 * kf - kc + 1) m |= 1u << bit;
            }
            ++bit;
        }
    }
    return m;
}

// Appends the children of `lin` at queue position `pos`; the first `cap - (pos - own_lo)` of them are also kept in
// shared memory (`s_nxt`, the CTA's share of the next level) and the DIR rows they will probe are prefetched.
__device__ __forceinline__ void write_children(unsigned m, int lin, int nc, long long pos, int parent,
                                               int32_t *__restrict__ q_lin, int32_t *__restrict__ q_par,
                                               int32_t *__restrict__ s_nxt, long long own_lo, int cap) {
    int bit = 0;
#pragma unroll
    for (int kf = 1; kf <= 3; ++kf) {
#pragma unroll
        for (int kc = 1; kc <= 3; ++kc) {
            if (kf == 2 && kc == 2) continue;
            if (m & (1u << bit)) {
                const int cl = lin + (kf - 2) * nc + (kc - 2);
                q_lin[pos] = cl;
                q_par[pos] = parent;
                if (pos - own_lo < cap) s_nxt[pos - own_lo] = cl;
                ++pos;
            }
            ++bit;
        }
    }
}

__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// All-to-all exchange of one 32-bit value per CTA = the only grid-wide synchronisation of a BFS level.
// CTA b publishes (stamp << 32 | value) in its slot; thread j of every CTA spins on slot j until it carries the
// stamp.  Returns the exclusive prefix of the values over the CTAs before this one and the total.
// Two slot sets alternate (a CTA can run at most one exchange ahead of the slowest one).  All CTAs are co-resident
// (cooperative launch), so the spin cannot deadlock.  The slot traffic is relaxed; `fence` adds the release / acquire
// fences that make the CTAs' earlier global writes visible to each other (needed only when they will read them).
template <bool FENCE>
__device__ __forceinline__ void cta_exchange(unsigned long long *slots, unsigned stamp, int value, int *s_scan,
                                             int *my_off, int *total, int *lopsided) {
    const int tid = threadIdx.x, b = blockIdx.x, nb = gridDim.x;
    unsigned long long *set = slots + (size_t)(stamp & 1u) * FIND_MAX_CTAS;
    __syncthreads();  // every thread's global writes of this phase precede the publication
    if (tid == 0) {
        if (FENCE) __threadfence();
        st_relaxed_u64(set + b, ((unsigned long long)stamp << 32) | (unsigned)value);
    }
    int v = 0;
    if (tid < nb) {
        unsigned long long x;
        do { x = ld_relaxed_u64(set + tid); } while ((unsigned)(x >> 32) != stamp);
        v = (int)(unsigned)x;
        if (FENCE) __threadfence();
    }
    int tot;
    const int ex = block_scan_excl(v, s_scan, &tot);  // (contains __syncthreads)
    if (tid == b) s_scan[33] = ex;
    // some CTA holds much more than an even share of the total?  (same answer in every CTA)
    *lopsided = __syncthreads_or(tid < nb && v > 2 * ((tot + nb - 1) / nb) + FIND_THREADS);
    *my_off = s_scan[33];
    *total = tot;
    __syncthreads();
}

// Breadth-first search in epochs of FIND_EPOCH levels with NO communication inside an epoch.
// At an epoch's start the numbered frontier (a contiguous range of queue positions) is split evenly among the CTAs,
// in queue order.  Every CTA then expands its own share level after level on its own: it probes its cells, scans the
// child counts locally and appends the children (raster index + index of the parent inside the previous level's
// segment) to a private region, which is also its next frontier.  Because the shares are ordered by CTA and every
// CTA keeps the Fortran's scan order inside its share, the queue order of level L (cuencas.f90:552-589) is "CTA 0's
// segment of L, CTA 1's segment of L, ...": after the epoch one table of child counts [level][CTA] gives every
// segment its final queue position, the CTAs copy their segments out (coalesced) and translate parent indices, and
// the last level becomes the next epoch's frontier.  Two fenced grid-wide exchanges per epoch instead of one or more
// per level; a level costs one dependent DIR load plus a CTA-local scan.
// A CTA whose region would overflow stops early; the epoch then ends at the smallest number of completed levels
// (later segments are simply ignored and re-expanded in the next epoch).
constexpr int FIND_EPOCH = 32;

__global__ void __launch_bounds__(FIND_THREADS, 1)
k_basin_find(const int32_t *__restrict__ dir, int nc, int nr, int32_t *__restrict__ q_lin,
             int32_t *__restrict__ q_par, long long cap, unsigned long long *slots, FindState *st,
             int32_t *__restrict__ priv_lin, int32_t *__restrict__ priv_par, long long region,
             int32_t *__restrict__ cnt_tab /* [FIND_EPOCH][nb] */, int32_t *__restrict__ done_tab /* [nb] */) {
    __shared__ int s_scan[34];
    __shared__ int s_wsum[FIND_THREADS / 32];
    __shared__ long long s_seg[FIND_EPOCH + 1];      // start of every level's segment inside my region
    __shared__ int s_mycnt[FIND_EPOCH];
    __shared__ long long s_off[FIND_EPOCH];          // final queue position of my segment of every level
    __shared__ int s_total[FIND_EPOCH];
    extern __shared__ int32_t s_dyn[];      // [FIND_CAP] current share, [FIND_CAP] next share, [FIND_CAP] bytes of masks
    int32_t *s_cur = s_dyn, *s_nxt = s_dyn + FIND_CAP;
    unsigned char *s_mask = reinterpret_cast<unsigned char *>(s_dyn + 2 * FIND_CAP);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, b = blockIdx.x, nb = gridDim.x;
    const float rnc = 1.0f / (float)nc;
    int32_t *my_lin = priv_lin + (size_t)b * region, *my_par = priv_par + (size_t)b * region;
    long long fs = 0, fe = 1;             // numbered frontier = queue positions [fs, fe)
    int depth = 1, err = 0;               // the outlet's level is already in the queue
    unsigned stamp = 1;
    for (;;) {
        // ---- my share of the frontier, in queue order ----
        const long long F = fe - fs;
        const long long chunk = ((F + nb - 1) / nb + 31) / 32 * 32;
        long long lo = fs + (long long)b * chunk, hi = lo + chunk;
        if (lo > fe) lo = fe;
        if (hi > fe) hi = fe;
        // ---- expand up to FIND_EPOCH levels on my own ----
        long long used = 0;               // entries of my region in use
        long long prev_start = 0;         // where the current level's segment starts in my region (levels >= 1)
        int mine = (int)(hi - lo);
        bool cur_in_smem = false;
        int done = 0;
        for (int e = 0; e < FIND_EPOCH; ++e) {
            if (mine == 0) {              // nothing below me any more: the remaining levels are empty
                for (int k = e + tid; k < FIND_EPOCH; k += FIND_THREADS) { cnt_tab[k * nb + b] = 0; s_mycnt[k] = 0; s_seg[k] = used; }
                done = FIND_EPOCH;
                break;
            }
            const int c = (mine + FIND_THREADS - 1) / FIND_THREADS;
            const long long seg = (long long)wid * 32 * c;   // lane l handles local cells seg + j*32 + l
            // phase 1: child masks and my count
            int cnt = 0;
            for (int j = 0; j < c; ++j) {
                const long long i = seg + (long long)j * 32 + lane;
                if (i < mine) {
                    const bool cached = i < FIND_CAP;
                    int lin;
                    if (cached && cur_in_smem) lin = s_cur[i];
                    else {
                        lin = (e == 0) ? __ldcg(q_lin + lo + i) : __ldcg(my_lin + prev_start + i);
                        if (cached) s_cur[i] = lin;
                    }
                    const unsigned m = probe_children(dir, nc, nr, rnc, lin);
                    if (cached) s_mask[i] = (unsigned char)m;
                    cnt += __popc(m);
                }
            }
            cnt = warp_sum_i(cnt);
            if (lane == 0) s_wsum[wid] = cnt;
            __syncthreads();
            int wpre = 0, mycnt = 0;
#pragma unroll
            for (int k = 0; k < FIND_THREADS / 32; ++k) {
                const int v = s_wsum[k];
                if (k < wid) wpre += v;
                mycnt += v;
            }
            if (used + mycnt > region) break;  // would not fit: this level is redone in the next epoch (CTA-uniform)
            // phase 2: append my children to my region (and keep them in shared memory as the next frontier)
            long long base = used + wpre;
            for (int j = 0; j < c; ++j) {
                const long long i = seg + (long long)j * 32 + lane;
                unsigned m = 0;
                int lin = 0;
                if (i < mine) {
                    if (i < FIND_CAP) { lin = s_cur[i]; m = s_mask[i]; }
                    else {
                        lin = (e == 0) ? __ldcg(q_lin + lo + i) : __ldcg(my_lin + prev_start + i);
                        m = probe_children(dir, nc, nr, rnc, lin);
                    }
                }
                const int k = __popc(m);
                const int incl = warp_scan_incl(k);
                if (m) write_children(m, lin, nc, base + incl - k, (int)i, my_lin, my_par, s_nxt, used, FIND_CAP);
                base += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (tid == 0) { cnt_tab[e * nb + b] = mycnt; s_mycnt[e] = mycnt; s_seg[e] = used; }
            __syncthreads();              // s_nxt complete, my region's new entries visible to my other warps
            int32_t *t = s_cur; s_cur = s_nxt; s_nxt = t;
            cur_in_smem = true;
            prev_start = used;
            used += mycnt;
            mine = mycnt;
            done = e + 1;
        }
        if (tid == 0) done_tab[b] = done;
        // ---- exchange A: everybody's counts are published ----
        int d0, d1, d2;
        cta_exchange<true>(slots, stamp++, 0, s_scan, &d0, &d1, &d2);
        // levels every CTA completed; totals and my final positions (thread e works on level e)
        int l_end = FIND_EPOCH;
        for (int k = 0; k < nb; ++k) l_end = min(l_end, __ldcg(done_tab + k));
        if (l_end == 0) { err = 2; break; }          // a single level does not fit a region (grid-uniform)
        if (tid < FIND_EPOCH) {
            int tot = 0, before = 0;
            if (tid < l_end)
                for (int k = 0; k < nb; ++k) {
                    const int v = __ldcg(cnt_tab + tid * nb + k);
                    if (k < b) before += v;
                    tot += v;
                }
            s_total[tid] = tot;
            s_off[tid] = before;
        }
        __syncthreads();
        if (tid == 0) {
            long long run = fe;
            for (int e = 0; e < l_end; ++e) { s_off[e] += run; run += s_total[e]; }
        }
        __syncthreads();
        // the BFS ends at the first empty level; levels after it do not exist
        int n_lev = l_end;
        for (int e = 0; e < l_end; ++e)
            if (s_total[e] == 0) { n_lev = e; break; }
        long long new_fs = fe, new_fe = fe;
        for (int e = 0; e < n_lev; ++e) { new_fs = new_fe; new_fe += s_total[e]; }
        if (new_fe > cap) { err = 1; break; }
        // ---- copy my segments to their queue positions; a parent's queue position = start of my segment of the
        // previous level (my share of the frontier for the first level) + its index in that segment ----
        for (int e = 0; e < n_lev; ++e) {
            const long long src = s_seg[e], dst = s_off[e];
            const long long pbase = (e == 0) ? lo : s_off[e - 1];
            const int n_e = s_mycnt[e];
            for (int i = tid; i < n_e; i += FIND_THREADS) {
                q_lin[dst + i] = my_lin[src + i];
                q_par[dst + i] = (int32_t)(pbase + my_par[src + i] + 1);
            }
        }
        depth += n_lev;
        const bool finished = (n_lev < l_end) || (new_fe == new_fs);
        fs = new_fs;
        fe = new_fe;
        if (finished) {
            // depth counts levels including the last, empty expansion (the Fortran's loop runs once more)
            break;
        }
        // ---- exchange B: the new frontier is in the queue for everybody ----
        cta_exchange<true>(slots, stamp++, 0, s_scan, &d0, &d1, &d2);
    }
    if (b == 0 && tid == 0) { st->n = fe; st->depth = depth; st->err = err; }
}

__global__ void k_find_seed(int32_t *q_lin, int32_t *q_par, int lin, unsigned long long *slots, FindState *st) {
    q_lin[0] = lin;
    q_par[0] = 0;
    for (int i = 0; i < 2 * FIND_MAX_CTAS; ++i) slots[i] = 0;
    st->n = 0; st->depth = 0; st->err = 0;
}

// cuencas.f90:592-607 : basin_f(:,i) = (drain id, col, row) of queue position n-i+1
__global__ void k_basin_cut(const int32_t *__restrict__ q_lin, const int32_t *__restrict__ q_par, int n, int nc,
                            int32_t *__restrict__ basin_f) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int p = n - 1 - i;
    const int lin = q_lin[p];
    basin_f[3 * (size_t)i + 0] = q_par[p];
    basin_f[3 * (size_t)i + 1] = lin % nc + 1;
    basin_f[3 * (size_t)i + 2] = lin / nc + 1;
}

// ---------------------------------------------------------------------------------------------------
// accumulation sweeps
// ---------------------------------------------------------------------------------------------------
__global__ void k_fill_u64(unsigned long long *w, long long n, unsigned long long v) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) w[i] = v;
}

// high half of w[parent] counts down one per child; flags a drain table that is not upstream-first
__global__ void k_count_children_u64(const int32_t *__restrict__ drain, int stride, int n,
                                     unsigned long long *__restrict__ w, int *__restrict__ err) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int d = drain[(size_t)i * stride];
    if (d == 0) return;
    const int p = n - d;
    if (d < 0 || p <= i || p >= n) { *err = 1; return; }
    atomicAdd(&w[p], 0xFFFFFFFF00000000ull);
}

// cuencas.f90:659-680.  w[i] = (arrivals - nchildren) << 32 | partial count.  A leaf climbs while it is the last
// child to arrive at each ancestor; the atomic's return value carries the finished subtotal.
// (the leaf test must not read w: a finished inner cell also shows a zero counter, so leaves are flagged up front)
__global__ void k_mark_leaves(const unsigned long long *__restrict__ w, int n, int32_t *__restrict__ leaf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) leaf[i] = ((unsigned)(w[i] >> 32) == 0u) ? 1 : 0;
}

__global__ void k_acum_climb(const int32_t *__restrict__ drain, int stride, int n,
                             unsigned long long *__restrict__ w, const int32_t *__restrict__ leaf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (!leaf[i]) return;  // has children: the last of them to arrive finishes this cell
    unsigned v = 1u;
    int cur = i;
    for (;;) {
        const int d = drain[(size_t)cur * stride];
        if (d == 0) break;
        const int p = n - d;
        const unsigned long long old = atomicAdd(&w[p], (1ull << 32) + (unsigned long long)v);
        if ((unsigned)(old >> 32) != 0xFFFFFFFFu) break;  // not the last arriver
        v = (unsigned)old + v;
        cur = p;
    }
}

__global__ void k_acum_extract(const unsigned long long *__restrict__ w, int n, int32_t *__restrict__ acum) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) acum[i] = (int32_t)(unsigned)w[i];
}

// --- ordered (floating point) variant: CSR of children, ascending ---
__global__ void k_count_children_i32(const int32_t *__restrict__ drain, int stride, int n,
                                     int32_t *__restrict__ cnt, int *__restrict__ err) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int d = drain[(size_t)i * stride];
    if (d == 0) return;
    const int p = n - d;
    if (d < 0 || p <= i || p >= n) { *err = 1; return; }
    atomicAdd(&cnt[p], 1);
}

__global__ void k_fill_children(const int32_t *__restrict__ drain, int stride, int n,
                                const int32_t *__restrict__ start, int32_t *__restrict__ cursor,
                                int32_t *__restrict__ child) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int d = drain[(size_t)i * stride];
    if (d == 0) return;
    const int p = n - d;
    const int slot = atomicAdd(&cursor[p], 1);
    child[start[p] + slot] = i;
}

__global__ void k_sort_children(const int32_t *__restrict__ start, const int32_t *__restrict__ cnt, int n,
                                int32_t *__restrict__ child) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int s = start[j], c = cnt[j];
    for (int a = 1; a < c; ++a) {  // insertion sort; D8 lists have at most 8 entries
        const int key = child[s + a];
        int k = a - 1;
        while (k >= 0 && child[s + k] > key) {
            child[s + k + 1] = child[s + k];
            --k;
        }
        child[s + k + 1] = key;
    }
}

// cuencas.f90:681-701 (positive_only = 0) and :1811-1833 (positive_only = 1)
__global__ void k_acum_var_climb(const int32_t *__restrict__ drain, int stride, int n,
                                 const int32_t *__restrict__ start, const int32_t *__restrict__ cnt,
                                 const int32_t *__restrict__ child, int32_t *__restrict__ pending,
                                 const float *__restrict__ var, float *val, int positive_only) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (cnt[i] != 0) return;
    val[i] = var[i];
    int cur = i;
    for (;;) {
        const int d = drain[(size_t)cur * stride];
        if (d == 0) break;
        const int p = n - d;
        __threadfence();  // publish val[cur] before announcing the arrival
        const int old = atomicSub(&pending[p], 1);
        if (old != 1) break;
        __threadfence();
        float acc = var[p];
        const int s = start[p], c = cnt[p];
        for (int k = 0; k < c; ++k) {
            const float v = __ldcg(val + child[s + k]);
            if (!positive_only || v > 0.0f) acc = acc + v;
        }
        val[p] = acc;
        cur = p;
    }
}

// ---------------------------------------------------------------------------------------------------
// device-wide exclusive scan (int32), three small kernels
// ---------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(const int32_t *__restrict__ in, int32_t *__restrict__ out,
                                                             long long n, long long *__restrict__ tile_sum) {
    __shared__ int sm[33];
    const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        s += v[k];
    }
    int tot;
    int ex = block_scan_excl(s, sm, &tot);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) out[base + k] = ex;
        ex += v[k];
    }
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024) k_scan_tile_sums(long long *tile_sum, int ntiles, long long *total) {
    __shared__ long long sm[1024];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < ntiles; base += 1024) {
        const int j = base + threadIdx.x;
        long long v = (j < ntiles) ? tile_sum[j] : 0;
        sm[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            long long t = (threadIdx.x >= o) ? sm[threadIdx.x - o] : 0;
            __syncthreads();
            sm[threadIdx.x] += t;
            __syncthreads();
        }
        if (j < ntiles) tile_sum[j] = carry + sm[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += sm[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_add(int32_t *__restrict__ out, long long n,
                                                           const long long *__restrict__ tile_sum) {
    const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
    const int add = (int)tile_sum[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < n) out[base + k] += add;
}

int wmf_device_exclusive_scan_i32(wmf_ctx *ctx, Scope &sc, const int32_t *in, int32_t *out, int64_t n,
                                  int64_t *total_host) {
    if (n <= 0) {
        if (total_host) *total_host = 0;
        return WMF_OK;
    }
    const int ntiles = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
    long long *tile_sum;
    WMF_TRY(sc.alloc(&tile_sum, (size_t)ntiles + 1));
    LAUNCH(ctx, k_scan_tiles, ntiles, SCAN_THREADS, 0, in, out, (long long)n, tile_sum);
    LAUNCH(ctx, k_scan_tile_sums, 1, 1024, 0, tile_sum, ntiles, tile_sum + ntiles);
    LAUNCH(ctx, k_scan_add, ntiles, SCAN_THREADS, 0, out, (long long)n, tile_sum);
    if (total_host) {
        long long t;
        CU_TRY(ctx, cudaMemcpyAsync(&t, tile_sum + ntiles, sizeof t, cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        *total_host = t;
    }
    return WMF_OK;
}

// ---------------------------------------------------------------------------------------------------
// per-cell kernels
// ---------------------------------------------------------------------------------------------------
// cuencas.f90:623-644 (long, pend, elev) ; acum comes from the climb sweep
__global__ void k_basics_cell(const int32_t *__restrict__ basin_f, const float *__restrict__ dem,
                              const int32_t *__restrict__ dir, int n, float dxp, float *__restrict__ lng,
                              float *__restrict__ pend, float *__restrict__ elev) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    auto cell_long = [&](int k) { return (dir[k] % 2 == 0) ? dxp : dxp * 1.41421354f; };  // dxp*sqrt(2.) in fp32
    auto cell_pend = [&](int k, float lk) {  // pend of a non-outlet cell k
        const int p = n - basin_f[3 * (size_t)k];
        float s = fabsf(dem[k] - dem[p]) / lk;
        return (s == 0.0f) ? 0.001f : s;
    };
    const float li = cell_long(i);
    lng[i] = li;
    elev[i] = dem[i];
    float s;
    if (basin_f[3 * (size_t)i] != 0) {
        s = cell_pend(i, li);
    } else {  // outlet copies the previous cell's slope (cuencas.f90:640)
        s = 0.0f;
        if (i > 0 && basin_f[3 * (size_t)(i - 1)] != 0) s = cell_pend(i - 1, cell_long(i - 1));
        if (s == 0.0f) s = 0.001f;
    }
    pend[i] = s;
}

// fp64 sums of two fp32 arrays (pend_media, elevacion) + column / row histograms for the median centroid
__global__ void k_basics_reduce(const int32_t *__restrict__ basin_f, const float *__restrict__ pend,
                                const float *__restrict__ dem, int n, double *__restrict__ sums,
                                int32_t *__restrict__ hist_col, int32_t *__restrict__ hist_row) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double a = 0.0, b = 0.0;
    int kc = -1, kr = -1;
    if (i < n) {
        a = (double)pend[i];
        b = (double)dem[i];
        kc = basin_f[3 * (size_t)i + 1] - 1;
        kr = basin_f[3 * (size_t)i + 2] - 1;
    }
    {   // consecutive cells of the table lie on one BFS ring and mostly share a row or a column: one atomic per
        // distinct key of the warp instead of one per cell
        const int lane = threadIdx.x & 31;
        const unsigned pc = __match_any_sync(0xffffffffu, kc), pr = __match_any_sync(0xffffffffu, kr);
        if (kc >= 0 && lane == __ffs(pc) - 1) atomicAdd(&hist_col[kc], __popc(pc));
        if (kr >= 0 && lane == __ffs(pr) - 1) atomicAdd(&hist_row[kr], __popc(pr));
    }
    a = warp_sum_d(a);
    b = warp_sum_d(b);
    __shared__ double sa[32], sb[32];
    if ((threadIdx.x & 31) == 0) { sa[threadIdx.x >> 5] = a; sb[threadIdx.x >> 5] = b; }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int nw = blockDim.x >> 5;
        a = (threadIdx.x < nw) ? sa[threadIdx.x] : 0.0;
        b = (threadIdx.x < nw) ? sb[threadIdx.x] : 0.0;
        a = warp_sum_d(a);
        b = warp_sum_d(b);
        if (threadIdx.x == 0) { atomicAdd(&sums[0], a); atomicAdd(&sums[1], b); }
    }
}

// cuencas.f90:797-808
__global__ void k_coordxy(const int32_t *__restrict__ basin_f, int n, float xll, float yll, float dx, int nrows,
                          float *__restrict__ x, float *__restrict__ y) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    x[i] = xll + dx * ((float)basin_f[3 * (size_t)i + 1] - 0.5f);
    y[i] = yll + dx * ((float)(nrows - basin_f[3 * (size_t)i + 2]) + 0.5f);
}

// sum / count of the non-nodata cells of a raster (cuencas.f90:1159-1161); fp64 accumulation
__global__ void k_map_mean(const float *__restrict__ mapa, long long tot, float nodatam, double *__restrict__ sum,
                           unsigned long long *__restrict__ cnt) {
    double s = 0.0;
    unsigned long long c = 0;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < tot; k += (long long)gridDim.x * blockDim.x) {
        const float v = mapa[k];
        if (v != nodatam) { s += (double)v; ++c; }
    }
    s = warp_sum_d(s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) { atomicAdd(sum, s); atomicAdd(cnt, c); }
}

// cuencas.f90:1163-1183
__global__ void k_map2basin(const int32_t *__restrict__ basin_f, int n, const float *__restrict__ mapa, float xll,
                            float yll, float dx, float dy, int nrows, float xllm, float yllm, int ncolsm, int nrowsm,
                            float dxm, float dym, float nodatam, const double *__restrict__ sum,
                            const unsigned long long *__restrict__ cnt, float *__restrict__ vec) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float xpos = xll + dx * ((float)basin_f[3 * (size_t)i + 1] - 0.5f);
    const float ypos = yll + dy * ((float)(nrows - basin_f[3 * (size_t)i + 2]) + 0.5f);
    float v = nodatam;
    if (xpos > xllm && xpos < (xllm + dxm * (float)ncolsm) && ypos > yllm && ypos < (yllm + (float)nrowsm * dym)) {
        int columna = (int)ceilf((xpos - xllm) / dxm);
        int fila = (int)ceilf((ypos - yllm) / dym);
        fila = nrowsm - fila + 1;
        columna = min(max(columna, 1), ncolsm);
        fila = min(max(fila, 1), nrowsm);
        v = mapa[(size_t)(fila - 1) * ncolsm + (columna - 1)];
        if (v == nodatam) v = (float)(*sum / (double)(*cnt));
    }
    vec[i] = v;
}

__global__ void k_minmax_colrow(const int32_t *__restrict__ basin_f, int n, int32_t *__restrict__ mm) {
    int cmin = INT_MAX, cmax = INT_MIN, fmin = INT_MAX, fmax = INT_MIN;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int c = basin_f[3 * (size_t)i + 1], f = basin_f[3 * (size_t)i + 2];
        cmin = min(cmin, c); cmax = max(cmax, c); fmin = min(fmin, f); fmax = max(fmax, f);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cmin = min(cmin, __shfl_xor_sync(0xffffffffu, cmin, o));
        cmax = max(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
        fmin = min(fmin, __shfl_xor_sync(0xffffffffu, fmin, o));
        fmax = max(fmax, __shfl_xor_sync(0xffffffffu, fmax, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&mm[0], cmin); atomicMax(&mm[1], cmax); atomicMin(&mm[2], fmin); atomicMax(&mm[3], fmax);
    }
}

__global__ void k_fill_f32(float *a, long long n, float v) {
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) a[k] = v;
}
__global__ void k_fill_i32(int32_t *a, long long n, int32_t v) {
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) a[k] = v;
}

// cuencas.f90:1224-1228 / :1048-1050 : scatter a basin vector into a raster window
__global__ void k_scatter_map(const int32_t *__restrict__ basin_f, const float *__restrict__ var, int n, int col0,
                              int fil0, int ncm, float *__restrict__ mapa) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = basin_f[3 * (size_t)i + 1] - col0, f = basin_f[3 * (size_t)i + 2] - fil0;
    mapa[(size_t)f * ncm + c] = var[i];
}

// cuencas.f90:1351-1353 (ge = 1) and :2127-2128 (ge = 0)
__global__ void k_threshold_mask(const int32_t *__restrict__ acum, long long n, int umbral, int ge,
                                 int32_t *__restrict__ cauce, unsigned long long *__restrict__ count) {
    unsigned long long c = 0;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
        const int a = acum[k];
        const int m = ge ? (a >= umbral) : (a > umbral);
        cauce[k] = m;
        c += m;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (count && (threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
}

// cuencas.f90:1455-1462
__global__ void k_stream_type(const int32_t *__restrict__ acum, int n, const int32_t *__restrict__ umbrales, int nu,
                              int32_t *__restrict__ types) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int a = acum[j];
    int t = 1;
    for (int i = 0; i < nu; ++i)
        if (a > umbrales[i]) t = i + 2;
    types[j] = t;
}

// cuencas.f90:2131-2174 : every hillslope cell walks down to its first channel cell
__global__ void k_geo_hand(const int32_t *__restrict__ basin_f, const float *__restrict__ elev,
                           const float *__restrict__ lng, const int32_t *__restrict__ cauce, int n,
                           float *__restrict__ hand, float *__restrict__ hdnd, int32_t *__restrict__ a_quien) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float h = 0.0f, hd = 0.0f;
    int aq = 0;
    if (cauce[i] != 1) {
        int it = i;
        float lsum = lng[i];
        for (;;) {
            const int d = basin_f[3 * (size_t)it];
            if (d == 0) break;  // non-channel outlet: 0,0,0
            const int p = n - d;
            if (cauce[p] == 1) {
                h = elev[i] - elev[p];
                hd = lsum;
                aq = p + 1;
                break;
            } else if (basin_f[3 * (size_t)p] == 0) {
                break;
            } else {
                it = p;
                lsum = lsum + lng[p];
            }
        }
    }
    hand[i] = h;
    hdnd[i] = hd;
    a_quien[i] = aq;
}

// cuencas.f90:2175-2219 : raster-wide HAND, one thread per raster cell
__global__ void k_geo_hand_global(const float *__restrict__ dem, const int32_t *__restrict__ dir,
                                  const int32_t *__restrict__ red, int nc, int nr, int gncols, int gnrows,
                                  float nodata, float *__restrict__ hand) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long tot = (long long)nc * nr;
    if (k >= tot) return;
    const int r0 = red[k];
    float h = (r0 == 1) ? 0.0f : nodata;
    if ((float)dir[k] != nodata && r0 == 0) {
        int col = (int)(k % nc), fil = (int)(k / nc);  // 0-based
        long long guard = 0;
        for (;;) {
            const int d = dir[(size_t)fil * nc + col];
            if ((float)d == nodata) break;
            if (d < 1 || d > 9 || d == 5) { h = nodata; break; }
            col += (d - 1) % 3 - 1;  // keypad: 7 8 9 / 4 5 6 / 1 2 3
            fil += 1 - (d - 1) / 3;
            if (col >= 0 && col < gncols && fil >= 0 && fil < gnrows && col < nc && fil < nr) {
                const int r = red[(size_t)fil * nc + col];
                if (r == 1) { h = dem[k] - dem[(size_t)fil * nc + col]; break; }
                if ((float)r == nodata) { h = nodata; break; }
            } else { h = nodata; break; }
            if (++guard > tot) break;
        }
    }
    hand[k] = h;
}

// geo_hand_global by pointer jumping.  HAND of a cell is dem(cell) - dem(first network cell on its way down), so only the
// END of the walk matters: T[p] = what one step from p gives (-1: the walk ends without a value, -(q+2): it reaches network
// cell q, q >= 0: it goes on from q) is squared in place until no cell is pending -- ~log2(longest way to the network)
// streaming rounds instead of one dependent walk per cell (C3, 400 M cells: 73 ms as walks).  Unsynchronised in-place
// updates are safe: whatever version of T[q] a thread reads describes a later point of the same walk.
__global__ void k_hg_step(const int32_t *__restrict__ dir, const int32_t *__restrict__ red, int nc, int nr, int gncols, int gnrows,
                          float nodata, int32_t *__restrict__ T) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (long long)nc * nr) return;
    const int d = dir[k];
    int t = -1;
    if ((float)d != nodata && d >= 1 && d <= 9 && d != 5) {
        const int col = (int)(k % nc) + (d - 1) % 3 - 1, fil = (int)(k / nc) + 1 - (d - 1) / 3;  // keypad: 7 8 9 / 4 5 6 / 1 2 3
        if (col >= 0 && col < gncols && fil >= 0 && fil < gnrows && col < nc && fil < nr) {
            const long long q = (long long)fil * nc + col;
            const int r = red[q];
            if (r == 1) t = -(int)q - 2;
            else if ((float)r != nodata) t = (int)q;
        }
    }
    T[k] = t;
}
// The same first step for a 64 x 64 tile, followed by the squarings that stay inside the tile, in shared memory: what
// leaves the kernel points at the network, at nothing, or at the first cell OUTSIDE the tile, so the global rounds that
// follow cross at least a tile per jump and most cells (the network is rarely 64 cells away) are settled already.
constexpr int HG_TS = 64, HG_TC = HG_TS * HG_TS, HG_THREADS = 512, HG_CPT = HG_TC / HG_THREADS;
__global__ void __launch_bounds__(HG_THREADS)
k_hg_tile(const int32_t *__restrict__ dir, const int32_t *__restrict__ red, int nc, int nr, int gncols, int gnrows, int ntx,
          float nodata, int32_t *__restrict__ T) {
    // encoding inside the tile: 0 .. 4095 pending at that cell of the tile; 4096 + q pending at raster cell q (outside);
    // -1 no value; -(q + 2) network cell q
    __shared__ int Ts[HG_TC];
    const int tid = threadIdx.x, r0 = (blockIdx.x / ntx) * HG_TS, c0 = (blockIdx.x % ntx) * HG_TS;
#pragma unroll
    for (int k = 0; k < HG_CPT; ++k) {
        const int li = k * HG_THREADS + tid, fil0 = r0 + (li >> 6), col0 = c0 + (li & (HG_TS - 1));
        int t = -1;
        if (fil0 < nr && col0 < nc) {
            const int d = dir[(size_t)fil0 * nc + col0];
            if ((float)d != nodata && d >= 1 && d <= 9 && d != 5) {
                const int col = col0 + (d - 1) % 3 - 1, fil = fil0 + 1 - (d - 1) / 3;
                if (col >= 0 && col < gncols && fil >= 0 && fil < gnrows && col < nc && fil < nr) {
                    const long long q = (long long)fil * nc + col;
                    const int r = red[q];
                    if (r == 1) t = -(int)q - 2;
                    else if ((float)r != nodata) {
                        const int ly = fil - r0, lx = col - c0;
                        t = (ly >= 0 && ly < HG_TS && lx >= 0 && lx < HG_TS) ? ly * HG_TS + lx : HG_TC + (int)q;
                    }
                }
            }
        }
        Ts[li] = t;
    }
    __syncthreads();
    for (int round = 0; round < 13; ++round) {
        int live = 0;
#pragma unroll
        for (int k = 0; k < HG_CPT; ++k) {
            const int li = k * HG_THREADS + tid;
            const int t = Ts[li];
            if (t >= 0 && t < HG_TC) {
                const int t2 = Ts[t];
                if (t2 != t) {
                    Ts[li] = t2;
                    live |= (t2 >= 0 && t2 < HG_TC);
                }
            }
        }
        if (!__syncthreads_or(live)) break;
    }
#pragma unroll
    for (int k = 0; k < HG_CPT; ++k) {
        const int li = k * HG_THREADS + tid, fil0 = r0 + (li >> 6), col0 = c0 + (li & (HG_TS - 1));
        if (fil0 < nr && col0 < nc) {
            const int t = Ts[li];
            int g = t;                                                                   // -1 / network cell
            if (t >= HG_TC) g = t - HG_TC;                                               // pending outside the tile
            else if (t >= 0) g = (r0 + (t >> 6)) * nc + c0 + (t & (HG_TS - 1));          // a cycle inside the tile: stays pending
            T[(size_t)fil0 * nc + col0] = g;
        }
    }
}
__global__ void k_hg_jump(long long tot, int32_t *__restrict__ T, int *__restrict__ more) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= tot) return;
    const int t = T[k];
    if (t < 0) return;
    const int t2 = T[t];
    if (t2 == t) return;                 // (a cell that steps onto itself: never resolves, like the walk's guard)
    T[k] = t2;
    if (t2 >= 0) *more = 1;
}
__global__ void k_hg_final(const float *__restrict__ dem, const int32_t *__restrict__ dir, const int32_t *__restrict__ red,
                           const int32_t *__restrict__ T, long long tot, float nodata, float *__restrict__ hand) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= tot) return;
    const int r0 = red[k];
    float h = (r0 == 1) ? 0.0f : nodata;
    if ((float)dir[k] != nodata && r0 == 0) {
        const int t = T[k];
        if (t <= -2) h = dem[k] - dem[-(long long)t - 2];
    }
    hand[k] = h;
}

// cuencas.f90:724-756
__global__ void k_time_to_out(const int32_t *__restrict__ basin_f, const float *__restrict__ lng,
                              const float *__restrict__ speed, int n, float *__restrict__ time) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float acc = lng[i] / speed[i];
    int drenaid = n - basin_f[3 * (size_t)i] + 1;  // 1-based
    if (drenaid < n) {
        for (;;) {
            acc = acc + lng[drenaid - 1] / speed[drenaid - 1];
            drenaid = n - basin_f[3 * (size_t)(drenaid - 1)] + 1;
            if (drenaid >= n) break;
        }
    }
    time[i] = acc;
}

// ---------------------------------------------------------------------------------------------------
// host entry points
// ---------------------------------------------------------------------------------------------------
static int check_flag(wmf_ctx *ctx, int *d_err, const char *what) {
    int e = 0;
    CU_TRY(ctx, cudaMemcpyAsync(&e, d_err, sizeof e, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (e) WMF_FAIL(ctx, WMF_ERR_TOPOLOGY, "%s", what);
    return WMF_OK;
}

// integer subtree sums into d_acum (device pointer)
static int acum_device(wmf_ctx *ctx, Scope &sc, const int32_t *d_drain, int stride, int n, int32_t *d_acum) {
    unsigned long long *w;
    int *d_err;
    WMF_TRY(sc.alloc(&w, (size_t)n));
    WMF_TRY(sc.alloc(&d_err, 1));
    CU_TRY(ctx, cudaMemsetAsync(d_err, 0, sizeof(int), ctx->stream));
    const int B = 256, G = grid_for(n, B);
    LAUNCH(ctx, k_fill_u64, G, B, 0, w, (long long)n, 1ull);
    LAUNCH(ctx, k_count_children_u64, G, B, 0, d_drain, stride, n, w, d_err);
    WMF_TRY(check_flag(ctx, d_err, "drain table is not upstream-first (need 0 <= id < n and parent index > cell index)"));
    LAUNCH(ctx, k_mark_leaves, G, B, 0, w, n, d_acum);  // the output buffer doubles as the leaf flag until extract
    LAUNCH(ctx, k_acum_climb, G, B, 0, d_drain, stride, n, w, d_acum);
    LAUNCH(ctx, k_acum_extract, G, B, 0, w, n, d_acum);
    return WMF_OK;
}

// basin_acum of a (3,N) table: the sizes basin_find computed on its way when the table is the one it produced (every
// drain id is compared), the barrier-free climb for any other upstream-first table
static int acum_table(wmf_ctx *ctx, Scope &sc, const int32_t *d_b, int n, int32_t *d_acum) {
    int used = 0;
    const char *e = getenv("WMF_ACUM_ALGO");
    if (!e || strcmp(e, "climb") != 0) WMF_TRY(wmf_pathfind_acum_cached(ctx, sc, d_b, n, d_acum, &used));
    if (used) return WMF_OK;
    return acum_device(ctx, sc, d_b, 3, n, d_acum);
}

static int acum_var_device(wmf_ctx *ctx, Scope &sc, const int32_t *d_drain, int stride, int n, const float *d_var,
                           float *d_out, int positive_only) {
    int32_t *cnt, *start, *cursor, *child, *pending;
    int *d_err;
    WMF_TRY(sc.alloc(&cnt, (size_t)n));
    WMF_TRY(sc.alloc(&start, (size_t)n));
    WMF_TRY(sc.alloc(&cursor, (size_t)n));
    WMF_TRY(sc.alloc(&child, (size_t)n));
    WMF_TRY(sc.alloc(&pending, (size_t)n));
    WMF_TRY(sc.alloc(&d_err, 1));
    CU_TRY(ctx, cudaMemsetAsync(d_err, 0, sizeof(int), ctx->stream));
    CU_TRY(ctx, cudaMemsetAsync(cnt, 0, sizeof(int32_t) * (size_t)n, ctx->stream));
    CU_TRY(ctx, cudaMemsetAsync(cursor, 0, sizeof(int32_t) * (size_t)n, ctx->stream));
    const int B = 256, G = grid_for(n, B);
    LAUNCH(ctx, k_count_children_i32, G, B, 0, d_drain, stride, n, cnt, d_err);
    WMF_TRY(check_flag(ctx, d_err, "drain table is not upstream-first (need 0 <= id < n and parent index > cell index)"));
    WMF_TRY(wmf_device_exclusive_scan_i32(ctx, sc, cnt, start, n, nullptr));
    LAUNCH(ctx, k_fill_children, G, B, 0, d_drain, stride, n, start, cursor, child);
    LAUNCH(ctx, k_sort_children, G, B, 0, start, cnt, n, child);
    CU_TRY(ctx, cudaMemcpyAsync(pending, cnt, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
    LAUNCH(ctx, k_acum_var_climb, G, B, 0, d_drain, stride, n, start, cnt, child, pending, d_var, d_out, positive_only);
    return WMF_OK;
}

extern "C" {

int wmf_basin_find(wmf_ctx *ctx, float x, float y, const int32_t *dir, int32_t nc, int32_t nr, int32_t *nceldas) {
    if (!ctx || !dir || !nceldas) return WMF_ERR_ARG;
    const wmf_geo &g = ctx->geo;
    if (nc != g.ncols || nr != g.nrows || nc <= 0 || nr <= 0)
        WMF_FAIL(ctx, WMF_ERR_ARG, "basin_find: DIR is (%d,%d) but cu.ncols,nrows = (%d,%d)", nc, nr, g.ncols, g.nrows);
    const long long celTot = (long long)nc * nr;
    if (celTot >= 2147483647LL) WMF_FAIL(ctx, WMF_ERR_ARG, "basin_find: raster has more than 2^31-1 cells");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    int32_t coli, fili;
    wmf_coord2fil_col(ctx, x, y, &coli, &fili);
    ctx->find_nc = nc;
    ctx->find_nr = nr;
    if (ctx->q_cap < celTot) {
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->q_lin) cudaFree(ctx->q_lin);
        if (ctx->q_par) cudaFree(ctx->q_par);
        if (ctx->q_acum) cudaFree(ctx->q_acum);
        ctx->q_lin = ctx->q_par = ctx->q_acum = nullptr;
        ctx->q_cap = 0;
        CU_TRY(ctx, cudaMalloc(&ctx->q_lin, sizeof(int32_t) * (size_t)celTot));
        CU_TRY(ctx, cudaMalloc(&ctx->q_par, sizeof(int32_t) * (size_t)celTot));
        CU_TRY(ctx, cudaMalloc(&ctx->q_acum, sizeof(int32_t) * (size_t)celTot));
        ctx->q_cap = celTot;
    }
    ctx->q_acum_valid = false;
    ctx->find_outside = (coli <= 0 || fili <= 0);
    if (coli <= 0 || fili <= 0) {
        // outlet outside the map: the Fortran returns a one-cell "basin" holding (-999 | col, -999 | fil)
        ctx->find_n = 1;
        ctx->find_depth = 1;
        // encode through the queue: lin = (fil-1)*nc + (col-1) cannot hold -999, keep it on the host side
        const int32_t lin = -1, par = 0;
        CU_TRY(ctx, cudaMemcpyAsync(ctx->q_lin, &lin, 4, cudaMemcpyHostToDevice, ctx->stream));
        CU_TRY(ctx, cudaMemcpyAsync(ctx->q_par, &par, 4, cudaMemcpyHostToDevice, ctx->stream));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->scalars[3] = (float)coli;  // stash (unused scalars until basin_basics runs)
        ctx->scalars[4] = (float)fili;
        *nceldas = 1;
        return WMF_OK;
    }
    Scope sc(ctx);
    const int32_t *d_dir;
    WMF_TRY(sc.in(dir, (size_t)celTot, &d_dir));
    const int lin0 = (fili - 1) * nc + (coli - 1);
    // WMF_FIND_ALGO = tiles (default: tile decomposition, pathfind.cu) | bfs (the level-by-level search below, kept as
    // the cross-check)
    const char *algo = getenv("WMF_FIND_ALGO");
    if (!algo || strcmp(algo, "bfs") != 0) {
        int32_t n = 0, depth = 0;
        WMF_TRY(wmf_pathfind(ctx, sc, d_dir, nc, nr, lin0, &n, &depth));
        ctx->find_n = n;
        ctx->find_depth = depth;
        ctx->q_acum_valid = true;
        *nceldas = n;
        return WMF_OK;
    }
    FindState *st;
    unsigned long long *slots;
    WMF_TRY(sc.alloc((void **)&st, sizeof(FindState)));
    WMF_TRY(sc.alloc(&slots, (size_t)2 * FIND_MAX_CTAS));
    LAUNCH(ctx, k_find_seed, 1, 1, 0, ctx->q_lin, ctx->q_par, lin0, slots, st);
    int occ = 0;
    CU_TRY(ctx, cudaFuncSetAttribute(k_basin_find, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FIND_SMEM));
    CU_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_basin_find, FIND_THREADS, FIND_SMEM));
    if (occ < 1) WMF_FAIL(ctx, WMF_ERR_CUDA, "basin_find kernel does not fit on an SM");
    // co-resident CTAs (the exchanges spin).  Measured find time: 1024^2 raster 32 CTAs 2.4 ms (64: 2.7); 8192^2 64 CTAs
    // 25.2 ms (32: 30.8, 128: 29.9); 20000^2 96 CTAs 73.3 ms (64: 90.6, 128: 79.6, 148: 83.6)
    int nb = std::min(ctx->num_sms, celTot >= (200LL << 20) ? 96 : (celTot >= (32LL << 20) ? 64 : 32));
    if (const char *e = getenv("WMF_FIND_CTAS")) nb = std::max(1, std::min(std::min(ctx->num_sms, FIND_MAX_CTAS), atoi(e)));
    // private regions: what a CTA discovers during one epoch (raster index + parent index per cell)
    long long region = 2 * celTot / nb + 65536;
    int32_t *priv_lin, *priv_par, *cnt_tab, *done_tab;
    WMF_TRY(sc.alloc(&priv_lin, (size_t)region * nb));
    WMF_TRY(sc.alloc(&priv_par, (size_t)region * nb));
    WMF_TRY(sc.alloc(&cnt_tab, (size_t)FIND_EPOCH * nb));
    WMF_TRY(sc.alloc(&done_tab, (size_t)nb));
    long long cap = celTot;
    void *args[] = {(void *)&d_dir, (void *)&nc, (void *)&nr, (void *)&ctx->q_lin, (void *)&ctx->q_par, (void *)&cap,
                    (void *)&slots, (void *)&st, (void *)&priv_lin, (void *)&priv_par, (void *)&region, (void *)&cnt_tab,
                    (void *)&done_tab};
    CU_TRY(ctx, cudaLaunchCooperativeKernel((void *)k_basin_find, dim3(nb), dim3(FIND_THREADS), args, FIND_SMEM, ctx->stream));
    ctx->launches++;
    FindState h;
    CU_TRY(ctx, cudaMemcpyAsync(&h, st, sizeof(FindState), cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (h.err == 2) WMF_FAIL(ctx, WMF_ERR_NOMEM, "basin_find: one BFS level does not fit the per-CTA scratch regions");
    if (h.err) WMF_FAIL(ctx, WMF_ERR_TOPOLOGY, "basin_find: BFS queue overflow -- the DIR map is cyclic");
    ctx->find_n = (int32_t)h.n;
    ctx->find_depth = h.depth;
    *nceldas = ctx->find_n;
    return WMF_OK;
}

int wmf_basin_find_depth(wmf_ctx *ctx, int32_t *depth) {
    if (!ctx || !depth) return WMF_ERR_ARG;
    *depth = ctx->find_depth;
    return WMF_OK;
}

int wmf_basin_cut(wmf_ctx *ctx, int32_t nceldas, int32_t *basin_f) {
    if (!ctx || !basin_f) return WMF_ERR_ARG;
    if (!ctx->q_lin || ctx->find_n <= 0) WMF_FAIL(ctx, WMF_ERR_STATE, "basin_cut: call basin_find first");
    if (nceldas <= 0 || nceldas > ctx->find_n)
        WMF_FAIL(ctx, WMF_ERR_ARG, "basin_cut: nceldas=%d but the last basin_find traced %d cells", nceldas, ctx->find_n);
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    Scope sc(ctx);
    int32_t *d_out;
    WMF_TRY(sc.out(basin_f, (size_t)3 * nceldas, &d_out));
    if (ctx->find_outside) {  // degenerate outlet outside the map
        const int32_t v[3] = {0, (int32_t)ctx->scalars[3], (int32_t)ctx->scalars[4]};
        CU_TRY(ctx, cudaMemcpyAsync(d_out, v, sizeof v, cudaMemcpyHostToDevice, ctx->stream));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        return sc.finish();
    }
    LAUNCH(ctx, k_basin_cut, grid_for(nceldas, 256), 256, 0, ctx->q_lin, ctx->q_par, nceldas, ctx->find_nc, d_out);
    return sc.finish();
}

int wmf_basin_acum(wmf_ctx *ctx, const int32_t *basin_f, int32_t n, int32_t *acum) {
    if (!ctx || !basin_f || !acum || n <= 0) return WMF_ERR_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    Scope sc(ctx);
    const int32_t *d_b;
    int32_t *d_a;
    WMF_TRY(sc.in(basin_f, (size_t)3 * n, &d_b));
    WMF_TRY(sc.out(acum, (size_t)n, &d_a));
    WMF_TRY(acum_table(ctx, sc, d_b, n, d_a));
    ctx->scalars[0] = (float)n * (ctx->geo.dxp * ctx->geo.dxp) / 1e6f;  // cuencas.f90:679
    return sc.finish();
}

int wmf_basin_acum_var(wmf_ctx *ctx, const int32_t *drain, int32_t n, const float *var, float *acum) {
    if (!ctx || !drain || !var || !acum || n <= 0) return WMF_ERR_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    Scope sc(ctx);
    const int32_t *d_d;
    const float *d_v;
    float *d_o;
    WMF_TRY(sc.in(drain, (size_t)n, &d_d));
    WMF_TRY(sc.in(var, (size_t)n, &d_v));
    WMF_TRY(sc.out(acum, (size_t)n, &d_o));
    WMF_TRY(acum_var_device(ctx, sc, d_d, 1, n, d_v, d_o, 0));
    return sc.finish();
}

int wmf_basin_propagate(wmf_ctx *ctx, const int32_t *basin_f, const float *var, int32_t n, float *var_prop) {
    if (!ctx || !basin_f || !var || !var_prop || n <= 0) return WMF_ERR_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    Scope sc(ctx);
    const int32_t *d_b;
    const float *d_v;
    float *d_o;
    WMF_TRY(sc.in(basin_f, (size_t)3 * n, &d_b));
    WMF_TRY(sc.in(var, (size_t)n, &d_v));
    WMF_TRY(sc.out(var_prop, (size_t)n, &d_o));
    WMF_TRY(acum_var_device(ctx, sc, d_b, 3, n, d_v, d_o, 1));
    return sc.finish();
}

int wmf_basin_basics(wmf_ctx *ctx, const int32_t *basin_f, const float *dem, const int32_t *dir, int32_t n,
                     int32_t *acum, float *lng, float *pend, float *elev) {
    if (!ctx || !basin_f || !dem || !dir || !acum || !lng || !pend || !elev || n <= 0) return WMF_ERR_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    const wmf_geo &g = ctx->geo;
    Scope sc(ctx);
    const int32_t *d_b, *d_dir;
    const float *d_dem;
    int32_t *d_a;
    float *d_l, *d_p, *d_e;
    WMF_TRY(sc.in(basin_f, (size_t)3 * n, &d_b));
    WMF_TRY(sc.in(dem, (size_t)n, &d_dem));
    WMF_TRY(sc.in(dir, (size_t)n, &d_dir));
    WMF_TRY(sc.out(acum, (size_t)n, &d_a));
    WMF_TRY(sc.out(lng, (size_t)n, &d_l));
    WMF_TRY(sc.out(pend, (size_t)n, &d_p));
    WMF_TRY(sc.out(elev, (size_t)n, &d_e));
    WMF_TRY(acum_table(ctx, sc, d_b, n, d_a));
    LAUNCH(ctx, k_basics_cell, grid_for(n, 256), 256, 0, d_b, d_dem, d_dir, n, g.dxp, d_l, d_p, d_e);
    // scalars: means in fp64, centroid = median column / row (the Fortran sorts X and Y, cuencas.f90:650-657)
    double *d_sums;
    int32_t *d_hist;
    WMF_TRY(sc.alloc(&d_sums, 2));
    WMF_TRY(sc.alloc(&d_hist, (size_t)g.ncols + g.nrows));
    CU_TRY(ctx, cudaMemsetAsync(d_sums, 0, 2 * sizeof(double), ctx->stream));
    CU_TRY(ctx, cudaMemsetAsync(d_hist, 0, sizeof(int32_t) * ((size_t)g.ncols + g.nrows), ctx->stream));
    LAUNCH(ctx, k_basics_reduce, grid_for(n, 256), 256, 0, d_b, d_p, d_dem, n, d_sums, d_hist, d_hist + g.ncols);
    double h_sums[2];
    std::vector<int32_t> h_hist((size_t)g.ncols + g.nrows);
    CU_TRY(ctx, cudaMemcpyAsync(h_sums, d_sums, sizeof h_sums, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaMemcpyAsync(h_hist.data(), d_hist, sizeof(int32_t) * h_hist.size(), cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->scalars[0] = (float)n * (g.dxp * g.dxp) / 1e6f;
    ctx->scalars[1] = (float)(h_sums[0] / (double)n);
    ctx->scalars[2] = (float)(h_sums[1] / (double)n);
    const long long k = (n % 2 == 0) ? n / 2 : (n + 1) / 2;  // 1-based rank in the sorted X / Y
    {
        // X grows with col: k-th smallest X = k-th smallest col
        long long run = 0;
        int col = 1;
        for (int c = 0; c < g.ncols; ++c) { run += h_hist[c]; if (run >= k) { col = c + 1; break; } }
        volatile float t = (float)col - 0.5f;
        ctx->scalars[3] = g.xll + g.dx * t;
        // Y shrinks with row: k-th smallest Y = k-th largest row
        run = 0;
        int row = g.nrows;
        for (int f = g.nrows - 1; f >= 0; --f) { run += h_hist[g.ncols + f]; if (run >= k) { row = f + 1; break; } }
        volatile float u = (float)(g.nrows - row) + 0.5f;
        ctx->scalars[4] = g.yll + g.dx * u;
    }
    return sc.finish();
}

int wmf_basin_coordxy(wmf_ctx *ctx, const int32_t *basin_f, int32_t n, float *x, float *y) {
    if (!ctx || !basin_f || !x || !y || n <= 0) return WMF_ERR_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    const wmf_geo &g = ctx->geo;
    Scope sc(ctx);
    const int32_t *d_b;
    float *d_x, *d_y;
    WMF_TRY(sc.in(basin_f, (size_t)3 * n, &d_b));
    WMF_TRY(sc.out(x, (size_t)n, &d_x));
    WMF_TRY(sc.out(y, (size_t)n, &d_y));
    LAUNCH(ctx, k_coordxy, grid_for(n, 256), 256, 0, d_b, n, g.xll, g.yll, g.dx, g.nrows, d_x, d_y);
    return sc.finish();
}

int wmf_basin_map2basin(wmf_ctx *ctx, const int32_t *basin_f, int32_t n, const float *mapa, float xllm, float yllm,
                        int32_t ncolsm, int32_t nrowsm, float dxm, float dym, float nodatam, float *vec) {
    if (!ctx || !basin_f || !mapa || !vec || n <= 0 || ncolsm <= 0 || nrowsm <= 0) return WMF_ERR_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    const wmf_geo &g = ctx->geo;
    Scope sc(ctx);
    const int32_t *d_b;
    const float *d_m;
    float *d_v;
    const long long tot = (long long)ncolsm * nrowsm;
    WMF_TRY(sc.in(basin_f, (size_t)3 * n, &d_b));
    WMF_TRY(sc.in(mapa, (size_t)tot, &d_m));
    WMF_TRY(sc.out(vec, (size_t)n, &d_v));
    double *d_sum;
    WMF_TRY(sc.alloc(&d_sum, 2));  // [0] sum, [1] reinterpreted as u64 count
    CU_TRY(ctx, cudaMemsetAsync(d_sum, 0, 2 * sizeof(double), ctx->stream));
    int G = grid_for(tot, 256);
    if (G > ctx->num_sms * 16) G = ctx->num_sms * 16;
    LAUNCH(ctx, k_map_mean, G, 256, 0, d_m, tot, nodatam, d_sum, (unsigned long long *)(d_sum + 1));
    LAUNCH(ctx, k_map2basin, grid_for(n, 256), 256, 0, d_b, n, d_m, g.xll, g.yll, g.dx, g.dy, g.nrows, xllm, yllm, ncolsm,
           nrowsm, dxm, dym, nodatam, d_sum, (const unsigned long long *)(d_sum + 1), d_v);
    return sc.finish();
}

static int minmax_colrow(wmf_ctx *ctx, Scope &sc, const int32_t *d_b, int n, int32_t mm[4]) {
    int32_t *d_mm;
    WMF_TRY(sc.alloc(&d_mm, 4));
    const int32_t init[4] = {INT32_MAX, INT32_MIN, INT32_MAX, INT32_MIN};
    CU_TRY(ctx, cudaMemcpyAsync(d_mm, init, sizeof init, cudaMemcpyHostToDevice, ctx->stream));
    int G = grid_for(n, 256);
    if (G > ctx->num_sms * 8) G = ctx->num_sms * 8;
    LAUNCH(ctx, k_minmax_colrow, G, 256, 0, d_b, n, d_mm);
    CU_TRY(ctx, cudaMemcpyAsync(mm, d_mm, sizeof init, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return WMF_OK;
}

int wmf_basin_2map_find(wmf_ctx *ctx, const int32_t *basin, int32_t n, int32_t *map_ncols, int32_t *map_nrows) {
    if (!ctx || !basin || !map_ncols || !map_nrows || n <= 0) return WMF_ERR_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    Scope sc(ctx);
    const int32_t *d_b;
    WMF_TRY(sc.in(basin, (size_t)3 * n, &d_b));
    int32_t mm[4];
    WMF_TRY(minmax_colrow(ctx, sc, d_b, n, mm));
    *map_ncols = mm[1] - mm[0] + 1;
    *map_nrows = mm[3] - mm[2] + 1;
    return sc.finish();
}

int wmf_basin_2map(wmf_ctx *ctx, const int32_t *basin, const float *var, int32_t n, int32_t map_ncols,
                   int32_t map_nrows, float *mapa, float *map_xll, float *map_yll) {
    if (!ctx || !basin || !var || !mapa || n <= 0 || map_ncols <= 0 || map_nrows <= 0) return WMF_ERR_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    const wmf_geo &g = ctx->geo;
    Scope sc(ctx);
    const int32_t *d_b;
    const float *d_v;
    float *d_m;
    WMF_TRY(sc.in(basin, (size_t)3 * n, &d_b));
    WMF_TRY(sc.in(var, (size_t)n, &d_v));
    WMF_TRY(sc.out(mapa, (size_t)map_ncols * map_nrows, &d_m));
    int32_t mm[4];
    WMF_TRY(minmax_colrow(ctx, sc, d_b, n, mm));
    if (mm[1] - mm[0] + 1 > map_ncols || mm[3] - mm[2] + 1 > map_nrows)
        WMF_FAIL(ctx, WMF_ERR_ARG, "basin_2map: window (%d,%d) smaller than the basin extent (%d,%d)", map_ncols, map_nrows,
                 mm[1] - mm[0] + 1, mm[3] - mm[2] + 1);
    if (map_xll) { volatile float t = (float)(mm[0] - 1); *map_xll = g.xll + g.dx * t; }
    if (map_yll) { volatile float t = (float)(g.nrows - mm[3]); *map_yll = g.yll + g.dx * t; }
    const long long tot = (long long)map_ncols * map_nrows;
    int G = grid_for(tot, 256);
    if (G > ctx->num_sms * 16) G = ctx->num_sms * 16;
    LAUNCH(ctx, k_fill_f32, G, 256, 0, d_m, tot, g.nodata);
    LAUNCH(ctx, k_scatter_map, grid_for(n, 256), 256, 0, d_b, d_v, n, mm[0], mm[2], map_ncols, d_m);
    return sc.finish();
}

int wmf_basin_float_var2map(wmf_ctx *ctx, const int32_t *basin_f, const float *vect, int32_t n, int32_t nc, int32_t nf,
                            float *mapa) {
    if (!ctx || !basin_f || !vect || !mapa || n <= 0 || nc <= 0 || nf <= 0) return WMF_ERR_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    Scope sc(ctx);
    const int32_t *d_b;
    const float *d_v;
    float *d_m;
    WMF_TRY(sc.in(basin_f, (size_t)3 * n, &d_b));
    WMF_TRY(sc.in(vect, (size_t)n, &d_v));
    WMF_TRY(sc.out(mapa, (size_t)nc * nf, &d_m));
    int32_t mm[4];
    WMF_TRY(minmax_colrow(ctx, sc, d_b, n, mm));
    if (mm[0] < 1 || mm[1] > nc || mm[2] < 1 || mm[3] > nf)
        WMF_FAIL(ctx, WMF_ERR_ARG, "basin_float_var2map: basin cells fall outside the (%d,%d) map", nc, nf);
    const long long tot = (long long)nc * nf;
    int G = grid_for(tot, 256);
    if (G > ctx->num_sms * 16) G = ctx->num_sms * 16;
    LAUNCH(ctx, k_fill_f32, G, 256, 0, d_m, tot, ctx->geo.nodata);
    LAUNCH(ctx, k_scatter_map, grid_for(n, 256), 256, 0, d_b, d_v, n, 1, 1, nc, d_m);
    return sc.finish();
}

static int threshold_mask(wmf_ctx *ctx, const int32_t *acum, int64_t n, int32_t umbral, int ge, int32_t *cauce,
                          int32_t *count) {
    if (!ctx || !acum || !cauce || n <= 0) return WMF_ERR_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    Scope sc(ctx);
    const int32_t *d_a;
    int32_t *d_c;
    WMF_TRY(sc.in(acum, (size_t)n, &d_a));
    WMF_TRY(sc.out(cauce, (size_t)n, &d_c));
    unsigned long long *d_cnt = nullptr;
    if (count) {
        WMF_TRY(sc.alloc(&d_cnt, 1));
        CU_TRY(ctx, cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long), ctx->stream));
    }
    int G = grid_for(n, 256);
    if (G > ctx->num_sms * 16) G = ctx->num_sms * 16;
    LAUNCH(ctx, k_threshold_mask, G, 256, 0, d_a, (long long)n, umbral, ge, d_c, d_cnt);
    if (count) {
        unsigned long long h;
        CU_TRY(ctx, cudaMemcpyAsync(&h, d_cnt, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        *count = (int32_t)h;
    }
    return sc.finish();
}

// cuencas.f90:1340-1379 in three one-thread-per-cell passes (the counts commute: integer atomics)
__global__ void k_nod_count(const int32_t *__restrict__ basin_f, const int32_t *__restrict__ acum, int n, int umbral,
                            int32_t *__restrict__ cauce, int32_t *__restrict__ nodos, int *__restrict__ n_cauce) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int c = 0;
    if (i < n) {
        c = (acum[i] >= umbral) ? 1 : 0;
        cauce[i] = c;
        const int d = basin_f[3 * (size_t)i];
        if (d != 0 && c) atomicAdd(&nodos[n - d], 1);   // channel cells draining into the cell (:1356-1363)
    }
    c = warp_sum_i(c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(n_cauce, c);
}
__global__ void k_nod_class(const int32_t *__restrict__ cauce, int n, int32_t *__restrict__ nodos, int *__restrict__ n_nodos) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int k = 0;
    if (i < n) {
        int v = nodos[i];
        if (v == 0 && cauce[i] == 1) v = 3;             // channel heads (:1365)
        if (i == n - 1) v = 2;                          // the outlet is a node too (:1366)
        if (v < 2) v = 0;                               // cells with a single channel tributary are not nodes (:1367)
        nodos[i] = v;
        k = v > 0 ? 1 : 0;
    }
    k = warp_sum_i(k);
    if ((threadIdx.x & 31) == 0 && k) atomicAdd(n_nodos, k);
}
__global__ void k_nod_trazado(const int32_t *__restrict__ basin_f, const int32_t *__restrict__ cauce,
                              const int32_t *__restrict__ nodos, int n, int32_t *__restrict__ trazado) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int d = basin_f[3 * (size_t)i];
    int t = (d != 0 && nodos[n - d] != 0 && cauce[i] == 1) ? 1 : 0;   // :1371-1377
    if (i == n - 1) t = 1;
    trazado[i] = t;
}

extern "C" int wmf_basin_stream_nod(wmf_ctx *ctx, const int32_t *basin_f, const int32_t *acum, int32_t n, int32_t umbral,
                                    int32_t *cauce, int32_t *nodos, int32_t *trazado, int32_t *n_nodos, int32_t *n_cauce) {
    if (!ctx || !basin_f || !acum || !cauce || !nodos || !trazado || n <= 0) return WMF_ERR_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    Scope sc(ctx);
    const int32_t *d_b, *d_a;
    int32_t *d_c, *d_n, *d_t;
    int *d_cnt;
    WMF_TRY(sc.in(basin_f, (size_t)3 * n, &d_b));
    WMF_TRY(sc.in(acum, (size_t)n, &d_a));
    WMF_TRY(sc.out(cauce, (size_t)n, &d_c));
    WMF_TRY(sc.out(nodos, (size_t)n, &d_n));
    WMF_TRY(sc.out(trazado, (size_t)n, &d_t));
    WMF_TRY(sc.alloc(&d_cnt, 2));
    CU_TRY(ctx, cudaMemsetAsync(d_cnt, 0, 2 * sizeof(int), ctx->stream));
    CU_TRY(ctx, cudaMemsetAsync(d_n, 0, sizeof(int32_t) * (size_t)n, ctx->stream));
    const int G = grid_for(n, 256);
    LAUNCH(ctx, k_nod_count, G, 256, 0, d_b, d_a, n, umbral, d_c, d_n, d_cnt);
    LAUNCH(ctx, k_nod_class, G, 256, 0, d_c, n, d_n, d_cnt + 1);
    LAUNCH(ctx, k_nod_trazado, G, 256, 0, d_b, d_c, d_n, n, d_t);
    int h[2] = {0, 0};
    CU_TRY(ctx, cudaMemcpyAsync(h, d_cnt, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    WMF_TRY(sc.finish());
    if (n_cauce) *n_cauce = h[0];
    if (n_nodos) *n_nodos = h[1];
    return WMF_OK;
}

int wmf_basin_stream_mask(wmf_ctx *ctx, const int32_t *acum, int32_t n, int32_t umbral, int32_t *cauce, int32_t *n_cauce) {
    return threshold_mask(ctx, acum, n, umbral, 1, cauce, n_cauce);
}

int wmf_geo_acum_to_cauce(wmf_ctx *ctx, const int32_t *acum, int64_t ncells, int32_t umbral, int32_t *cauce) {
    return threshold_mask(ctx, acum, ncells, umbral, 0, cauce, nullptr);
}

int wmf_basin_stream_type(wmf_ctx *ctx, const int32_t *acum, int32_t n, const int32_t *umbrales, int32_t nu,
                          int32_t *stream_types) {
    if (!ctx || !acum || !stream_types || n <= 0 || nu < 0 || (nu > 0 && !umbrales)) return WMF_ERR_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    Scope sc(ctx);
    const int32_t *d_a, *d_u;
    int32_t *d_t;
    WMF_TRY(sc.in(acum, (size_t)n, &d_a));
    WMF_TRY(sc.in(umbrales, (size_t)nu, &d_u));
    WMF_TRY(sc.out(stream_types, (size_t)n, &d_t));
    LAUNCH(ctx, k_stream_type, grid_for(n, 256), 256, 0, d_a, n, d_u, nu, d_t);
    return sc.finish();
}

int wmf_geo_hand(wmf_ctx *ctx, const int32_t *basin_f, const float *elev, const float *lng, const int32_t *cauce,
                 int32_t n, float *hand, float *hdnd, int32_t *a_quien) {
    if (!ctx || !basin_f || !elev || !lng || !cauce || !hand || !hdnd || !a_quien || n <= 0) return WMF_ERR_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    Scope sc(ctx);
    const int32_t *d_b, *d_c;
    const float *d_e, *d_l;
    float *d_h, *d_hd;
    int32_t *d_aq;
    WMF_TRY(sc.in(basin_f, (size_t)3 * n, &d_b));
    WMF_TRY(sc.in(elev, (size_t)n, &d_e));
    WMF_TRY(sc.in(lng, (size_t)n, &d_l));
    WMF_TRY(sc.in(cauce, (size_t)n, &d_c));
    WMF_TRY(sc.out(hand, (size_t)n, &d_h));
    WMF_TRY(sc.out(hdnd, (size_t)n, &d_hd));
    WMF_TRY(sc.out(a_quien, (size_t)n, &d_aq));
    LAUNCH(ctx, k_geo_hand, grid_for(n, 256), 256, 0, d_b, d_e, d_l, d_c, n, d_h, d_hd, d_aq);
    return sc.finish();
}

int wmf_geo_hand_global(wmf_ctx *ctx, const float *dem, const int32_t *dir, const int32_t *red, int32_t nc, int32_t nr,
                        float *hand) {
    if (!ctx || !dem || !dir || !red || !hand || nc <= 0 || nr <= 0) return WMF_ERR_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    const wmf_geo &g = ctx->geo;
    Scope sc(ctx);
    const long long tot = (long long)nc * nr;
    const float *d_dem;
    const int32_t *d_dir, *d_red;
    float *d_h;
    WMF_TRY(sc.in(dem, (size_t)tot, &d_dem));
    WMF_TRY(sc.in(dir, (size_t)tot, &d_dir));
    WMF_TRY(sc.in(red, (size_t)tot, &d_red));
    WMF_TRY(sc.out(hand, (size_t)tot, &d_h));
    const char *algo = getenv("WMF_HAND_ALGO");     // jump (default) | walk (one dependent walk per cell, the cross-check)
    if (tot >= 2147483640LL - HG_TC || (algo && strcmp(algo, "walk") == 0)) {
        LAUNCH(ctx, k_geo_hand_global, grid_for(tot, 256), 256, 0, d_dem, d_dir, d_red, nc, nr, g.ncols, g.nrows, g.nodata, d_h);
        return sc.finish();
    }
    int32_t *T;
    int *more;
    WMF_TRY(sc.alloc(&T, (size_t)tot));
    WMF_TRY(sc.alloc(&more, 1));
    const int G = grid_for(tot, 256);
    if (algo && strcmp(algo, "jump_global") == 0) {   // (the version without the tile pass, for comparison)
        LAUNCH(ctx, k_hg_step, G, 256, 0, d_dir, d_red, nc, nr, g.ncols, g.nrows, g.nodata, T);
    } else {
        const int ntx = (nc + HG_TS - 1) / HG_TS, nty = (nr + HG_TS - 1) / HG_TS;
        LAUNCH(ctx, k_hg_tile, ntx * nty, HG_THREADS, 0, d_dir, d_red, nc, nr, g.ncols, g.nrows, ntx, g.nodata, T);
    }
    for (int round = 0; round < 40; ++round) {      // (what is still pending after 40 squarings is a cycle: no value, as the walk's guard)
        CU_TRY(ctx, cudaMemsetAsync(more, 0, sizeof(int), ctx->stream));
        LAUNCH(ctx, k_hg_jump, G, 256, 0, tot, T, more);
        int h_more = 0;
        CU_TRY(ctx, cudaMemcpyAsync(&h_more, more, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        if (!h_more) break;
    }
    LAUNCH(ctx, k_hg_final, G, 256, 0, d_dem, d_dir, d_red, T, tot, g.nodata, d_h);
    return sc.finish();
}

int wmf_basin_time_to_out(wmf_ctx *ctx, const int32_t *basin_f, const float *lng, const float *speed, int32_t n,
                          float *time) {
    if (!ctx || !basin_f || !lng || !speed || !time || n <= 0) return WMF_ERR_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    Scope sc(ctx);
    const int32_t *d_b;
    const float *d_l, *d_s;
    float *d_t;
    WMF_TRY(sc.in(basin_f, (size_t)3 * n, &d_b));
    WMF_TRY(sc.in(lng, (size_t)n, &d_l));
    WMF_TRY(sc.in(speed, (size_t)n, &d_s));
    WMF_TRY(sc.out(time, (size_t)n, &d_t));
    LAUNCH(ctx, k_time_to_out, grid_for(n, 256), 256, 0, d_b, d_l, d_s, n, d_t);
    return sc.finish();
}

}  // extern "C"
