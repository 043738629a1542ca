// subbasin.cu -- sub-basin (hills) topology on the GPU: the routines SimuBasin.__init__ and set_Geomorphology call right
// after the basin cut, even for the cells model (wmf.py:3019-3023, 3664-3676):
//   basin_subbasin_nod / cut (cuencas.f90:1835-1930), basin_subbasin_find (:1979-2010), basin_subbasin_horton (:1931-1978),
//   basin_subbasin_long (:2011-2051), basin_subbasin_map2subbasin (:2052-2092), basin_subbasin_stream_prop (:2093-2115),
//   basin_stream_slope (:1380-1438).
// The reference runs them as nested sequential scans (O(nodes x cells)); here the per-cell work is one thread per cell
// (pointer walks down the drain table, bounded by a reach or a hillslope), the per-hill sums run one thread per hill over
// the hill's cells gathered in ascending cell order (stable radix sort by hill id), so every fp32 sum has the reference's
// order, and the small node tree is numbered by one CTA, breadth first, tributaries in descending cell order as the
// reference's queue finds them.
#include <cub/cub.cuh>

#include "common.cuh"

namespace {

#define DRAIN(b, i) ((b)[3 * (size_t)(i)])

// ---------------------------------------------------------------- basin_subbasin_nod
__global__ void k_sn_count(const int32_t *__restrict__ b, const int32_t *__restrict__ acum, int n, int umbral,
                           int32_t *__restrict__ cauce, int32_t *__restrict__ cnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = acum[i] >= umbral ? 1 : 0;
    cauce[i] = c;
    const int d = DRAIN(b, i);
    if (c && d != 0) atomicAdd(cnt + (n - d), 1);       // junctions: cells with two or more channel cells draining in (:1856-1861)
}
__global__ void k_sn_flag(const int32_t *__restrict__ b, const int32_t *__restrict__ cauce, const int32_t *__restrict__ cnt, int n,
                          int32_t *__restrict__ nf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int d = DRAIN(b, i);
    int f = (i == n - 1) ? 1 : 0;                       // nodos_fin(nceldas) = 1
    if (cauce[i] == 1 && d != 0 && cnt[n - d] >= 2) f = 1;   // a channel cell draining into a junction (:1866-1877)
    nf[i] = f;
}
__global__ void k_sn_compact(const int32_t *__restrict__ nf, const int32_t *__restrict__ pos, int n, int32_t *__restrict__ nodes) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && nf[i]) nodes[pos[i]] = i;
}
// parent node of every node: the first node cell met walking down the drain table (:1893-1908); child counts
__global__ void k_sn_parent(const int32_t *__restrict__ b, const int32_t *__restrict__ nf, const int32_t *__restrict__ pos,
                            const int32_t *__restrict__ nodes, int nn, int n, int32_t *__restrict__ par, int32_t *__restrict__ cc) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nn) return;
    const int cell = nodes[k];
    if (cell == n - 1) { par[k] = -1; return; }
    int d = n - DRAIN(b, cell);
    while (nf[d] == 0) d = n - DRAIN(b, d);
    const int p = pos[d];
    par[k] = p;
    atomicAdd(cc + p, 1);
}
// children lists (CSR): filled in any order, then each list is put in DESCENDING node index = descending cell order, the
// order in which the reference's scan `do j=1,posi-1: posi-j` meets them
__global__ void k_sn_fill(const int32_t *__restrict__ par, const int32_t *__restrict__ cstart, int nn, int32_t *__restrict__ fillp,
                          int32_t *__restrict__ child) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nn || par[k] < 0) return;
    const int p = par[k];
    child[cstart[p] + atomicAdd(fillp + p, 1)] = k;
}
__global__ void k_sn_sortlists(const int32_t *__restrict__ cstart, const int32_t *__restrict__ cc, int nn, int32_t *__restrict__ child) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nn) return;
    int32_t *l = child + cstart[p];
    const int m = cc[p];
    for (int a = 1; a < m; ++a) {                       // insertion sort, lists hold a handful of tributaries
        const int v = l[a];
        int q = a - 1;
        while (q >= 0 && l[q] < v) { l[q + 1] = l[q]; --q; }
        l[q + 1] = v;
    }
}
// breadth-first numbering by ONE CTA: disc[k] = discovery number (outlet 1); frontier buffers fa / fb hold node indices
__global__ void __launch_bounds__(1024) k_sn_bfs(const int32_t *__restrict__ cstart, const int32_t *__restrict__ cc,
                                                 const int32_t *__restrict__ child, int nn, int32_t *__restrict__ disc,
                                                 int32_t *__restrict__ fa, int32_t *__restrict__ fb) {
    __shared__ int sm[33];
    __shared__ int s_carry;
    int32_t *cur = fa, *nxt = fb;
    int ncur = 1, found = 1;
    if (threadIdx.x == 0) { cur[0] = nn - 1; disc[nn - 1] = 1; }     // the outlet is the last node of the ascending list
    __syncthreads();
    while (ncur > 0) {
        if (threadIdx.x == 0) s_carry = 0;
        __syncthreads();
        for (int base = 0; base < ncur; base += blockDim.x) {
            const int j = base + threadIdx.x;
            const int node = (j < ncur) ? cur[j] : -1;
            const int m = (node >= 0) ? cc[node] : 0;
            int total;
            const int off = block_scan_excl(m, sm, &total) + s_carry;
            if (node >= 0) {
                const int32_t *l = child + cstart[node];
                for (int q = 0; q < m; ++q) {
                    nxt[off + q] = l[q];
                    disc[l[q]] = found + off + q + 1;
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) s_carry += total;
            __syncthreads();
        }
        const int nnext = s_carry;
        __syncthreads();
        found += nnext;
        ncur = nnext;
        int32_t *t = cur; cur = nxt; nxt = t;
        __threadfence_block();
        __syncthreads();
    }
}
// outputs: nodos_fin(cell) = discovery number; sub_basins(1, nn-disc+1) = cell (1-based), sub_basins(2, ..) = parent's number
__global__ void k_sn_emit(const int32_t *__restrict__ nodes, const int32_t *__restrict__ par, const int32_t *__restrict__ disc, int nn,
                          int32_t *__restrict__ nf, int32_t *__restrict__ sb) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nn) return;
    const int dnum = disc[k];
    nf[nodes[k]] = dnum;
    const int col = nn - dnum;      // 0-based column
    sb[2 * (size_t)col] = nodes[k] + 1;
    sb[2 * (size_t)col + 1] = par[k] >= 0 ? disc[par[k]] : 0;
}

// ---------------------------------------------------------------- basin_subbasin_find (:1979-2010)
__global__ void k_sub_find(const int32_t *__restrict__ b, const int32_t *__restrict__ nodos, int n, int32_t *__restrict__ sub_pert) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int v = nodos[i];
    if (v == 0) {
        int d = n - DRAIN(b, i);
        while (d < n && nodos[d] == 0) d = n - DRAIN(b, d);
        v = d < n ? nodos[d] : 0;
    }
    sub_pert[i] = v;
}

// ---------------------------------------------------------------- per-hill reductions in the reference's summation order
// key of cell i = hill id when the cell takes part (channel cell of a hill), else 0; cells gathered per hill, ascending
__global__ void k_hill_keys(const int32_t *__restrict__ sub_pert, const int32_t *__restrict__ cauce, int n, int nn,
                            uint32_t *__restrict__ key, int32_t *__restrict__ val, int32_t *__restrict__ hist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool on = false;
    int h = 0;
    if (i < n) {
        h = sub_pert[i];
        on = (!cauce || cauce[i] == 1) && h >= 1 && h <= nn;
        key[i] = on ? (uint32_t)h : 0u;
        val[i] = i;
        if (on) atomicAdd(hist + h, 1);
    }
    const unsigned off = __ballot_sync(0xffffffffu, i < n && !on);      // key 0: one atomic per warp
    if ((threadIdx.x & 31) == 0 && off) atomicAdd(hist, __popc(off));
}
// mode 0: sum of (float)(int)v  [basin_subbasin_long: INTEGER dummy `long`]   1: sum of v   2: max of v
__global__ void k_hill_reduce(const int32_t *__restrict__ order, const int32_t *__restrict__ hstart, const int32_t *__restrict__ hist,
                              const float *__restrict__ v, int nn, int mode, float *__restrict__ out) {
    const int h = blockIdx.x * blockDim.x + threadIdx.x + 1;     // hill ids are 1-based
    if (h > nn) return;
    const int m = hist[h];
    const int32_t *l = order + hstart[h];
    float s = (mode == 2) ? -3.4028235e38f : 0.0f;
    for (int q = 0; q < m; ++q) {
        const float x = v[l[q]];
        if (mode == 0) s = s + (float)(int32_t)x;
        else if (mode == 1) s = s + x;
        else s = fmaxf(s, x);
    }
    out[h - 1] = s;
}

struct HillGroups { int32_t *order, *hstart, *hist; };

int hill_groups(wmf_ctx *ctx, Scope &sc, const int32_t *d_sp, const int32_t *d_ca, int n, int nn, HillGroups *g) {
    uint32_t *key, *key2;
    int32_t *val;
    WMF_TRY(sc.alloc(&key, (size_t)n)); WMF_TRY(sc.alloc(&key2, (size_t)n));
    WMF_TRY(sc.alloc(&val, (size_t)n)); WMF_TRY(sc.alloc(&g->order, (size_t)n));
    WMF_TRY(sc.alloc(&g->hist, (size_t)nn + 2)); WMF_TRY(sc.alloc(&g->hstart, (size_t)nn + 2));
    CU_TRY(ctx, cudaMemsetAsync(g->hist, 0, sizeof(int32_t) * ((size_t)nn + 2), ctx->stream));
    LAUNCH(ctx, k_hill_keys, grid_for(n, 256), 256, 0, d_sp, d_ca, n, nn, key, val, g->hist);
    // stable LSD radix sort: the cells of one hill stay in ascending cell order; key 0 (not taking part) sorts first
    size_t tmp_bytes = 0;
    int bits = 1;
    while ((1ll << bits) <= nn) ++bits;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key, key2, val, g->order, n, 0, bits, ctx->stream);
    void *tmp;
    WMF_TRY(sc.alloc(&tmp, tmp_bytes));
    CU_TRY(ctx, cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, key, key2, val, g->order, n, 0, bits, ctx->stream));
    ctx->launches += 3;
    return wmf_device_exclusive_scan_i32(ctx, sc, g->hist, g->hstart, (int64_t)nn + 1, nullptr);   // hist[0] = cells with key 0
}

// ---------------------------------------------------------------- basin_subbasin_horton (:1931-1978)
// children of the node numbered p = the columns j with sub_basins(2,j) == p; one CTA relaxes the tree from the leaves
__global__ void k_hz_count(const int32_t *__restrict__ sb, int nn, int32_t *__restrict__ cc) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nn) return;
    const int p = sb[2 * (size_t)j + 1];
    if (p >= 1 && p <= nn) atomicAdd(cc + (nn - p), 1);           // column of the node numbered p (0-based)
}
__global__ void k_hz_fill(const int32_t *__restrict__ sb, const int32_t *__restrict__ cstart, int nn, int32_t *__restrict__ fillp,
                          int32_t *__restrict__ child) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nn) return;
    const int p = sb[2 * (size_t)j + 1];
    if (p < 1 || p > nn) return;
    const int c = nn - p;
    child[cstart[c] + atomicAdd(fillp + c, 1)] = j;
}
__global__ void k_hz_sort_asc(const int32_t *__restrict__ cstart, const int32_t *__restrict__ cc, int nn, int32_t *__restrict__ child) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nn) return;
    int32_t *l = child + cstart[p];
    const int m = cc[p];
    for (int a = 1; a < m; ++a) {                       // ascending column = the order of the reference's scan j = 1..i-1
        const int v = l[a];
        int q = a - 1;
        while (q >= 0 && l[q] > v) { l[q + 1] = l[q]; --q; }
        l[q + 1] = v;
    }
}
// one relaxation sweep over all nodes: a node whose (first ten) tributaries are all ordered gets its own order; the host
// repeats the sweep until nothing changes (as many sweeps as the node tree is high)
__global__ void k_hz_leaves(const int32_t *__restrict__ sb, const int32_t *__restrict__ cc, int nn, int32_t *__restrict__ sh,
                            int32_t *__restrict__ nod_horton) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < nn && cc[j] == 0) { sh[j] = 1; nod_horton[sb[2 * (size_t)j] - 1] = 1; }     // no tributary: order 1 (:1942-1949)
}
__global__ void k_hz_sweep(const int32_t *__restrict__ sb, const int32_t *__restrict__ cstart, const int32_t *__restrict__ cc,
                           const int32_t *__restrict__ child, int nn, const int32_t *__restrict__ sh_in, int32_t *__restrict__ sh_out,
                           int32_t *__restrict__ nod_horton, int32_t *__restrict__ changed) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nn) return;
    const int cur = sh_in[j];
    if (cur != 0) { sh_out[j] = cur; return; }
    const int32_t *l = child + cstart[j];
    // the reference looks only at tributaries in columns below its own (j < i) and keeps at most ten (drenantes(10))
    int cont = 0, mx = 0, mn = 0;
    for (int q = 0; q < cc[j] && cont < 10; ++q) {
        if (l[q] >= j) break;
        const int o = sh_in[l[q]];
        if (o == 0) { sh_out[j] = 0; return; }
        mx = cont ? max(mx, o) : o;
        mn = cont ? min(mn, o) : o;
        ++cont;
    }
    const int o = (mx == mn) ? mx + 1 : mx;
    sh_out[j] = o;
    nod_horton[sb[2 * (size_t)j] - 1] = o;
    *changed = 1;
}

// ---------------------------------------------------------------- basin_subbasin_long, longest path (:2033-2050)
__global__ void k_long_path(const int32_t *__restrict__ sb, const int32_t *__restrict__ sh, const float *__restrict__ hl, int nn,
                            unsigned long long *__restrict__ best) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;     // 0-based column
    if (i >= nn || sh[i] != 1) return;
    float acc = hl[i];
    int d = nn - sb[2 * (size_t)i + 1];                      // 0-based column of the node it drains to; == nn when that is 0
    while (d >= 0 && d < nn && sb[2 * (size_t)d + 1] >= 1) {
        acc = acc + hl[d];
        d = nn - sb[2 * (size_t)d + 1];
    }
    if (acc > 0.0f)      // strictly greater than the running maximum, first column wins ties: max over (value, -column)
        atomicMax(best, ((unsigned long long)__float_as_uint(acc) << 32) | (unsigned)(0x7fffffff - i));
}
// sub_basin_long is indexed by column i <-> hill id posi = nn-i+1: reverse the per-hill sums
__global__ void k_reverse(const float *__restrict__ in, int nn, float *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nn) out[i] = in[nn - 1 - i];
}
__global__ void k_prop_final(const float *__restrict__ sum_slope, const int32_t *__restrict__ hist, int nn, float *__restrict__ slope) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nn) slope[i] = sum_slope[i] / (float)hist[i + 1];      // 0/0 = NaN for a hill without channel cells, as the Fortran
}
__global__ void k_m2s_final(const float *__restrict__ red, const int32_t *__restrict__ hist, int nn, int sum_mean, float *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;     // column i <-> hill nn-i
    if (i >= nn) return;
    const int h = nn - i;
    const int m = hist[h];
    float r = 0.0f;
    if (m > 0) r = (sum_mean == 0) ? red[h - 1] / (float)m : red[h - 1];
    out[i] = r;
}

// ---------------------------------------------------------------- basin_stream_slope (:1380-1438)
struct Reach { float S, L; int32_t cont, second; };
// one thread per start node (nodos > 0, not the outlet): slope and (INTEGER) length of the reach down to the next node
__global__ void k_ss_walk(const int32_t *__restrict__ b, const float *__restrict__ elev, const float *__restrict__ lng,
                          const int32_t *__restrict__ nodos, int n, float dxp, Reach *__restrict__ reach, int32_t *__restrict__ lastsec,
                          int32_t *__restrict__ owner) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    lastsec[i] = -1;
    if (!(nodos[i] > 0 && DRAIN(b, i) != 0)) return;
    atomicMax(owner + i, i);               // the reach's cells are claimed on the way down: the largest start node wins
    const float e0 = elev[i], d0 = lng[i];
    float e = e0, dist = d0;
    int cont = 1, second = -1;
    int d = n - DRAIN(b, i);
    for (;;) {
        if (nodos[d] == 0) {
            ++cont;
            e = elev[d];
            dist = dist + lng[d];
            if (cont == 2) second = d;
            atomicMax(owner + d, i);
        } else {
            if (cont == 1) { cont = 2; e = elev[d]; dist = dist + lng[d]; }    // celdas_tramo(2) is NOT written here
            break;
        }
        d = n - DRAIN(b, d);
    }
    float S = fabsf((e0 - e) / (d0 - dist));
    if (S <= 0.0f) S = 0.001f;
    int32_t Li = (int32_t)dist;            // L_tramo is implicitly INTEGER
    if (Li <= 0) Li = (int32_t)dxp;
    reach[i] = Reach{S, (float)Li, cont, second};
    if (second >= 0) lastsec[i] = i;       // this reach leaves a cell in celdas_tramo(2)
}
// when the cell below a start node is itself a node, the reference's assignment loop also writes the cell the latest earlier
// reach left in celdas_tramo(2): one more claim
__global__ void k_ss_stale(const int32_t *__restrict__ b, const int32_t *__restrict__ nodos, const Reach *__restrict__ reach,
                           const int32_t *__restrict__ lastsec_scan, int n, int32_t *__restrict__ owner) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || i == 0) return;
    if (!(nodos[i] > 0 && DRAIN(b, i) != 0) || reach[i].second >= 0) return;
    const int prev = lastsec_scan[i - 1];
    if (prev >= 0) atomicMax(owner + reach[prev].second, i);
}
// every claimed cell takes the slope and length of the reach that owns it
__global__ void k_ss_write(const int32_t *__restrict__ owner, const Reach *__restrict__ reach, int n, float *__restrict__ ss,
                           float *__restrict__ sl) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const int o = owner[c];
    if (o >= 0) { ss[c] = reach[o].S; sl[c] = reach[o].L; }
}
__global__ void k_fill2(float *a, float *b2, float v, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { a[i] = v; b2[i] = v; }
}

// stream_find_to_corr, cuencas.f90:372-417: one thread walks down DIR to the first cell of the channel map
__global__ void k_find_to_corr(const int32_t *__restrict__ dir, const int32_t *__restrict__ cauce, int nc, int nf, int col, int fil,
                               wmf_geo g, float *__restrict__ out) {
    float lat = 0.0f, lon = 0.0f;
    if (col >= 1 && col <= nc && fil >= 1 && fil <= nf) {
        int dire = dir[(size_t)(fil - 1) * nc + (col - 1)];
        bool go = true;
        while (go) {
            for (int kf = 1; kf <= 3; ++kf)
                for (int kc = 1; kc <= 3; ++kc)
                    if (dire == 9 - 3 * kf + kc) {      // `dire` changes inside the scan: a sweep can take several steps
                        col = col + kc - 2;
                        fil = fil + kf - 2;
                        dire = (col < 1 || col > nc || fil < 1 || fil > nf) ? -1 : dir[(size_t)(fil - 1) * nc + (col - 1)];
                        lat = g.xll + g.dx * ((float)col - 0.5f);
                        lon = g.yll + g.dx * ((float)(g.nrows - fil) + 0.5f);
                    }
            if (col > nc || fil > nf || col <= 0 || fil <= 0 || (float)dire == g.nodata || dire <= 0 ||
                cauce[(size_t)(fil - 1) * nc + (col - 1)] != 0)
                go = false;
        }
    }
    out[0] = lat;
    out[1] = lon;
}

// find_xy_in_basin (cuencas.f90:2659-2678) for a batch of points: posit[j] = smallest 1-based index of the cell at (col,fil)
__global__ void k_find_xy(const int32_t *__restrict__ b, int n, const int32_t *__restrict__ cf, int npts, int32_t *__restrict__ posit) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int col = b[3 * (size_t)i + 1], fil = b[3 * (size_t)i + 2];
    for (int j = 0; j < npts; ++j)
        if (cf[2 * j] == col && cf[2 * j + 1] == fil) atomicMin(posit + j, i + 1);
}
// basin_point2var (:1230-1261) / basin_stream_point2stream (:1524-1577): the points are handled in order by one thread, as
// the Fortran does (a later point overwrites an earlier one that lands on the same cell)
__global__ void k_points(const int32_t *__restrict__ b, const int32_t *__restrict__ cauce, const int32_t *__restrict__ ids,
                         const int32_t *__restrict__ posit, int npts, int n, wmf_geo g, int32_t *__restrict__ res,
                         int32_t *__restrict__ pts, float *__restrict__ xy_new) {
    if (threadIdx.x || blockIdx.x) return;
    for (int i = 0; i < npts; ++i) {
        int p = posit[i] <= n ? posit[i] : 0;
        if (!cauce) {                       // basin_point2var
            if (p > 0) { res[i] = 0; pts[p - 1] = ids[i]; }
            else res[i] = 1;
            continue;
        }
        int r = 0;
        if (p > 0) {
            r = 1;
            if (cauce[p - 1] > 0) { r = 2; pts[p - 1] = ids[i]; }
            else {
                while (cauce[p - 1] == 0 && p < n) p = n + 1 - b[3 * (size_t)(p - 1)];
                if (p < n) { pts[p - 1] = ids[i]; r = 2; }
            }
        }
        res[i] = r;
        if (r == 2) {
            xy_new[2 * i] = g.xll + g.dx * ((float)b[3 * (size_t)(p - 1) + 1] - 0.5f);
            xy_new[2 * i + 1] = g.yll + g.dx * ((float)(g.nrows - b[3 * (size_t)(p - 1) + 2]) + 0.5f);
        } else {
            xy_new[2 * i] = g.nodata;
            xy_new[2 * i + 1] = g.nodata;
        }
    }
}
// column / row of every point: x_col = ceiling((x-xll)/dx), y_fil = ceiling(nrows - (y-yll)/dx), in fp32 as the Fortran
__global__ void k_point_colfil(const float *__restrict__ xy, int npts, wmf_geo g, int32_t *__restrict__ cf, int32_t *__restrict__ posit) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= npts) return;
    cf[2 * j] = (int32_t)ceilf((xy[2 * j] - g.xll) / g.dx);
    cf[2 * j + 1] = (int32_t)ceilf((float)g.nrows - (xy[2 * j + 1] - g.yll) / g.dx);
    posit[j] = 0x7fffffff;
}

// basin_stream_sections, cuencas.f90:1464-1523: raster lookup of the basin positions (first match = smallest index), then one
// thread per (cell, sample of the cross section)
__global__ void k_sec_lookup(const int32_t *__restrict__ b, int n, int nc, int nr, int32_t *__restrict__ look) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = b[3 * (size_t)i + 1], f = b[3 * (size_t)i + 2];
    if (c >= 1 && c <= nc && f >= 1 && f <= nr) atomicMin(look + (size_t)(f - 1) * nc + (c - 1), i + 1);
}
__global__ void k_sections(const int32_t *__restrict__ b, const int32_t *__restrict__ cauces, const int32_t *__restrict__ dirs,
                           const float *__restrict__ dem, const int32_t *__restrict__ look, int num, int n, int nc, int nr,
                           float *__restrict__ sec, int32_t *__restrict__ cel) {
    const int S = 2 * num + 1;
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (long long)n * S) return;
    const int i = (int)(g / S), k = (int)(g % S);
    float e = 0.0f;
    int p = 0;
    if (cauces[i] == 1) {
        int colMov = 0, filMov = 0;
        const int d = dirs[i];
        if (d == 4 || d == 6) filMov = 1;
        else if (d == 8 || d == 2) colMov = 1;
        else if (d == 7 || d == 3) { colMov = 1; filMov = -1; }
        else if (d == 1 || d == 9) { colMov = 1; filMov = 1; }
        const int j = k - num;
        const int c = b[3 * (size_t)i + 1] + j * colMov, f = b[3 * (size_t)i + 2] + j * filMov;
        if (c >= 1 && c <= nc && f >= 1 && f <= nr) {
            e = dem[(size_t)(f - 1) * nc + (c - 1)];
            const int l = look[(size_t)(f - 1) * nc + (c - 1)];
            p = (l == 0x7f7f7f7f) ? 0 : l;
        } else {
            e = -9999.0f;
            p = -9999;
        }
    }
    sec[g] = e;
    cel[g] = p;
}

int points_run(wmf_ctx *ctx, const int32_t *basin_f, const int32_t *cauce, const int32_t *id_coord, const float *xy_coord, int32_t ncoord,
               int32_t n, int32_t *res_coord, int32_t *basin_pts, float *xy_new, bool stream) {
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    Scope sc(ctx);
    const int32_t *d_b, *d_ca = nullptr, *d_id;
    const float *d_xy;
    int32_t *d_res, *d_pts, *cf, *posit;
    float *d_new = nullptr;
    WMF_TRY(sc.in(basin_f, (size_t)3 * n, &d_b));
    if (stream) WMF_TRY(sc.in(cauce, (size_t)n, &d_ca));
    WMF_TRY(sc.in(id_coord, (size_t)ncoord, &d_id));
    WMF_TRY(sc.in(xy_coord, (size_t)2 * ncoord, &d_xy));
    WMF_TRY(sc.out(res_coord, (size_t)ncoord, &d_res));
    WMF_TRY(sc.out(basin_pts, (size_t)n, &d_pts));
    if (stream) WMF_TRY(sc.out(xy_new, (size_t)2 * ncoord, &d_new));
    if (stream && !d_new) WMF_TRY(sc.alloc(&d_new, (size_t)2 * ncoord));
    WMF_TRY(sc.alloc(&cf, (size_t)2 * ncoord));
    WMF_TRY(sc.alloc(&posit, (size_t)ncoord));
    CU_TRY(ctx, cudaMemsetAsync(d_pts, 0, sizeof(int32_t) * (size_t)n, ctx->stream));
    LAUNCH(ctx, k_point_colfil, grid_for(ncoord, 128), 128, 0, d_xy, ncoord, ctx->geo, cf, posit);
    LAUNCH(ctx, k_find_xy, grid_for(n, 256), 256, 0, d_b, n, cf, ncoord, posit);
    LAUNCH(ctx, k_points, 1, 32, 0, d_b, d_ca, d_id, posit, ncoord, n, ctx->geo, d_res, d_pts, d_new);
    return sc.finish();
}

}  // namespace

int wmf_hill_groups(wmf_ctx *ctx, Scope &sc, const int32_t *sub_pert, const int32_t *cauce, int n, int n_hills, int32_t **order,
                    int32_t **hstart, int32_t **hist) {
    HillGroups g;
    WMF_TRY(hill_groups(ctx, sc, sub_pert, cauce, n, n_hills, &g));
    *order = g.order; *hstart = g.hstart; *hist = g.hist;
    return WMF_OK;
}

extern "C" {

int wmf_stream_find_to_corr(wmf_ctx *ctx, float x, float y, const int32_t *dir, const int32_t *cauce, int32_t nc, int32_t nf,
                            float *lat, float *lon) {
    if (!ctx || !dir || !cauce || !lat || !lon || nc <= 0 || nf <= 0) return WMF_ERR_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    Scope sc(ctx);
    const int32_t *d_d, *d_c;
    float *d_o, h[2];
    WMF_TRY(sc.in(dir, (size_t)nc * nf, &d_d));
    WMF_TRY(sc.in(cauce, (size_t)nc * nf, &d_c));
    WMF_TRY(sc.alloc(&d_o, 2));
    int32_t col, fil;
    WMF_TRY(wmf_coord2fil_col(ctx, x, y, &col, &fil));
    LAUNCH(ctx, k_find_to_corr, 1, 1, 0, d_d, d_c, nc, nf, col, fil, ctx->geo, d_o);
    CU_TRY(ctx, cudaMemcpyAsync(h, d_o, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    WMF_TRY(sc.finish());
    *lat = h[0];
    *lon = h[1];
    return WMF_OK;
}


int wmf_basin_subbasin_nod(wmf_ctx *ctx, const int32_t *basin_f, const int32_t *acum, int32_t n, int32_t umbral, int32_t *cauce,
                           int32_t *nodos_fin, int32_t *n_nodos) {
    if (!ctx || !basin_f || !acum || !cauce || !nodos_fin || !n_nodos || n <= 0) return WMF_ERR_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    Scope sc(ctx);
    const int32_t *d_b, *d_a;
    int32_t *d_c, *d_nf, *cnt, *pos;
    WMF_TRY(sc.in(basin_f, (size_t)3 * n, &d_b));
    WMF_TRY(sc.in(acum, (size_t)n, &d_a));
    WMF_TRY(sc.out(cauce, (size_t)n, &d_c));
    WMF_TRY(sc.out(nodos_fin, (size_t)n, &d_nf));
    WMF_TRY(sc.alloc(&cnt, (size_t)n)); WMF_TRY(sc.alloc(&pos, (size_t)n));
    CU_TRY(ctx, cudaMemsetAsync(cnt, 0, sizeof(int32_t) * (size_t)n, ctx->stream));
    const int G = grid_for(n, 256);
    LAUNCH(ctx, k_sn_count, G, 256, 0, d_b, d_a, n, umbral, d_c, cnt);
    LAUNCH(ctx, k_sn_flag, G, 256, 0, d_b, d_c, cnt, n, d_nf);
    int64_t nn64 = 0;
    WMF_TRY(wmf_device_exclusive_scan_i32(ctx, sc, d_nf, pos, n, &nn64));
    const int nn = (int)nn64;
    int32_t *nodes, *par, *cc, *cstart, *fillp, *child, *disc, *fa, *fb;
    WMF_TRY(sc.alloc(&nodes, (size_t)nn)); WMF_TRY(sc.alloc(&par, (size_t)nn)); WMF_TRY(sc.alloc(&cc, (size_t)nn + 1));
    WMF_TRY(sc.alloc(&cstart, (size_t)nn + 1)); WMF_TRY(sc.alloc(&fillp, (size_t)nn)); WMF_TRY(sc.alloc(&child, (size_t)nn));
    WMF_TRY(sc.alloc(&disc, (size_t)nn)); WMF_TRY(sc.alloc(&fa, (size_t)nn)); WMF_TRY(sc.alloc(&fb, (size_t)nn));
    CU_TRY(ctx, cudaMemsetAsync(cc, 0, sizeof(int32_t) * ((size_t)nn + 1), ctx->stream));
    CU_TRY(ctx, cudaMemsetAsync(fillp, 0, sizeof(int32_t) * (size_t)nn, ctx->stream));
    CU_TRY(ctx, cudaMemsetAsync(disc, 0, sizeof(int32_t) * (size_t)nn, ctx->stream));
    const int GN = grid_for(nn, 256);
    LAUNCH(ctx, k_sn_compact, G, 256, 0, d_nf, pos, n, nodes);
    LAUNCH(ctx, k_sn_parent, GN, 256, 0, d_b, d_nf, pos, nodes, nn, n, par, cc);
    WMF_TRY(wmf_device_exclusive_scan_i32(ctx, sc, cc, cstart, (int64_t)nn + 1, nullptr));
    LAUNCH(ctx, k_sn_fill, GN, 256, 0, par, cstart, nn, fillp, child);
    LAUNCH(ctx, k_sn_sortlists, GN, 256, 0, cstart, cc, nn, child);
    LAUNCH(ctx, k_sn_bfs, 1, 1024, 0, cstart, cc, child, nn, disc, fa, fb);
    if (ctx->sub_basins) { cudaFree(ctx->sub_basins); ctx->sub_basins = nullptr; }
    CU_TRY(ctx, cudaMalloc(&ctx->sub_basins, sizeof(int32_t) * 2 * (size_t)nn));
    ctx->sub_n = nn;
    LAUNCH(ctx, k_sn_emit, GN, 256, 0, nodes, par, disc, nn, d_nf, ctx->sub_basins);
    *n_nodos = nn;
    return sc.finish();
}

int wmf_basin_subbasin_cut(wmf_ctx *ctx, int32_t n_nodos, int32_t *sub_basins) {
    if (!ctx || !sub_basins || n_nodos <= 0) return WMF_ERR_ARG;
    if (!ctx->sub_basins || ctx->sub_n != n_nodos)
        WMF_FAIL(ctx, WMF_ERR_ARG, "basin_subbasin_cut: basin_subbasin_nod has not been run for %d nodes", n_nodos);
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    Scope sc(ctx);
    int32_t *d;
    WMF_TRY(sc.out(sub_basins, (size_t)2 * n_nodos, &d));
    CU_TRY(ctx, cudaMemcpyAsync(d, ctx->sub_basins, sizeof(int32_t) * 2 * (size_t)n_nodos, cudaMemcpyDeviceToDevice, ctx->stream));
    WMF_TRY(sc.finish());
    cudaFree(ctx->sub_basins);              // the Fortran deallocates sub_basins_temp here (:1929)
    ctx->sub_basins = nullptr;
    ctx->sub_n = 0;
    return WMF_OK;
}

int wmf_basin_subbasin_find(wmf_ctx *ctx, const int32_t *basin_f, const int32_t *nodos, int32_t n, int32_t *sub_pert) {
    if (!ctx || !basin_f || !nodos || !sub_pert || n <= 0) return WMF_ERR_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    Scope sc(ctx);
    const int32_t *d_b, *d_n;
    int32_t *d_s;
    WMF_TRY(sc.in(basin_f, (size_t)3 * n, &d_b));
    WMF_TRY(sc.in(nodos, (size_t)n, &d_n));
    WMF_TRY(sc.out(sub_pert, (size_t)n, &d_s));
    LAUNCH(ctx, k_sub_find, grid_for(n, 256), 256, 0, d_b, d_n, n, d_s);
    return sc.finish();
}

int wmf_basin_subbasin_horton(wmf_ctx *ctx, const int32_t *sub_basins, int32_t n_nodos, int32_t n, int32_t *sub_horton,
                              int32_t *nod_horton) {
    if (!ctx || !sub_basins || !sub_horton || !nod_horton || n_nodos <= 0 || n <= 0) return WMF_ERR_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    Scope sc(ctx);
    const int32_t *d_sb;
    int32_t *d_sh, *d_nh, *cc, *cstart, *fillp, *child;
    const int nn = n_nodos;
    WMF_TRY(sc.in(sub_basins, (size_t)2 * nn, &d_sb));
    WMF_TRY(sc.out(sub_horton, (size_t)nn, &d_sh));
    WMF_TRY(sc.out(nod_horton, (size_t)n, &d_nh));
    WMF_TRY(sc.alloc(&cc, (size_t)nn + 1)); WMF_TRY(sc.alloc(&cstart, (size_t)nn + 1));
    WMF_TRY(sc.alloc(&fillp, (size_t)nn)); WMF_TRY(sc.alloc(&child, (size_t)nn));
    CU_TRY(ctx, cudaMemsetAsync(cc, 0, sizeof(int32_t) * ((size_t)nn + 1), ctx->stream));
    CU_TRY(ctx, cudaMemsetAsync(fillp, 0, sizeof(int32_t) * (size_t)nn, ctx->stream));
    CU_TRY(ctx, cudaMemsetAsync(d_sh, 0, sizeof(int32_t) * (size_t)nn, ctx->stream));
    CU_TRY(ctx, cudaMemsetAsync(d_nh, 0, sizeof(int32_t) * (size_t)n, ctx->stream));
    const int GN = grid_for(nn, 256);
    LAUNCH(ctx, k_hz_count, GN, 256, 0, d_sb, nn, cc);
    WMF_TRY(wmf_device_exclusive_scan_i32(ctx, sc, cc, cstart, (int64_t)nn + 1, nullptr));
    LAUNCH(ctx, k_hz_fill, GN, 256, 0, d_sb, cstart, nn, fillp, child);
    LAUNCH(ctx, k_hz_sort_asc, GN, 256, 0, cstart, cc, nn, child);
    LAUNCH(ctx, k_hz_leaves, GN, 256, 0, d_sb, cc, nn, d_sh, d_nh);
    int32_t *sh2, *d_changed;
    WMF_TRY(sc.alloc(&sh2, (size_t)nn));
    WMF_TRY(sc.alloc(&d_changed, 1));
    int32_t *a = d_sh, *b2 = sh2;
    for (int it = 0; it < nn + 2; it += 8) {            // eight sweeps between two looks at the flag
        CU_TRY(ctx, cudaMemsetAsync(d_changed, 0, sizeof(int32_t), ctx->stream));
        for (int k = 0; k < 8; ++k) {
            LAUNCH(ctx, k_hz_sweep, GN, 256, 0, d_sb, cstart, cc, child, nn, a, b2, d_nh, d_changed);
            int32_t *t = a; a = b2; b2 = t;
        }
        int32_t h_changed = 0;
        CU_TRY(ctx, cudaMemcpyAsync(&h_changed, d_changed, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        if (!h_changed) break;
    }
    // an even number of sweeps per round: the current orders are back in d_sh
    return sc.finish();
}

int wmf_basin_subbasin_long(wmf_ctx *ctx, const int32_t *sub_pert, const int32_t *cauce, const float *lng, const int32_t *sub_basin,
                            const int32_t *sub_horton, int32_t n_nodos, int32_t n, float *sub_basin_long, float *max_long,
                            int32_t *nodo_max_long) {
    if (!ctx || !sub_pert || !cauce || !lng || !sub_basin || !sub_horton || !sub_basin_long || !max_long || !nodo_max_long ||
        n_nodos <= 0 || n <= 0)
        return WMF_ERR_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    Scope sc(ctx);
    const int nn = n_nodos;
    const int32_t *d_sp, *d_ca, *d_sb, *d_sh;
    const float *d_l;
    float *d_out, *byhill;
    WMF_TRY(sc.in(sub_pert, (size_t)n, &d_sp)); WMF_TRY(sc.in(cauce, (size_t)n, &d_ca)); WMF_TRY(sc.in(lng, (size_t)n, &d_l));
    WMF_TRY(sc.in(sub_basin, (size_t)2 * nn, &d_sb)); WMF_TRY(sc.in(sub_horton, (size_t)nn, &d_sh));
    WMF_TRY(sc.out(sub_basin_long, (size_t)nn, &d_out));
    WMF_TRY(sc.alloc(&byhill, (size_t)nn));
    HillGroups g;
    WMF_TRY(hill_groups(ctx, sc, d_sp, d_ca, n, nn, &g));
    const int GN = grid_for(nn, 256);
    LAUNCH(ctx, k_hill_reduce, GN, 256, 0, g.order, g.hstart, g.hist, d_l, nn, 0, byhill);
    LAUNCH(ctx, k_reverse, GN, 256, 0, byhill, nn, d_out);
    unsigned long long *best, h_best = 0;
    WMF_TRY(sc.alloc(&best, 1));
    CU_TRY(ctx, cudaMemsetAsync(best, 0, sizeof(unsigned long long), ctx->stream));
    LAUNCH(ctx, k_long_path, GN, 256, 0, d_sb, d_sh, d_out, nn, best);
    CU_TRY(ctx, cudaMemcpyAsync(&h_best, best, sizeof h_best, cudaMemcpyDeviceToHost, ctx->stream));
    WMF_TRY(sc.finish());
    if (h_best == 0) { *max_long = 0.0f; *nodo_max_long = 0; }
    else {
        const uint32_t bits = (uint32_t)(h_best >> 32);
        memcpy(max_long, &bits, 4);
        const int col = 0x7fffffff - (int)(h_best & 0xffffffffu);
        *nodo_max_long = nn - col;          // n_nodos - i + 1 with i = col + 1
    }
    return WMF_OK;
}

int wmf_basin_subbasin_stream_prop(wmf_ctx *ctx, const int32_t *sub_pert, const int32_t *cauce, const float *lng, const float *slope,
                                   int32_t n, int32_t n_nodos, float *stream_slope, float *stream_long) {
    if (!ctx || !sub_pert || !cauce || !lng || !slope || !stream_slope || !stream_long || n_nodos <= 0 || n <= 0) return WMF_ERR_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    Scope sc(ctx);
    const int nn = n_nodos;
    const int32_t *d_sp, *d_ca;
    const float *d_l, *d_s;
    float *d_ss, *d_sl, *sum_s;
    WMF_TRY(sc.in(sub_pert, (size_t)n, &d_sp)); WMF_TRY(sc.in(cauce, (size_t)n, &d_ca));
    WMF_TRY(sc.in(lng, (size_t)n, &d_l)); WMF_TRY(sc.in(slope, (size_t)n, &d_s));
    WMF_TRY(sc.out(stream_slope, (size_t)nn, &d_ss)); WMF_TRY(sc.out(stream_long, (size_t)nn, &d_sl));
    WMF_TRY(sc.alloc(&sum_s, (size_t)nn));
    HillGroups g;
    WMF_TRY(hill_groups(ctx, sc, d_sp, d_ca, n, nn, &g));
    const int GN = grid_for(nn, 256);
    LAUNCH(ctx, k_hill_reduce, GN, 256, 0, g.order, g.hstart, g.hist, d_l, nn, 1, d_sl);       // indexed by hill id i (:2107)
    LAUNCH(ctx, k_hill_reduce, GN, 256, 0, g.order, g.hstart, g.hist, d_s, nn, 1, sum_s);
    LAUNCH(ctx, k_prop_final, GN, 256, 0, sum_s, g.hist, nn, d_ss);
    return sc.finish();
}

int wmf_basin_subbasin_map2subbasin(wmf_ctx *ctx, const int32_t *sub_pert, const float *basin_var, int32_t n_nodos, int32_t n,
                                    const int32_t *cauce, int32_t sum_mean, float *subbasin_sum) {
    if (!ctx || !sub_pert || !basin_var || !cauce || !subbasin_sum || n_nodos <= 0 || n <= 0 || sum_mean < 0 || sum_mean > 2)
        return WMF_ERR_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    Scope sc(ctx);
    const int nn = n_nodos;
    const int32_t *d_sp, *d_ca;
    const float *d_v;
    float *d_o, *red;
    WMF_TRY(sc.in(sub_pert, (size_t)n, &d_sp)); WMF_TRY(sc.in(cauce, (size_t)n, &d_ca)); WMF_TRY(sc.in(basin_var, (size_t)n, &d_v));
    WMF_TRY(sc.out(subbasin_sum, (size_t)nn, &d_o));
    WMF_TRY(sc.alloc(&red, (size_t)nn));
    HillGroups g;
    WMF_TRY(hill_groups(ctx, sc, d_sp, d_ca, n, nn, &g));
    const int GN = grid_for(nn, 256);
    LAUNCH(ctx, k_hill_reduce, GN, 256, 0, g.order, g.hstart, g.hist, d_v, nn, sum_mean == 2 ? 2 : 1, red);
    LAUNCH(ctx, k_m2s_final, GN, 256, 0, red, g.hist, nn, sum_mean, d_o);
    return sc.finish();
}

int wmf_basin_stream_slope(wmf_ctx *ctx, const int32_t *basin_f, const float *basin_elev, const float *basin_long, const int32_t *nodos,
                           int32_t n, float *stream_s, float *stream_l) {
    if (!ctx || !basin_f || !basin_elev || !basin_long || !nodos || !stream_s || !stream_l || n <= 0) return WMF_ERR_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    Scope sc(ctx);
    const int32_t *d_b, *d_n;
    const float *d_e, *d_l;
    float *d_ss, *d_sl;
    WMF_TRY(sc.in(basin_f, (size_t)3 * n, &d_b)); WMF_TRY(sc.in(nodos, (size_t)n, &d_n));
    WMF_TRY(sc.in(basin_elev, (size_t)n, &d_e)); WMF_TRY(sc.in(basin_long, (size_t)n, &d_l));
    WMF_TRY(sc.out(stream_s, (size_t)n, &d_ss)); WMF_TRY(sc.out(stream_l, (size_t)n, &d_sl));
    Reach *reach;
    int32_t *lastsec, *lastscan, *owner;
    WMF_TRY(sc.alloc(&reach, (size_t)n)); WMF_TRY(sc.alloc(&lastsec, (size_t)n)); WMF_TRY(sc.alloc(&lastscan, (size_t)n));
    WMF_TRY(sc.alloc(&owner, (size_t)n));
    CU_TRY(ctx, cudaMemsetAsync(owner, 0xff, sizeof(int32_t) * (size_t)n, ctx->stream));    // -1
    const int G = grid_for(n, 256);
    LAUNCH(ctx, k_fill2, G, 256, 0, d_ss, d_sl, ctx->geo.nodata, n);                        // stream_s = stream_l = noData (:1393-1394)
    LAUNCH(ctx, k_ss_walk, G, 256, 0, d_b, d_e, d_l, d_n, n, ctx->geo.dxp, reach, lastsec, owner);
    size_t tmp_bytes = 0;
    cub::DeviceScan::InclusiveScan(nullptr, tmp_bytes, lastsec, lastscan, cub::Max(), n, ctx->stream);
    void *tmp;
    WMF_TRY(sc.alloc(&tmp, tmp_bytes));
    CU_TRY(ctx, cub::DeviceScan::InclusiveScan(tmp, tmp_bytes, lastsec, lastscan, cub::Max(), n, ctx->stream));
    ctx->launches += 2;
    LAUNCH(ctx, k_ss_stale, G, 256, 0, d_b, d_n, reach, lastscan, n, owner);
    LAUNCH(ctx, k_ss_write, G, 256, 0, owner, reach, n, d_ss, d_sl);
    return sc.finish();
}

}  // extern "C"

extern "C" {

int wmf_basin_point2var(wmf_ctx *ctx, const int32_t *basin_f, const int32_t *id_coord, const float *xy_coord, int32_t ncoord,
                        int32_t nceldas, int32_t *res_coord, int32_t *basin_pts) {
    if (!ctx || !basin_f || !id_coord || !xy_coord || !res_coord || !basin_pts || ncoord <= 0 || nceldas <= 0) return WMF_ERR_ARG;
    return points_run(ctx, basin_f, nullptr, id_coord, xy_coord, ncoord, nceldas, res_coord, basin_pts, nullptr, false);
}

int wmf_basin_stream_point2stream(wmf_ctx *ctx, const int32_t *basin_f, const int32_t *cauce, const int32_t *id_coord,
                                  const float *xy_coord, int32_t ncoord, int32_t nceldas, int32_t *res_coord, int32_t *basin_pts,
                                  float *xy_new) {
    if (!ctx || !basin_f || !cauce || !id_coord || !xy_coord || !res_coord || !basin_pts || !xy_new || ncoord <= 0 || nceldas <= 0)
        return WMF_ERR_ARG;
    return points_run(ctx, basin_f, cauce, id_coord, xy_coord, ncoord, nceldas, res_coord, basin_pts, xy_new, true);
}

}  // extern "C"

extern "C" int wmf_basin_stream_sections(wmf_ctx *ctx, const int32_t *basin_f, const int32_t *cauces, const int32_t *directions,
                                         const float *dem, int32_t num_celdas, int32_t nceldas, int32_t ncols, int32_t nrows,
                                         float *secciones, int32_t *secciones_cel) {
    if (!ctx || !basin_f || !cauces || !directions || !dem || !secciones || !secciones_cel || num_celdas < 0 || nceldas <= 0 ||
        ncols <= 0 || nrows <= 0)
        return WMF_ERR_ARG;
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    Scope sc(ctx);
    const int S = 2 * num_celdas + 1, n = nceldas;
    const int32_t *d_b, *d_c, *d_d;
    const float *d_dem;
    float *d_s;
    int32_t *d_p, *look;
    WMF_TRY(sc.in(basin_f, (size_t)3 * n, &d_b)); WMF_TRY(sc.in(cauces, (size_t)n, &d_c)); WMF_TRY(sc.in(directions, (size_t)n, &d_d));
    WMF_TRY(sc.in(dem, (size_t)ncols * nrows, &d_dem));
    WMF_TRY(sc.out(secciones, (size_t)S * n, &d_s)); WMF_TRY(sc.out(secciones_cel, (size_t)S * n, &d_p));
    WMF_TRY(sc.alloc(&look, (size_t)ncols * nrows));
    CU_TRY(ctx, cudaMemsetAsync(look, 0x7f, sizeof(int32_t) * (size_t)ncols * nrows, ctx->stream));   // 0x7f7f7f7f: "no cell yet"
    LAUNCH(ctx, k_sec_lookup, grid_for(n, 256), 256, 0, d_b, n, ncols, nrows, look);
    LAUNCH(ctx, k_sections, grid_for((int64_t)n * S, 256), 256, 0, d_b, d_c, d_d, d_dem, look, num_celdas, n, ncols, nrows, d_s, d_p);
    return sc.finish();
}
