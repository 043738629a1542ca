// pathfind.cu -- basin_find (cuencas.f90:525-591) and the accumulation it implies (:659-680) WITHOUT a level-by-level
// search: every stage is a streaming pass over the raster or over the tile-boundary graph.
//
// The reference's FIFO queue visits the basin by level (hops to the outlet) and, inside a level, in the order of the DFS
// preorder taken with the children in the 3x3 probe order (kf outer, kc inner, :555-571).  With
//      size(v) = cells draining through v (== basin_acum, :659-680)
//      off(c)  = 1 + sum of size(s) over the siblings s of c probed before c
//      pre(v)  = pre(parent) + off(v),  depth(v) = depth(parent) + 1,   pre(outlet) = depth(outlet) = 0
// the queue position of v is its rank under (depth, pre), and basin_f(1,.) is the queue position of the parent.
// size / depth / pre are leaf-to-root and root-to-leaf sums on a tree that is ~20 000 levels deep for a 20000^2 raster, so
// they are computed by TILES: a 64x64 tile is solved in shared memory by pointer doubling (<= 12 rounds), the paths that
// leave a tile form a graph over the tiles' perimeter cells only (~2 % of the raster, a few hundred hops deep), which is
// solved by a barrier-free climb (sizes) and by pointer jumping (depth, pre), and a last pass adds the two.  The ranking is
// a scatter by preorder (a permutation: no histogram) followed by a stable LSD radix sort on the depth.
//
//   k_pf_local1   tile: local terminal of every perimeter cell, cells collected by every exit cell          (4 B/cell read)
//   k_pf_bg_*     boundary graph: next exit cell, global size of the exit cells ("last arriver climbs")
//   k_pf_local2   tile: global size of every cell (external inflows added), sibling offsets, local hops and
//                 preorder sums to the terminal                                                             (4 r + 12 w)
//   k_pf_exit_flag + scan, k_pf_bd_*
//                 the exit cells (a third of the perimeter slots) renumbered densely; their depth / preorder by pointer
//                 jumping (one CTA runs all the rounds of a raster up to 1024^2: k_pf_bd_jump_small)
//   k_pf_emit     cell: depth, preorder -> one (key, value) pair written at position preorder               (8 r + 8 w)
//   k_rs_*        stable 8-bit LSD radix passes on the depth (2 for depths < 65536): digit histograms per 4096-item tile,
//                 one device scan, the tile ranked with one match.any per item and sorted in shared memory before the
//                 copy-out (digit runs leave as consecutive addresses)                                      (20 B per pass)
//   k_pf_final_*  queue order: raster index, parent position (run-length expansion of the child counts), size
//
// Measured on B200 (profiles/README.md): 20000 x 20000 raster, 399.9 M cells, 19 998 levels: basin_find 25.5 ms (the
// level-by-level kernel of round 1, still behind WMF_FIND_ALGO=bfs: 73 ms), bit-exact against the reference's Fortran.
//
// oracle/tile_model.py restates the same decomposition on the CPU (tests/test_tile_model.py checks it against the
// reference's queue); tests/test_gpu_topo.py / test_gpu_fullsize.py compare this path with the oracle bit for bit.
#include <limits.h>
#include <stdlib.h>

#include <algorithm>
#include <string>

#include "common.cuh"

namespace pf {

constexpr int TS = 64;                 // tile edge
constexpr int TC = TS * TS;            // cells per tile
constexpr int NSLOT = 256;             // node slots per tile (252 perimeter cells)
constexpr int NPERIM = 4 * TS - 4;
constexpr int NTHR = 512;              // threads per tile CTA
constexpr int CPT = TC / NTHR;         // cells per thread
constexpr int MAXROUNDS = 12;          // 2^12 = longest possible path inside a tile
constexpr int ROOT_OUT = -1, DEAD = -2;
constexpr unsigned short NONE16 = 0xFFFF;
constexpr unsigned T_ROOT_OUT = 0xFFFEu, T_DEAD = 0xFFFFu;
constexpr int EW = TS + 2;             // tile + halo ring

__device__ __forceinline__ int perim_index(int ly, int lx) {
    if (ly == 0) return lx;
    if (ly == TS - 1) return TS + lx;
    if (lx == 0) return 2 * TS + ly - 1;
    return 2 * TS + (TS - 2) + ly - 1;  // lx == TS - 1
}
__device__ __forceinline__ void perim_cell(int k, int &ly, int &lx) {
    if (k < TS) { ly = 0; lx = k; }
    else if (k < 2 * TS) { ly = TS - 1; lx = k - TS; }
    else if (k < 2 * TS + TS - 2) { ly = k - 2 * TS + 1; lx = 0; }
    else { ly = k - (2 * TS + TS - 2) + 1; lx = TS - 1; }
}
__device__ __forceinline__ int node_of(int gr, int gc, int ntx) {
    return ((gr >> 6) * ntx + (gc >> 6)) * NSLOT + perim_index(gr & (TS - 1), gc & (TS - 1));
}
__device__ __forceinline__ void code_move(int d, int &dc, int &dr) {  // keypad 7 8 9 / 4 5 6 / 1 2 3, rows grow downwards
    // dc = (d - 1) % 3 - 1, dr = 1 - (d - 1) / 3 for d = 1..9, from two 2-bit tables (value + 1 at bits 2d, 2d + 1)
    constexpr unsigned DC = (0u << 2) | (1u << 4) | (2u << 6) | (0u << 8) | (1u << 10) | (2u << 12) | (0u << 14) | (1u << 16) | (2u << 18);
    constexpr unsigned DR = (2u << 2) | (2u << 4) | (2u << 6) | (1u << 8) | (1u << 10) | (1u << 12) | (0u << 14) | (0u << 16) | (0u << 18);
    dc = (int)((DC >> (2 * d)) & 3u) - 1;
    dr = (int)((DR >> (2 * d)) & 3u) - 1;
}
// effective direction code of raster cell (gr, gc): 0 = no downstream (outside the raster, not a D8 code, points off the
// map, or the outlet -- the search never follows the outlet's own direction, cuencas.f90:552-589)
__device__ __forceinline__ int eff_code(const int32_t *__restrict__ dir, int nc, int nr, int gr, int gc, int lin0) {
    if (gr < 0 || gr >= nr || gc < 0 || gc >= nc) return 0;
    const int lin = gr * nc + gc;
    if (lin == lin0) return 0;
    const int d = __ldg(dir + lin);
    if (d < 1 || d > 9 || d == 5) return 0;
    int dc, dr;
    code_move(d, dc, dr);
    const int tc = gc + dc, tr = gr + dr;
    if (tc < 0 || tc >= nc || tr < 0 || tr >= nr) return 0;
    return d;
}

// ---------------------------------------------------------------------------------------------------------------------
// stage A: per tile, where the perimeter cells' paths leave the tile and how many cells every exit cell collects
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHR, 2)
k_pf_local1(const int32_t *__restrict__ dir, int nc, int nr, int ntx, int lin0, int32_t *__restrict__ pterm,
            int32_t *__restrict__ bdown, uint32_t *__restrict__ lsize) {
    __shared__ unsigned char code[TC];
    __shared__ unsigned short T[TC];
    __shared__ unsigned cnt[NSLOT];
    const int tid = threadIdx.x, tile = blockIdx.x;
    const int r0 = (tile / ntx) * TS, c0 = (tile % ntx) * TS;
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
        const int li = k * NTHR + tid, ly = li >> 6, lx = li & (TS - 1);
        const int d = eff_code(dir, nc, nr, r0 + ly, c0 + lx, lin0);
        int t = li;
        if (d) {
            int dc, dr;
            code_move(d, dc, dr);
            const int ny = ly + dr, nx = lx + dc;
            if (ny >= 0 && ny < TS && nx >= 0 && nx < TS) t = ny * TS + nx;
        }
        code[li] = (unsigned char)d;
        T[li] = (unsigned short)t;
    }
    if (tid < NSLOT) cnt[tid] = 0;
    __syncthreads();
    // pointer jumping to the local terminal (T == self).  Unsynchronised in-place updates are safe here: T[x] always is
    // x itself or a cell further down x's path, whatever mix of old and new values a thread reads.
    for (int round = 0; round < MAXROUNDS; ++round) {
        int changed = 0;
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
            const int li = k * NTHR + tid;
            const int t = T[li], tt = T[t];
            if (tt != t) { T[li] = (unsigned short)tt; changed = 1; }
        }
        if (!__syncthreads_or(changed)) break;
    }
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
        const int li = k * NTHR + tid;
        const int t = T[li];
        if (T[t] == t && code[t] != 0) atomicAdd(&cnt[perim_index(t >> 6, t & (TS - 1))], 1u);  // an exit cell collects me
    }
    __syncthreads();
    if (tid < NSLOT) {
        const int node = tile * NSLOT + tid;
        int pt = DEAD, bd = -1;
        unsigned ls = 0;
        if (tid < NPERIM) {
            int ly, lx;
            perim_cell(tid, ly, lx);
            const int li = ly * TS + lx;
            const int t = T[li];
            if (T[t] == t) {  // (cells of a cycle inside the tile never reach a terminal: dead)
                if (code[t] != 0) pt = perim_index(t >> 6, t & (TS - 1));
                else {
                    const int gr = r0 + (t >> 6), gc = c0 + (t & (TS - 1));
                    pt = (gr < nr && gc < nc && gr * nc + gc == lin0) ? ROOT_OUT : DEAD;
                }
            }
            if (t == li && code[li] != 0) {  // an exit cell: its downstream neighbour lies in another tile
                int dc, dr;
                code_move(code[li], dc, dr);
                bd = node_of(r0 + ly + dr, c0 + lx + dc, ntx);
                ls = cnt[tid];
            }
        }
        pterm[node] = pt;
        bdown[node] = bd;
        lsize[node] = ls;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// stage B: the boundary graph (exit cell -> the exit cell its water reaches next) and the global size of the exit cells
// ---------------------------------------------------------------------------------------------------------------------
// w[x] = (arrivals - tributaries) << 32 | cells collected so far   (as k_acum_climb in topo.cu)
__global__ void k_pf_bg_init(int n_nodes, const int32_t *__restrict__ pterm, const int32_t *__restrict__ bdown,
                             const uint32_t *__restrict__ lsize, int32_t *__restrict__ bnext,
                             unsigned long long *__restrict__ w) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n_nodes) return;
    const int bd = bdown[x];
    int nx = DEAD;
    if (bd >= 0) {
        const int pt = pterm[bd];
        nx = pt >= 0 ? (bd & ~(NSLOT - 1)) + pt : pt;
    }
    bnext[x] = nx;
    w[x] = (unsigned long long)lsize[x];
}
__global__ void k_pf_bg_count(int n_nodes, const int32_t *__restrict__ bnext, unsigned long long *__restrict__ w) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n_nodes) return;
    const int nx = bnext[x];
    if (nx >= 0) atomicAdd(&w[nx], 0xFFFFFFFF00000000ull);
}
__global__ void k_pf_bg_leaves(int n_nodes, const int32_t *__restrict__ bdown, const unsigned long long *__restrict__ w,
                               unsigned char *__restrict__ leaf) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n_nodes) return;
    leaf[x] = (bdown[x] >= 0 && (unsigned)(w[x] >> 32) == 0u) ? 1 : 0;
}
__global__ void k_pf_bg_climb(int n_nodes, const int32_t *__restrict__ bnext, const unsigned char *__restrict__ leaf,
                              unsigned long long *__restrict__ w) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n_nodes || !leaf[x]) return;
    unsigned v = (unsigned)w[x];  // a leaf's total is final: nobody adds to it
    int cur = x;
    for (;;) {
        const int p = bnext[cur];
        if (p < 0) break;
        const unsigned long long old = atomicAdd(&w[p], (1ull << 32) + (unsigned long long)v);
        if ((unsigned)(old >> 32) != 0xFFFFFFFFu) break;  // not the last tributary to arrive
        v = (unsigned)old + v;
        cur = p;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// stage C: per tile, global sizes, sibling offsets, hops and preorder sums to the local terminal
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int halo_slot(int ey, int ex) {  // (ey, ex) in [-1, TS], on the halo ring
    if (ey == -1) return ex + 1;
    if (ey == TS) return EW + ex + 1;
    if (ex == -1) return 2 * EW + ey;
    return 2 * EW + TS + ey;  // ex == TS
}
constexpr int NHALO = 2 * EW + 2 * TS;

struct Local2Smem {
    uint32_t bufA[TC], bufB[TC];
    unsigned short j0[TC], j1[TC];
    uint32_t hG[NHALO];
    unsigned char code[EW * EW];
};

__global__ void __launch_bounds__(NTHR, 4)
k_pf_local2(const int32_t *__restrict__ dir, int nc, int nr, int ntx, int lin0, const unsigned long long *__restrict__ w,
            uint32_t *__restrict__ S, uint2 *__restrict__ LR, uint32_t *__restrict__ boff, uint2 *__restrict__ eho) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Local2Smem &sm = *reinterpret_cast<Local2Smem *>(smem_raw);
    const int tid = threadIdx.x, tile = blockIdx.x;
    const int r0 = (tile / ntx) * TS, c0 = (tile % ntx) * TS;
    // the 3x3 probe order of cuencas.f90:555-571 and the code a neighbour must carry to be a child
    constexpr int PDC[8] = {-1, 0, 1, -1, 1, -1, 0, 1}, PDR[8] = {-1, -1, -1, 0, 0, 1, 1, 1}, PCODE[8] = {3, 2, 1, 6, 4, 9, 8, 7};
    {   // (row, column) of the extended tile advance by NTHR cells per trip without a division
        constexpr int DY = NTHR / EW, DX = NTHR % EW;
        int ey = tid / EW, ex = tid % EW;
        for (int idx = tid; idx < EW * EW; idx += NTHR) {
            sm.code[idx] = (unsigned char)eff_code(dir, nc, nr, r0 + ey - 1, c0 + ex - 1, lin0);
            ex += DX;
            ey += DY;
            if (ex >= EW) { ex -= EW; ++ey; }
        }
    }
    __syncthreads();
    uint32_t *Ac = sm.bufA, *An = sm.bufB;
    unsigned short *Jc = sm.j0, *Jn = sm.j1;
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
        const int li = k * NTHR + tid, ly = li >> 6, lx = li & (TS - 1);
        const int d = sm.code[(ly + 1) * EW + lx + 1];
        int j = NONE16;
        if (d) {
            int dc, dr;
            code_move(d, dc, dr);
            const int ny = ly + dr, nx = lx + dc;
            if (ny >= 0 && ny < TS && nx >= 0 && nx < TS) j = ny * TS + nx;
        }
        Jc[li] = (unsigned short)j;
        Ac[li] = 1u;
    }
    __syncthreads();
    // what enters from the neighbouring tiles: a halo cell pointing at one of my cells is an exit cell of its own tile
    for (int h = tid; h < NHALO; h += NTHR) {
        int ey, ex;
        if (h < EW) { ey = -1; ex = h - 1; }
        else if (h < 2 * EW) { ey = TS; ex = h - EW - 1; }
        else if (h < 2 * EW + TS) { ey = h - 2 * EW; ex = -1; }
        else { ey = h - 2 * EW - TS; ex = TS; }
        const int d = sm.code[(ey + 1) * EW + ex + 1];
        unsigned g = 0;
        if (d) {
            int dc, dr;
            code_move(d, dc, dr);
            const int ty = ey + dr, tx = ex + dc;
            if (ty >= 0 && ty < TS && tx >= 0 && tx < TS) {
                g = (unsigned)w[node_of(r0 + ey, c0 + ex, ntx)];
                atomicAdd(&Ac[ty * TS + tx], g);
            }
        }
        sm.hG[h] = g;
    }
    __syncthreads();
    // leaf-to-root sums by doubling: J_k(u) = the cell 2^k steps down u's path inside the tile (or none),
    // A_k(v) = sum over the cells less than 2^k steps above v;  A_{k+1}(v) = A_k(v) + sum_{J_k(u) = v} A_k(u)
    for (int round = 0; round < MAXROUNDS; ++round) {
#pragma unroll
        for (int k = 0; k < CPT; ++k) An[k * NTHR + tid] = Ac[k * NTHR + tid];
        __syncthreads();
        int live = 0;
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
            const int li = k * NTHR + tid;
            const int j = Jc[li];
            int jn = NONE16;
            if (j != NONE16) {
                atomicAdd(&An[j], Ac[li]);
                jn = Jc[j];
                live |= (jn != NONE16);
            }
            Jn[li] = (unsigned short)jn;
        }
        { uint32_t *t = Ac; Ac = An; An = t; }
        { unsigned short *t = Jc; Jc = Jn; Jn = t; }
        if (!__syncthreads_or(live)) break;
    }
    // Ac = global size of every cell of the tile (== basin_acum for the cells of the basin)
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
        const int li = k * NTHR + tid, gr = r0 + (li >> 6), gc = c0 + (li & (TS - 1));
        if (gr < nr && gc < nc) S[(size_t)gr * nc + gc] = Ac[li];
        An[li] = 0u;  // becomes off(), then the preorder sum to the terminal
    }
    __syncthreads();
    // sibling offsets: the cell that is drained into numbers its tributaries in probe order
    unsigned nchpack = 0;
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
        const int li = k * NTHR + tid, ly = li >> 6, lx = li & (TS - 1);
        unsigned run = 1u, nch = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int ey = ly + PDR[q], ex = lx + PDC[q];
            if (sm.code[(ey + 1) * EW + ex + 1] == PCODE[q]) {
                if (ey >= 0 && ey < TS && ex >= 0 && ex < TS) {
                    An[ey * TS + ex] = run;
                    run += Ac[ey * TS + ex];
                } else {
                    boff[node_of(r0 + ey, c0 + ex, ntx)] = run;
                    run += sm.hG[halo_slot(ey, ex)];
                }
                ++nch;
            }
        }
        nchpack |= nch << (4 * k);
    }
    __syncthreads();
    // root-to-leaf sums by pointer jumping on one 64-bit word per cell: (T = cell reached so far | H = hops to it << 16 |
    // O = sum of off() over the cells passed << 32; the terminal's own offset belongs to the boundary graph).  A word is
    // read and written whole, so it always describes a valid jump and the updates need no read / write phases: a cell adds
    // the word of the cell it points at, whichever version of it it sees, until that cell is a terminal (T == itself).
    int tv[CPT];
    unsigned offv[CPT];
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
        const int li = k * NTHR + tid, ly = li >> 6, lx = li & (TS - 1);
        const int d = sm.code[(ly + 1) * EW + lx + 1];
        int t = li;
        if (d) {
            int dc, dr;
            code_move(d, dc, dr);
            const int ny = ly + dr, nx = lx + dc;
            if (ny >= 0 && ny < TS && nx >= 0 && nx < TS) t = ny * TS + nx;
        }
        tv[k] = t;
        offv[k] = (t != li) ? An[li] : 0u;
    }
    __syncthreads();
    unsigned long long *P64 = reinterpret_cast<unsigned long long *>(sm.bufA);   // bufA + bufB: 4096 words
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
        const int li = k * NTHR + tid;
        P64[li] = (unsigned long long)tv[k] | ((unsigned long long)(tv[k] != li ? 1u : 0u) << 16) | ((unsigned long long)offv[k] << 32);
    }
    __syncthreads();
    unsigned done = 0;
    for (int round = 0; round <= MAXROUNDS; ++round) {
        int wrote = 0;
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
            if ((done >> k) & 1u) continue;
            const int li = k * NTHR + tid;
            const unsigned long long p = P64[li];
            const int t = (int)(p & 0xFFFFu);
            const unsigned long long q = P64[t];
            if ((int)(q & 0xFFFFu) == t) { done |= 1u << k; continue; }   // t is a terminal (my own word when I am one)
            // (hops stay below 2^16 -- a path inside a tile has at most 4095 -- so the low words add without a carry into T)
            const unsigned plo = (unsigned)p, qlo = (unsigned)q;
            const unsigned lo = (qlo & 0xFFFFu) | ((plo & 0xFFFF0000u) + (qlo & 0xFFFF0000u));
            const unsigned hi = (unsigned)(p >> 32) + (unsigned)(q >> 32);
            P64[li] = (unsigned long long)lo | ((unsigned long long)hi << 32);
            wrote = 1;
        }
        if (!__syncthreads_or(wrote)) break;
    }
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
        const int li = k * NTHR + tid, ly = li >> 6, lx = li & (TS - 1);
        const int gr = r0 + ly, gc = c0 + lx;
        const unsigned long long p = P64[li];
        const int t = (int)(p & 0xFFFFu);
        unsigned tc = T_DEAD;
        if ((int)(P64[t] & 0xFFFFu) == t) {
            if (sm.code[((t >> 6) + 1) * EW + (t & (TS - 1)) + 1] != 0) tc = (unsigned)perim_index(t >> 6, t & (TS - 1));
            else {
                const int tr = r0 + (t >> 6), tcc = c0 + (t & (TS - 1));
                tc = (tr < nr && tcc < nc && tr * nc + tcc == lin0) ? T_ROOT_OUT : T_DEAD;
            }
        }
        const unsigned h = (unsigned)(p >> 16) & 0xFFFu, o = (unsigned)(p >> 32), nch = (nchpack >> (4 * k)) & 15u;   // (h <= 4095 unless dead)
        if (gr < nr && gc < nc) LR[(size_t)gr * nc + gc] = make_uint2(tc | (h << 16) | (nch << 28), o);
        if (ly == 0 || ly == TS - 1 || lx == 0 || lx == TS - 1) eho[tile * NSLOT + perim_index(ly, lx)] = make_uint2(h, o);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// stage D: depth and preorder of the exit cells
// ---------------------------------------------------------------------------------------------------------------------
// Only exit cells (a third of the perimeter slots) take part from here on: they are renumbered densely, in slot order
// (cmap = exclusive scan of "is an exit cell"), and the jump rounds run over the dense arrays.
__global__ void k_pf_exit_flag(int n_nodes, const int32_t *__restrict__ bdown, int32_t *__restrict__ flag) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x < n_nodes) flag[x] = bdown[x] >= 0 ? 1 : 0;
}
__global__ void k_pf_bd_init(int n_nodes, const int32_t *__restrict__ bdown, const int32_t *__restrict__ bnext,
                             const uint32_t *__restrict__ boff, const uint2 *__restrict__ eho, const int32_t *__restrict__ cmap,
                             int32_t *__restrict__ nx, uint2 *__restrict__ dp, unsigned char *__restrict__ age) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n_nodes) return;
    const int bd = bdown[x];
    if (bd < 0) return;
    const int c = cmap[x];
    const int n = bnext[x];
    const uint2 e = eho[bd];
    nx[c] = n >= 0 ? cmap[n] : n;
    dp[c] = make_uint2(1u + e.x, boff[x] + e.y);  // one hop into the next tile + that cell's way to its terminal
    age[c] = 0;
}
// A node whose jump has ended is copied into the other buffer once more (age 1 -> 2) and then left alone.
__global__ void k_pf_bd_jump(int n_nodes, const int32_t *__restrict__ nx_in, const uint2 *__restrict__ dp_in,
                             int32_t *__restrict__ nx_out, uint2 *__restrict__ dp_out, unsigned char *__restrict__ age,
                             int *__restrict__ more) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n_nodes) return;
    const unsigned char a = age[x];
    if (a >= 2) return;
    int n = nx_in[x];
    uint2 v = dp_in[x];
    if (n >= 0) {
        const int n2 = nx_in[n];
        const uint2 u = dp_in[n];
        v.x += u.x;
        v.y += u.y;
        n = n2;
        if (n2 >= 0) *more = 1;
    } else {
        age[x] = a + 1;
    }
    nx_out[x] = n;
    dp_out[x] = v;
}

// Small boundary graphs (C1: 64 tiles): all the jump rounds in ONE CTA, __syncthreads() between rounds instead of a launch, a
// flag read-back and a host synchronisation per round.  The result is left in (nx_a, dp_a).
constexpr int BDS_THREADS = 1024, BDS_MAX_NODES = 1 << 16;
__global__ void __launch_bounds__(BDS_THREADS)
k_pf_bd_jump_small(int n_slots, const int32_t *__restrict__ cmap, const int32_t *__restrict__ flag, int32_t *__restrict__ nx_a,
                   uint2 *__restrict__ dp_a, int32_t *__restrict__ nx_b, uint2 *__restrict__ dp_b) {
    const int n_nodes = cmap[n_slots - 1] + flag[n_slots - 1];   // exit cells
    int32_t *nin = nx_a, *nout = nx_b;
    uint2 *din = dp_a, *dout = dp_b;
    for (int round = 0; round < 40; ++round) {
        int more = 0;
        for (int x = threadIdx.x; x < n_nodes; x += BDS_THREADS) {
            int n = nin[x];
            uint2 v = din[x];
            if (n >= 0) {
                const int n2 = nin[n];
                const uint2 u = din[n];
                v.x += u.x;
                v.y += u.y;
                n = n2;
                if (n2 >= 0) more = 1;
            }
            nout[x] = n;
            dout[x] = v;
        }
        { int32_t *t = nin; nin = nout; nout = t; }
        { uint2 *t = din; din = dout; dout = t; }
        if (!__syncthreads_or(more)) break;  // (also orders this round's global writes before the next round's reads)
    }
    if (nin != nx_a) {   // an odd number of rounds: the result sits in the second buffer
        for (int x = threadIdx.x; x < n_nodes; x += BDS_THREADS) {
            nx_a[x] = nin[x];
            dp_a[x] = din[x];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// stage E: every cell of the basin writes (depth | children << 28, raster index) at position preorder
// ---------------------------------------------------------------------------------------------------------------------
// One CTA per tile: the DFS preorder is spatially local (the next cell in preorder is the first tributary, or the next
// sibling of a near ancestor), so the cells of a tile own long runs of consecutive preorder positions; written by one
// CTA within microseconds they merge into full sectors in L2.
__global__ void __launch_bounds__(NTHR, 3)
k_pf_emit(const uint2 *__restrict__ LR, int nc, int nr, int ntx, const int32_t *__restrict__ cmap, int n_exit,
          const int32_t *__restrict__ nx, const uint2 *__restrict__ dp,
          uint32_t n_basin, uint2 *__restrict__ items /* (depth | children << 28, raster index) at position preorder */,
          unsigned *__restrict__ stat /* [0] max depth, [1] errors */) {
    // depth / preorder of the tile's (at most 252) exit cells, once per tile instead of three dependent gathers per cell
    __shared__ uint2 s_dp[NSLOT];
    __shared__ int s_nx[NSLOT];
    const int tile = blockIdx.x, tid = threadIdx.x;
    const int r0 = (tile / ntx) * TS, c0 = (tile % ntx) * TS;
    unsigned m = 0;
    uint2 lr[CPT];
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
        const int li = k * NTHR + tid, gr = r0 + (li >> 6), gc = c0 + (li & (TS - 1));
        lr[k] = (gc < nc && gr < nr) ? __ldg(LR + (size_t)gr * nc + gc) : make_uint2(T_DEAD, 0u);
    }
    if (tid < NSLOT) {
        const int node = __ldg(cmap + tile * NSLOT + tid);   // (slots that are no exit cells are never looked up)
        const bool ok = node < n_exit;
        s_nx[tid] = ok ? __ldg(nx + node) : DEAD;
        s_dp[tid] = ok ? __ldg(dp + node) : make_uint2(0u, 0u);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
        const int li = k * NTHR + tid, gr = r0 + (li >> 6), gc = c0 + (li & (TS - 1));
        const unsigned tc = lr[k].x & 0xFFFFu, h = (lr[k].x >> 16) & 0xFFFu, nch = lr[k].x >> 28;
        unsigned depth = 0, pre = 0;
        bool ok = false;
        if (tc == T_ROOT_OUT) { depth = h; pre = lr[k].y; ok = true; }
        else if (tc != T_DEAD && s_nx[tc] == ROOT_OUT) {
            const uint2 v = s_dp[tc];
            depth = v.x + h;
            pre = v.y + lr[k].y;
            ok = true;
        }
        if (ok) {
            if (pre < n_basin && depth < (1u << 28)) {
                items[pre] = make_uint2(depth | (nch << 28), (uint32_t)((size_t)gr * nc + gc));   // one 8-byte store
                m = max(m, depth);
            } else {
                atomicAdd(&stat[1], 1u);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(&stat[0], m);
}

// The reference's search treats the outlet like any other cell: if the cell the outlet itself drains to belongs to the
// basin, the outlet is found again as a tributary and the queue never ends (cuencas.f90:552-589).  stat[2] = 1 then.
__global__ void k_pf_outlet_check(const int32_t *__restrict__ dir, const uint2 *__restrict__ LR, int nc, int nr, int ntx, int lin0,
                                  const int32_t *__restrict__ cmap, const int32_t *__restrict__ nx, unsigned *__restrict__ stat) {
    const int d = dir[lin0];
    if (d < 1 || d > 9 || d == 5) return;
    int dc, dr;
    code_move(d, dc, dr);
    const int gr = lin0 / nc + dr, gc = lin0 % nc + dc;
    if (gr < 0 || gr >= nr || gc < 0 || gc >= nc) return;
    const uint2 lr = LR[(size_t)gr * nc + gc];
    const unsigned tc = lr.x & 0xFFFFu;
    bool in_basin = false;
    if (tc == T_ROOT_OUT) in_basin = true;
    else if (tc != T_DEAD) in_basin = nx[cmap[((gr >> 6) * ntx + (gc >> 6)) * NSLOT + (int)tc]] == ROOT_OUT;
    if (in_basin) stat[2] = 1u;
}

// ---------------------------------------------------------------------------------------------------------------------
// stable LSD radix pass, 8 bits: per-tile digit histograms, one device scan, ranked scatter through shared memory
// ---------------------------------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 512, RS_WARPS = RS_THREADS / 32, RS_TILE = 4096, RS_PER_WARP = RS_TILE / RS_WARPS, RS_SUB = RS_PER_WARP / 32;

// (IL: the first pass reads the interleaved (key, value) pairs k_pf_emit wrote)
template <bool IL>
__global__ void __launch_bounds__(RS_THREADS)
k_rs_hist(const uint32_t *__restrict__ keys, uint32_t n, int shift, int nblk, int32_t *__restrict__ hist /* [256][nblk] */) {
    __shared__ unsigned h[256];
    if (threadIdx.x < 256) h[threadIdx.x] = 0;
    __syncthreads();
    const size_t base = (size_t)blockIdx.x * RS_TILE;
    if (IL) {
        const uint2 *items = reinterpret_cast<const uint2 *>(keys);
        constexpr int PT = RS_TILE / RS_THREADS;
        uint32_t kk[PT];
#pragma unroll
        for (int i = 0; i < PT; ++i) {
            const size_t idx = base + (size_t)i * RS_THREADS + threadIdx.x;
            kk[i] = idx < n ? __ldg(items + idx).x : 0u;
        }
#pragma unroll
        for (int i = 0; i < PT; ++i)
            if (base + (size_t)i * RS_THREADS + threadIdx.x < n) atomicAdd(&h[(kk[i] >> shift) & 255u], 1u);
        __syncthreads();
        if (threadIdx.x < 256) hist[(size_t)threadIdx.x * nblk + blockIdx.x] = (int32_t)h[threadIdx.x];
        return;
    }
    constexpr int V = RS_TILE / (4 * RS_THREADS);   // 128-bit loads per thread (the tile start is 16-byte aligned)
    uint4 q[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const size_t idx = base + 4 * ((size_t)i * RS_THREADS + threadIdx.x);
        if (idx + 3 < n) q[i] = __ldg(reinterpret_cast<const uint4 *>(keys + idx));
        else {
            q[i].x = idx < n ? keys[idx] : 0xFFFFFFFFu;
            q[i].y = idx + 1 < n ? keys[idx + 1] : 0xFFFFFFFFu;
            q[i].z = idx + 2 < n ? keys[idx + 2] : 0xFFFFFFFFu;
            q[i].w = 0xFFFFFFFFu;
        }
    }
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const size_t idx = base + 4 * ((size_t)i * RS_THREADS + threadIdx.x);
        const uint32_t kk[4] = {q[i].x, q[i].y, q[i].z, q[i].w};
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (idx + c < n) atomicAdd(&h[(kk[c] >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 256) hist[(size_t)threadIdx.x * nblk + blockIdx.x] = (int32_t)h[threadIdx.x];
}

// Every warp owns 512 consecutive items of the tile, so "warp, sub-step, lane" is the input order: the ranks are stable.
// The tile is first sorted by digit inside shared memory; the copy-out then writes each digit's run with consecutive
// threads on consecutive addresses (runs of ~16 items instead of 4-byte writes all over the output).
struct RsSmem {
    uint32_t k[RS_TILE], v[RS_TILE];
    unsigned wh[RS_WARPS][256];
    unsigned gbase[256];
    int scan[33];
};

template <bool IL>
__global__ void __launch_bounds__(RS_THREADS, 3)
k_rs_scatter(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, uint32_t n, int shift, int nblk,
             const int32_t *__restrict__ hist_scan, uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RsSmem &sm = *reinterpret_cast<RsSmem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int i = tid; i < RS_WARPS * 256; i += RS_THREADS) (&sm.wh[0][0])[i] = 0;
    __syncthreads();
    const size_t tile0 = (size_t)blockIdx.x * RS_TILE, base = tile0 + (size_t)wid * RS_PER_WARP;
    uint32_t k[RS_SUB], v[RS_SUB];
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int s = 0; s < RS_SUB; ++s) {
        const size_t idx = base + s * 32 + lane;
        const bool valid = idx < n;
        if (IL) {
            const uint2 it = valid ? __ldg(reinterpret_cast<const uint2 *>(keys) + idx) : make_uint2(0u, 0u);
            k[s] = it.x;
            v[s] = it.y;
        } else {
            k[s] = valid ? __ldg(keys + idx) : 0u;
            v[s] = valid ? __ldg(vals + idx) : 0u;
        }
    }
    // one match per item: its rank among the warp's earlier items of the same digit (running per-warp counters), kept in a
    // register until the digit bases are known
    unsigned rk[RS_SUB];
#pragma unroll
    for (int s = 0; s < RS_SUB; ++s) {
        const bool valid = base + s * 32 + lane < n;
        const int d = valid ? (int)((k[s] >> shift) & 255u) : 256 + lane;   // invalid lanes match nobody
        const unsigned m = __match_any_sync(0xffffffffu, d);
        rk[s] = valid ? sm.wh[wid][d] + __popc(m & lt) : 0u;
        __syncwarp();
        if (valid && (m & lt) == 0) sm.wh[wid][d] += __popc(m);   // the lowest lane of every digit group
        __syncwarp();
    }
    __syncthreads();
    // digit `tid`: where each warp's items of that digit start inside the sorted tile, and where the run goes globally
    unsigned tot = 0;
    if (tid < 256) {
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) tot += sm.wh[w][tid];
    }
    int cta_tot;
    const int start = block_scan_excl((int)tot, sm.scan, &cta_tot);
    if (tid < 256) {
        unsigned run = (unsigned)start;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            const unsigned c = sm.wh[w][tid];
            sm.wh[w][tid] = run;
            run += c;
        }
        sm.gbase[tid] = (unsigned)hist_scan[(size_t)tid * nblk + blockIdx.x] - (unsigned)start;
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < RS_SUB; ++s) {
        if (base + s * 32 + lane < n) {
            const unsigned pos = sm.wh[wid][(k[s] >> shift) & 255u] + rk[s];
            sm.k[pos] = k[s];
            sm.v[pos] = v[s];
        }
    }
    __syncthreads();
    for (int i = tid; i < cta_tot; i += RS_THREADS) {
        const uint32_t kk = sm.k[i];
        const unsigned g = sm.gbase[(kk >> shift) & 255u] + (unsigned)i;
        keys_out[g] = kk;
        vals_out[g] = sm.v[i];
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// stage F: queue order -> q_par (position of the parent, 1-based), q_acum
// ---------------------------------------------------------------------------------------------------------------------
constexpr int FN_THREADS = 256, FN_WARPS = FN_THREADS / 32, FN_TILE = 2048, FN_PER_WARP = FN_TILE / FN_WARPS, FN_SUB = FN_PER_WARP / 32;

__global__ void __launch_bounds__(FN_THREADS)
k_pf_final_sums(const uint32_t *__restrict__ keys, uint32_t n, int32_t *__restrict__ blk_sum) {
    __shared__ int sm[33];
    const size_t base = (size_t)blockIdx.x * FN_TILE;
    int s = 0;
#pragma unroll 4
    for (int i = threadIdx.x; i < FN_TILE; i += FN_THREADS)
        if (base + i < n) s += (int)(__ldg(keys + base + i) >> 28);
    int tot;
    block_scan_excl(s, sm, &tot);
    if (threadIdx.x == 0) blk_sum[blockIdx.x] = tot;
}

// the children of queue position j occupy positions [cs(j), cs(j) + nch(j)), cs(j) = 1 + children of the positions before j.
// A warp takes 512 consecutive positions, 32 at a time: loads, the size gather and the parent writes are coalesced
// (consecutive lanes own consecutive child ranges).
__global__ void __launch_bounds__(FN_THREADS, 6)
k_pf_final_write(const uint32_t *__restrict__ keys, const int32_t *__restrict__ q_lin, uint32_t n,
                 const int32_t *__restrict__ blk_scan, const uint32_t *__restrict__ S, int32_t *__restrict__ q_par,
                 int32_t *__restrict__ q_acum, unsigned *__restrict__ stat) {
    __shared__ int wtot[FN_WARPS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const size_t base = (size_t)blockIdx.x * FN_TILE + (size_t)wid * FN_PER_WARP;
    int c[FN_SUB], ex[FN_SUB];
    int run = 0;
#pragma unroll
    for (int s = 0; s < FN_SUB; ++s) {
        const size_t j = base + s * 32 + lane;
        c[s] = (j < n) ? (int)(__ldg(keys + j) >> 28) : 0;
        const int incl = warp_scan_incl(c[s]);
        ex[s] = run + incl - c[s];
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) wtot[wid] = run;
    __syncthreads();
    long long cs0 = 1ll + blk_scan[blockIdx.x];
    for (int w = 0; w < wid; ++w) cs0 += wtot[w];
    int32_t lin[FN_SUB];
#pragma unroll
    for (int s = 0; s < FN_SUB; ++s) {
        const size_t j = base + s * 32 + lane;
        lin[s] = (j < n) ? __ldg(q_lin + j) : -1;
    }
    uint32_t sz[FN_SUB];
#pragma unroll
    for (int s = 0; s < FN_SUB; ++s) sz[s] = (lin[s] >= 0) ? __ldg(S + lin[s]) : 0u;
#pragma unroll
    for (int s = 0; s < FN_SUB; ++s) {
        const size_t j = base + s * 32 + lane;
        if (j < n) {
            const long long cs = cs0 + ex[s];
            for (int q = 0; q < c[s]; ++q) {
                if (cs + q < (long long)n) q_par[cs + q] = (int32_t)(j + 1);
                else atomicAdd(&stat[1], 1u);
            }
            q_acum[j] = (int32_t)sz[s];
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) q_par[0] = 0;
}

// basin_acum on the table basin_find itself produced: the sizes are already known.  ok[0] = 0 when a drain id differs.
__global__ void k_pf_acum_cached(const int32_t *__restrict__ basin_f, const int32_t *__restrict__ q_par,
                                 const int32_t *__restrict__ q_acum, int n, int32_t *__restrict__ acum, int *__restrict__ bad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int p = n - 1 - i;
    if (basin_f[3 * (size_t)i] != q_par[p]) *bad = 1;
    acum[i] = q_acum[p];
}

// debug dumps only: the dense (nx, dp) back on the perimeter slots
__global__ void k_pf_expand(int n_nodes, const int32_t *__restrict__ bdown, const int32_t *__restrict__ cmap, const int32_t *__restrict__ nx,
                            const uint2 *__restrict__ dp, int32_t *__restrict__ nx_full, uint2 *__restrict__ dp_full) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n_nodes) return;
    const bool e = bdown[x] >= 0;
    nx_full[x] = e ? nx[cmap[x]] : DEAD;
    dp_full[x] = e ? dp[cmap[x]] : make_uint2(0u, 0u);
}

static void dump(wmf_ctx *ctx, const char *dir, const char *name, const void *dev, size_t bytes) {
    if (!dir) return;
    std::vector<char> h(bytes);
    cudaStreamSynchronize(ctx->stream);
    cudaMemcpy(h.data(), dev, bytes, cudaMemcpyDeviceToHost);
    const std::string path = std::string(dir) + "/pf_" + name + ".bin";
    if (FILE *f = fopen(path.c_str(), "wb")) { fwrite(h.data(), 1, bytes, f); fclose(f); }
}

}  // namespace pf

// Host side.  Leaves the queue in ctx->q_lin / q_par (what basin_cut reads) and the sizes in ctx->q_acum.
int wmf_pathfind(wmf_ctx *ctx, Scope &sc, const int32_t *d_dir, int nc, int nr, int lin0, int32_t *n_out, int32_t *depth_out) {
    using namespace pf;
    const char *dumpdir = getenv("WMF_PF_DUMP");
    const int ntx = (nc + TS - 1) / TS, nty = (nr + TS - 1) / TS;
    const long long ntiles_ll = (long long)ntx * nty;
    if (ntiles_ll * NSLOT >= INT_MAX) WMF_FAIL(ctx, WMF_ERR_ARG, "basin_find: raster too large for the tile graph");
    const int ntiles = (int)ntiles_ll, n_nodes = ntiles * NSLOT;
    const size_t cells = (size_t)nc * nr;
    int32_t *pterm, *bdown, *bnext, *nxA, *nxB;
    uint32_t *lsize, *boff, *S;
    unsigned long long *w;
    unsigned char *leaf;
    uint2 *eho, *dpA, *dpB, *LR;
    unsigned *stat;
    int *more;
    WMF_TRY(sc.alloc(&pterm, (size_t)n_nodes)); WMF_TRY(sc.alloc(&bdown, (size_t)n_nodes)); WMF_TRY(sc.alloc(&bnext, (size_t)n_nodes));
    WMF_TRY(sc.alloc(&lsize, (size_t)n_nodes)); WMF_TRY(sc.alloc(&boff, (size_t)n_nodes)); WMF_TRY(sc.alloc(&w, (size_t)n_nodes));
    WMF_TRY(sc.alloc(&leaf, (size_t)n_nodes)); WMF_TRY(sc.alloc(&eho, (size_t)n_nodes));
    WMF_TRY(sc.alloc(&nxA, (size_t)n_nodes)); WMF_TRY(sc.alloc(&nxB, (size_t)n_nodes));
    WMF_TRY(sc.alloc(&dpA, (size_t)n_nodes)); WMF_TRY(sc.alloc(&dpB, (size_t)n_nodes));
    WMF_TRY(sc.alloc(&S, cells)); WMF_TRY(sc.alloc(&LR, cells));
    WMF_TRY(sc.alloc(&stat, 4)); WMF_TRY(sc.alloc(&more, 1));
    CU_TRY(ctx, cudaMemsetAsync(stat, 0, 4 * sizeof(unsigned), ctx->stream));
    CU_TRY(ctx, cudaMemsetAsync(boff, 0, sizeof(uint32_t) * (size_t)n_nodes, ctx->stream));
    const int B = 256, GN = grid_for(n_nodes, B);
    // A
    LAUNCH(ctx, k_pf_local1, ntiles, NTHR, 0, d_dir, nc, nr, ntx, lin0, pterm, bdown, lsize);
    // B
    LAUNCH(ctx, k_pf_bg_init, GN, B, 0, n_nodes, pterm, bdown, lsize, bnext, w);
    LAUNCH(ctx, k_pf_bg_count, GN, B, 0, n_nodes, bnext, w);
    LAUNCH(ctx, k_pf_bg_leaves, GN, B, 0, n_nodes, bdown, w, leaf);
    LAUNCH(ctx, k_pf_bg_climb, GN, B, 0, n_nodes, bnext, leaf, w);
    // C
    CU_TRY(ctx, cudaFuncSetAttribute(k_pf_local2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Local2Smem)));
    LAUNCH(ctx, k_pf_local2, ntiles, NTHR, sizeof(Local2Smem), d_dir, nc, nr, ntx, lin0, w, S, LR, boff, eho);
    uint32_t n_basin = 0;
    CU_TRY(ctx, cudaMemcpyAsync(&n_basin, S + lin0, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    // D
    const bool small_graph = n_nodes <= BDS_MAX_NODES;
    int32_t *flag, *cmap;
    WMF_TRY(sc.alloc(&flag, (size_t)n_nodes)); WMF_TRY(sc.alloc(&cmap, (size_t)n_nodes));
    LAUNCH(ctx, k_pf_exit_flag, GN, B, 0, n_nodes, bdown, flag);
    int64_t n_exit = n_nodes;   // (small graphs read the count on the device: no host round trip)
    WMF_TRY(wmf_device_exclusive_scan_i32(ctx, sc, flag, cmap, n_nodes, small_graph ? nullptr : &n_exit));
    LAUNCH(ctx, k_pf_bd_init, GN, B, 0, n_nodes, bdown, bnext, boff, eho, cmap, nxA, dpA, leaf /* reused as the age */);
    if (small_graph) LAUNCH(ctx, k_pf_bd_jump_small, 1, BDS_THREADS, 0, n_nodes, cmap, flag, nxA, dpA, nxB, dpB);
    const int GE = grid_for(std::max<int64_t>(n_exit, 1), B);
    for (int round = 0; round < 40 && !small_graph; ++round) {
        CU_TRY(ctx, cudaMemsetAsync(more, 0, sizeof(int), ctx->stream));
        LAUNCH(ctx, k_pf_bd_jump, GE, B, 0, (int)n_exit, nxA, dpA, nxB, dpB, leaf, more);
        std::swap(nxA, nxB);
        std::swap(dpA, dpB);
        int h_more = 0;
        CU_TRY(ctx, cudaMemcpyAsync(&h_more, more, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        if (!h_more) break;  // (what is still pending after 40 doublings is a cycle across tiles: not in the basin)
    }
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (n_basin == 0 || n_basin > cells) WMF_FAIL(ctx, WMF_ERR_TOPOLOGY, "basin_find: inconsistent basin size %u", n_basin);
    if (dumpdir) {
        dump(ctx, dumpdir, "pterm", pterm, 4 * (size_t)n_nodes); dump(ctx, dumpdir, "bdown", bdown, 4 * (size_t)n_nodes);
        dump(ctx, dumpdir, "lsize", lsize, 4 * (size_t)n_nodes); dump(ctx, dumpdir, "bnext", bnext, 4 * (size_t)n_nodes);
        dump(ctx, dumpdir, "w", w, 8 * (size_t)n_nodes); dump(ctx, dumpdir, "boff", boff, 4 * (size_t)n_nodes);
        LAUNCH(ctx, k_pf_expand, GN, B, 0, n_nodes, bdown, cmap, nxA, dpA, nxB, dpB);
        dump(ctx, dumpdir, "eho", eho, 8 * (size_t)n_nodes); dump(ctx, dumpdir, "nx", nxB, 4 * (size_t)n_nodes);
        dump(ctx, dumpdir, "dp", dpB, 8 * (size_t)n_nodes); dump(ctx, dumpdir, "S", S, 4 * cells); dump(ctx, dumpdir, "LR", LR, 8 * cells);
    }
    // E
    const uint32_t n = n_basin;
    uint2 *items;
    uint32_t *k1, *v1;
    WMF_TRY(sc.alloc(&items, (size_t)n)); WMF_TRY(sc.alloc(&k1, (size_t)n)); WMF_TRY(sc.alloc(&v1, (size_t)n));
    uint32_t *k0 = reinterpret_cast<uint32_t *>(items), *v0 = k0 + n;   // the pairs are dead after the first pass: reused as split arrays
    // a slot nobody writes shows up as depth 2^28-1.  (Dropping this fill for a written-count check was measured: basin_find gets
    // SLOWER, 25.5 -> 26.2 ms at C3 -- the scattered 8-byte writes of k_pf_emit are cheaper on lines the fill has just touched.)
    CU_TRY(ctx, cudaMemsetAsync(items, 0xff, sizeof(uint2) * (size_t)n, ctx->stream));
    LAUNCH(ctx, k_pf_emit, ntiles, NTHR, 0, LR, nc, nr, ntx, cmap, (int)n_exit, nxA, dpA, n, items, stat);
    LAUNCH(ctx, k_pf_outlet_check, 1, 1, 0, d_dir, LR, nc, nr, ntx, lin0, cmap, nxA, stat);
    unsigned h_stat[4] = {0, 0, 0, 0};
    CU_TRY(ctx, cudaMemcpyAsync(h_stat, stat, sizeof h_stat, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (h_stat[2]) WMF_FAIL(ctx, WMF_ERR_TOPOLOGY, "basin_find: the DIR map is cyclic -- the cell the outlet drains to drains back into the outlet");
    if (h_stat[1]) WMF_FAIL(ctx, WMF_ERR_TOPOLOGY, "basin_find: %u cells fell outside the preorder range (internal error or depth >= 2^28)", h_stat[1]);
    const unsigned maxdepth = h_stat[0];
    if (dumpdir) dump(ctx, dumpdir, "items0", items, 8 * (size_t)n);
    // sort by depth, stable: the order inside a level stays the preorder
    int bits = 0;
    while (bits < 28 && (maxdepth >> bits) != 0) ++bits;
    const int passes = std::max(1, (bits + 7) / 8);   // (at least one: it also splits the pairs into keys and values)
    const int nblk = (int)(((size_t)n + RS_TILE - 1) / RS_TILE);
    int32_t *hist = nullptr, *hist_scan = nullptr;
    WMF_TRY(sc.alloc(&hist, (size_t)256 * nblk));
    WMF_TRY(sc.alloc(&hist_scan, (size_t)256 * nblk));
    uint32_t *kin = k0, *vin = v0, *kout = k1, *vout = v1;
    CU_TRY(ctx, cudaFuncSetAttribute(k_rs_scatter<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RsSmem)));
    CU_TRY(ctx, cudaFuncSetAttribute(k_rs_scatter<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RsSmem)));
    for (int p = 0; p < passes; ++p) {
        uint32_t *vdst = (p == passes - 1) ? reinterpret_cast<uint32_t *>(ctx->q_lin) : vout;
        if (p == 0) LAUNCH(ctx, k_rs_hist<true>, nblk, RS_THREADS, 0, kin, n, 8 * p, nblk, hist);
        else LAUNCH(ctx, k_rs_hist<false>, nblk, RS_THREADS, 0, kin, n, 8 * p, nblk, hist);
        WMF_TRY(wmf_device_exclusive_scan_i32(ctx, sc, hist, hist_scan, (int64_t)256 * nblk, nullptr));
        if (p == 0) LAUNCH(ctx, k_rs_scatter<true>, nblk, RS_THREADS, sizeof(RsSmem), kin, vin, n, 8 * p, nblk, hist_scan, kout, vdst);
        else LAUNCH(ctx, k_rs_scatter<false>, nblk, RS_THREADS, sizeof(RsSmem), kin, vin, n, 8 * p, nblk, hist_scan, kout, vdst);
        std::swap(kin, kout);
        vin = vdst;
        vout = (vdst == v1) ? v0 : v1;
    }
    // F
    const int fblk = (int)(((size_t)n + FN_TILE - 1) / FN_TILE);
    int32_t *blk_sum, *blk_scan;
    WMF_TRY(sc.alloc(&blk_sum, (size_t)fblk));
    WMF_TRY(sc.alloc(&blk_scan, (size_t)fblk));
    LAUNCH(ctx, k_pf_final_sums, fblk, FN_THREADS, 0, kin, n, blk_sum);
    WMF_TRY(wmf_device_exclusive_scan_i32(ctx, sc, blk_sum, blk_scan, fblk, nullptr));
    LAUNCH(ctx, k_pf_final_write, fblk, FN_THREADS, 0, kin, ctx->q_lin, n, blk_scan, S, ctx->q_par, ctx->q_acum, stat);
    CU_TRY(ctx, cudaMemcpyAsync(h_stat, stat, sizeof h_stat, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CU_TRY(ctx, cudaGetLastError());
    if (h_stat[1]) WMF_FAIL(ctx, WMF_ERR_TOPOLOGY, "basin_find: the child counts do not add up to the basin size (internal error)");
    *n_out = (int32_t)n;
    *depth_out = (int32_t)maxdepth + 1;
    return WMF_OK;
}

// basin_acum for the very table basin_find produced (checked drain id by drain id); returns 1 in *used when it applied
int wmf_pathfind_acum_cached(wmf_ctx *ctx, Scope &sc, const int32_t *d_basin_f, int n, int32_t *d_acum, int *used) {
    *used = 0;
    if (!ctx->q_acum_valid || n != ctx->find_n || !ctx->q_acum) return WMF_OK;
    int *bad;
    WMF_TRY(sc.alloc(&bad, 1));
    CU_TRY(ctx, cudaMemsetAsync(bad, 0, sizeof(int), ctx->stream));
    LAUNCH(ctx, pf::k_pf_acum_cached, grid_for(n, 256), 256, 0, d_basin_f, ctx->q_par, ctx->q_acum, n, d_acum, bad);
    int h_bad = 1;
    CU_TRY(ctx, cudaMemcpyAsync(&h_bad, bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    *used = h_bad ? 0 : 1;
    return WMF_OK;
}
