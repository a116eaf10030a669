// Map rebuild kernels (map_generation side of the hot path):
//   uam_dem_mask            image > threshold | image == -9999            data_manager.py:14-17
//   uam_rasterize_occupancy Map.collides at cell centres, fp64 bit-exact   map.py:41-43, quadratic_obstacle.py:89-94
//   uam_rasterize_layers    per-region penalty field at cell centres       problem.py:72-80 (unweighted)
//   uam_edt                 exact squared Euclidean distance transform     (build-defined extension)
//
// Rasterisation is tile-culled: a CTA owns a 16 x 64 cell tile, first compacts (in shape order, so float64
// sums keep the reference's order) the shapes that can be non-zero / can contain a point anywhere in the tile
// by a conservative separating-inequality test on the tile corners, then evaluates only those per cell.
// A culled shape contributes exactly 0 (psi has a zero factor) or "not contained", so culling never changes
// a bit of the result.
#include <algorithm>
#include <cmath>

#include "uam_internal.cuh"

namespace {

#define UAM_TILE_H 16
#define UAM_TILE_W 64
#define UAM_LIST_CAP 1024

// -------------------------------------------------------------------------------------------------------
// DEM mask
// -------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
uam_k_dem_mask(const float* __restrict__ img, long long n, float thr, int eq_mode, uint8_t* __restrict__ mask) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long n4 = n >> 2;
    const bool aligned = ((((uintptr_t)img) & 15) == 0) && ((((uintptr_t)mask) & 3) == 0);
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (aligned) {
        for (; i < n4; i += stride) {
            const float4 v = __ldcs(reinterpret_cast<const float4*>(img) + i);
            uchar4 m;
            m.x = eq_mode ? (v.x == thr) : (v.x > thr);
            m.y = eq_mode ? (v.y == thr) : (v.y > thr);
            m.z = eq_mode ? (v.z == thr) : (v.z > thr);
            m.w = eq_mode ? (v.w == thr) : (v.w > thr);
            __stcs(reinterpret_cast<uchar4*>(mask) + i, m);
        }
        for (long long j = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride)
            mask[j] = eq_mode ? (img[j] == thr) : (img[j] > thr);
    } else {
        for (; i < n; i += stride) mask[i] = eq_mode ? (img[i] == thr) : (img[i] > thr);
    }
}

// -------------------------------------------------------------------------------------------------------
// tile culling
// -------------------------------------------------------------------------------------------------------
// Ordered compaction of the candidates at positions [s_begin, s_end) that survive the tile test into list[0..n)
// (n <= CAP).  A candidate is shape cand[p] (or shape p itself when cand is NULL).  Returns the next position to
// continue from.  Must be called by all threads of a 256-thread CTA.
// psic (nullable: the occupancy pass has no use for it): a shape whose psi(centre) is 0 or NaN is never culled -- the
// reference's psi(x) / psi(centre) is 0/0 = NaN at EVERY point for it (problem.py:79), not only near the shape.
__device__ int uam_cull_shapes(const UamEdge* __restrict__ edges, const UamShape* __restrict__ shapes,
                               const int* __restrict__ cand, int s_begin, int s_end, double xa, double xb, double ya,
                               double yb, double thr, int* list, int* n_out, int* warp_cnt, const double* __restrict__ psic = nullptr) {
    int n = 0;
    int s0 = s_begin;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    while (s0 < s_end && n + (int)blockDim.x <= UAM_LIST_CAP) {
        const int pos = s0 + threadIdx.x;
        bool keep = false;
        int s = 0;
        if (pos < s_end) {
            s = cand ? __ldg(cand + pos) : pos;
            const int4 meta = __ldg(reinterpret_cast<const int4*>(&shapes[s].e0));
            keep = true;
            bool never = false;
            if (psic && meta.w) {
                const double pc = __ldg(psic + s);
                never = pc == 0.0 || pc != pc;
            }
            for (int i = meta.x; i < meta.y && keep && !never; ++i) {
                const UamEdge r = uam_load_edge(edges + i);
                if (uam_edge_excludes_tile(r, xa, xb, ya, yb, thr)) keep = false;
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_cnt[warp] = __popc(bal);
        __syncthreads();
        int off = n, tot = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
            const int c = warp_cnt[w];
            if (w < warp) off += c;
            tot += c;
        }
        if (keep) list[off + __popc(bal & ((1u << lane) - 1u))] = s;
        n += tot;
        s0 += blockDim.x;
        __syncthreads();
    }
    *n_out = n;
    return min(s0, s_end);
}

// Coarse pass: one CTA per `size` x `size`-cell block keeps, in shape order, the shapes of [s_begin, s_end) that can matter
// anywhere in the block; the per-tile kernels then only test those.  Hierarchical: with parent lists (blocks of parent_size
// cells, parent_gx of them per row) a block only tests its parent's survivors -- 2048-cell blocks first, then the 256-cell
// supertiles: 4096 shapes x 4096 supertiles cost 0.7 M tests instead of 16.8 M (ncu r02: 0.46 ms -> the noise).
#define UAM_SUPER 256
#define UAM_SUPER0 2048
__global__ void __launch_bounds__(256)
uam_k_cull_coarse(const UamEdge* __restrict__ edges, const UamShape* __restrict__ shapes, int s_begin, int s_end,
                  double thr, int H, int W, double x0, double dx, double y0, double dy, int size,
                  const int* __restrict__ parent_list, const int* __restrict__ parent_count, int parent_size, int parent_gx,
                  int* __restrict__ out_list, int* __restrict__ out_count, const double* __restrict__ psic) {
    __shared__ int list[UAM_LIST_CAP];
    __shared__ int warp_cnt[8];
    const int sup = blockIdx.y * gridDim.x + blockIdx.x;
    const int j0 = blockIdx.x * size, i0 = blockIdx.y * size;
    const int j1 = min(j0 + size, W), i1 = min(i0 + size, H);
    const double xe0 = x0 + j0 * dx, xe1 = x0 + j1 * dx, ye0 = y0 + i0 * dy, ye1 = y0 + i1 * dy;
    const double xa = fmin(xe0, xe1), xb = fmax(xe0, xe1), ya = fmin(ye0, ye1), yb = fmax(ye0, ye1);
    int* dst = out_list + (size_t)sup * (s_end - s_begin);
    const int* cand = nullptr;
    int lo = s_begin, hi = s_end;
    if (parent_list) {
        const int p = (i0 / parent_size) * parent_gx + j0 / parent_size;
        cand = parent_list + (size_t)p * (s_end - s_begin);
        lo = 0;
        hi = parent_count[p];
    }
    int total = 0;
    int s_next = lo;
    while (s_next < hi) {
        int n;
        s_next = uam_cull_shapes(edges, shapes, cand, s_next, hi, xa, xb, ya, yb, thr, list, &n, warp_cnt, psic);
        for (int t = threadIdx.x; t < n; t += blockDim.x) dst[total + t] = list[t];
        total += n;
        __syncthreads();
    }
    if (threadIdx.x == 0) out_count[sup] = total;
}

__device__ __forceinline__ int uam_supertile_of_block() {
    // fine tiles are 16 rows x 64 columns: 16 tile-rows x 4 tile-columns per supertile
    const int sup_x = (gridDim.x + 3) / 4;
    return (blockIdx.y / (UAM_SUPER / UAM_TILE_H)) * sup_x + blockIdx.x / (UAM_SUPER / UAM_TILE_W);
}

__device__ __forceinline__ double uam_cell_centre(int j, double x0, double dx) {
    return __dadd_rn(x0, __dmul_rn(__dadd_rn((double)j, 0.5), dx));     // x0 + (j + 1/2) * dx
}

// -------------------------------------------------------------------------------------------------------
// occupancy
// -------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
uam_k_rasterize_occupancy(const UamEdge* __restrict__ edges, const UamShape* __restrict__ shapes, int n_obs, int H,
                          int W, double x0, double dx, double y0, double dy, const int* __restrict__ coarse_list,
                          const int* __restrict__ coarse_count, uint8_t* __restrict__ occ) {
    __shared__ int list[UAM_LIST_CAP];
    __shared__ int warp_cnt[8];
    const int tj0 = blockIdx.x * UAM_TILE_W, ti0 = blockIdx.y * UAM_TILE_H;
    const int tj1 = min(tj0 + UAM_TILE_W, W), ti1 = min(ti0 + UAM_TILE_H, H);
    const double xe0 = x0 + tj0 * dx, xe1 = x0 + tj1 * dx, ye0 = y0 + ti0 * dy, ye1 = y0 + ti1 * dy;
    const double xa = fmin(xe0, xe1), xb = fmax(xe0, xe1), ya = fmin(ye0, ye1), yb = fmax(ye0, ye1);
    const int i = ti0 + (threadIdx.x >> 4);
    const int j = tj0 + ((threadIdx.x & 15) << 2);
    const double y = uam_cell_centre(i, y0, dy);
    double x[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) x[c] = uam_cell_centre(j + c, x0, dx);
    bool in_any[4] = {false, false, false, false};
    const int sup = uam_supertile_of_block();
    const int* cand = coarse_list + (size_t)sup * n_obs;
    const int n_cand = coarse_count[sup];
    int s_next = 0;
    while (s_next < n_cand) {
        int n;
        s_next = uam_cull_shapes(edges, shapes, cand, s_next, n_cand, xa, xb, ya, yb, 1e-14, list, &n, warp_cnt);
        for (int t = 0; t < n; ++t) {
            const int s = list[t];
            const int4 meta = __ldg(reinterpret_cast<const int4*>(&shapes[s].e0));
            bool in[4] = {true, true, true, true};
            for (int e = meta.x; e < meta.y; ++e) {
                const UamEdge r = uam_load_edge(edges + e);
#pragma unroll
                for (int c = 0; c < 4; ++c) in[c] = in[c] && (uam_h_exact(r, x[c], y) <= 1e-14);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) in_any[c] = in_any[c] || in[c];
        }
        __syncthreads();
    }
    if (i < H) {
        uint8_t* row = occ + (size_t)i * W;
        if (j + 3 < W && (W & 3) == 0 && ((((uintptr_t)occ) & 3) == 0)) {
            uchar4 m;
            m.x = in_any[0]; m.y = in_any[1]; m.z = in_any[2]; m.w = in_any[3];
            *reinterpret_cast<uchar4*>(row + j) = m;
        } else {
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (j + c < W) row[j + c] = in_any[c];
        }
    }
}

// -------------------------------------------------------------------------------------------------------
// row intervals of a convex shape (scanline form of the two rasterisers)
// -------------------------------------------------------------------------------------------------------
// Along one raster row the fp64 value of an inequality at the cell centres, evaluated by uam_h_exact in the reference's
// operation order, is a MONOTONE function of the column (every rounding step is monotone: x_j = x0 + (j + 1/2) dx,
// x_j - Ax, the product with a constant, the difference with a constant), and for an ellipse it is monotone on either side
// of the centre column ((x_j - cx) / r1 is monotone and its square is monotone in its magnitude).  So the set of cells of
// the row where the predicate "h <= thr" (contains, quadratic_obstacle.py:89-94) or "h - e < 0" (psi != 0,
// quadratic_obstacle.py:33-35) holds is a prefix, a suffix or -- for the ellipse -- the union of a suffix and a prefix
// around the centre: one interval per inequality, found by bisection WITH THE EXACT PREDICATE.  The cells of a convex
// shape on a row are the intersection of its inequalities' intervals.  No approximate geometry decides a cell, so the
// result has the bits of the per-cell evaluation; the gain is ~8 predicate evaluations per (row, inequality) instead of
// one per cell.  Needs finite records (uam_map_set_shapes checks) -- NaN breaks monotonicity; callers fall back to the
// per-cell kernels otherwise.
template <int KIND>      // 0: h <= t (contains)   1: h - t < 0 (psi != 0 for enlargement t)
__device__ __forceinline__ bool uam_row_pred(const UamEdge& r, double x, double y, double t) {
    const double h = uam_h_exact(r, x, y);
    return KIND == 0 ? (h <= t) : (__dsub_rn(h, t) < 0.0);
}
// cells of [a, b) (columns relative to jbase) where the predicate holds, the predicate being monotone on [a, b)
template <int KIND>
__device__ __forceinline__ void uam_monotone_cells(const UamEdge& r, double y, double x0, double dx, int jbase, int a, int b,
                                                   double t, int& lo, int& hi) {
    lo = hi = a;
    if (a >= b) return;
    const bool pa = uam_row_pred<KIND>(r, uam_cell_centre(jbase + a, x0, dx), y, t);
    const bool pb = (b - 1 == a) ? pa : uam_row_pred<KIND>(r, uam_cell_centre(jbase + b - 1, x0, dx), y, t);
    if (pa && pb) { hi = b; return; }
    if (!pa && !pb) return;
    int l = a, h = b - 1;                               // pred(l) == pa, pred(h) == pb, pa != pb
    while (h - l > 1) {
        const int m = (l + h) >> 1;
        if (uam_row_pred<KIND>(r, uam_cell_centre(jbase + m, x0, dx), y, t) == pa) l = m; else h = m;
    }
    if (pa) { lo = a; hi = l + 1; } else { lo = h; hi = b; }
}
// [lo, hi) := [lo, hi) intersected with the cells of [0, n) where inequality r holds
template <int KIND>
__device__ __forceinline__ void uam_row_interval_edge(const UamEdge& r, double y, double x0, double dx, int jbase, int n, double t,
                                                      int& lo, int& hi) {
    int l, h;
    if ((int)r.kind == UAM_EDGE_ELLIPSE) {
        // split where x_j - cx changes sign (monotone in j)
        int a = 0, b = n;
        const bool s0 = uam_cell_centre(jbase, x0, dx) >= r.p0;
        const bool s1 = uam_cell_centre(jbase + n - 1, x0, dx) >= r.p0;
        int split = n;                                  // first column on the other side of the centre than column 0
        if (s0 != s1) {
            int ll = 0, hh = n - 1;
            while (hh - ll > 1) {
                const int m = (ll + hh) >> 1;
                if ((uam_cell_centre(jbase + m, x0, dx) >= r.p0) == s0) ll = m; else hh = m;
            }
            split = hh;
        }
        int l1, h1, l2, h2;
        uam_monotone_cells<KIND>(r, y, x0, dx, jbase, a, split, t, l1, h1);
        uam_monotone_cells<KIND>(r, y, x0, dx, jbase, split, b, t, l2, h2);
        if (h1 > l1 && h2 > l2) { l = l1; h = h2; }      // a suffix of [0, split) and a prefix of [split, n): contiguous
        else if (h1 > l1) { l = l1; h = h1; }
        else { l = l2; h = h2; }
    } else {
        uam_monotone_cells<KIND>(r, y, x0, dx, jbase, 0, n, t, l, h);
    }
    lo = max(lo, l);
    hi = min(hi, h);
}

// One CTA per 256 x 256-cell supertile, thread r = row r of the supertile.  For every candidate obstacle of the supertile
// (coarse list) the thread finds the row's cell interval and sets the bits in its row of a shared-memory bitmap (32 bytes
// per row: no atomics, the row is the thread's own); the bitmap is then expanded to bytes with 16-byte stores.  Every cell of
// the raster is written exactly once.
__global__ void __launch_bounds__(256)
uam_k_occupancy_scan(const UamEdge* __restrict__ edges, const UamShape* __restrict__ shapes, int n_obs, int H, int W,
                     double x0, double dx, double y0, double dy, const int* __restrict__ coarse_list,
                     const int* __restrict__ coarse_count, uint8_t* __restrict__ occ) {
    __shared__ unsigned bits[UAM_SUPER][UAM_SUPER / 32 + 1];        // + 1: rows start in different banks
    const int sup = blockIdx.y * gridDim.x + blockIdx.x;
    const int j0 = blockIdx.x * UAM_SUPER, i0 = blockIdx.y * UAM_SUPER;
    const int ncols = min(UAM_SUPER, W - j0);
    const int row = threadIdx.x;
    const int i = i0 + row;
    unsigned w[UAM_SUPER / 32];
#pragma unroll
    for (int k = 0; k < UAM_SUPER / 32; ++k) w[k] = 0u;
    if (i < H) {
        const double y = uam_cell_centre(i, y0, dy);
        const int* cand = coarse_list + (size_t)sup * n_obs;
        const int n_cand = coarse_count[sup];
        for (int c = 0; c < n_cand; ++c) {
            const int s = __ldg(cand + c);
            const int4 meta = __ldg(reinterpret_cast<const int4*>(&shapes[s].e0));
            int lo = 0, hi = ncols;
            for (int e = meta.x; e < meta.y && lo < hi; ++e) {
                const UamEdge r = uam_load_edge(edges + e);
                uam_row_interval_edge<0>(r, y, x0, dx, j0, ncols, 1e-14, lo, hi);
            }
            if (lo < hi) {
#pragma unroll
                for (int k = 0; k < UAM_SUPER / 32; ++k) {
                    const int a = max(lo - 32 * k, 0), b = min(hi - 32 * k, 32);
                    if (a < b) w[k] |= (b - a == 32) ? 0xffffffffu : (((1u << (b - a)) - 1u) << a);
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < UAM_SUPER / 32; ++k) bits[row][k] = w[k];
    __syncthreads();
    // expand: 16 cells (one 16-byte store) per thread per trip; a warp covers two rows of the supertile per trip
    const bool vec = (W & 15) == 0 && ((((uintptr_t)occ) & 15) == 0);
    for (int t = threadIdx.x; t < UAM_SUPER * (UAM_SUPER / 16); t += blockDim.x) {
        const int r = t >> 4, q = t & 15;               // row, 16-cell group
        const int ii = i0 + r, jj = j0 + q * 16;
        if (ii >= H || jj >= W) continue;
        const unsigned m = (bits[r][q >> 1] >> ((q & 1) * 16)) & 0xffffu;
        uint8_t* dst = occ + (size_t)ii * W + jj;
        if (vec && jj + 16 <= W) {
            uint4 v;
            v.x = ((m >> 0) & 1u) | (((m >> 1) & 1u) << 8) | (((m >> 2) & 1u) << 16) | (((m >> 3) & 1u) << 24);
            v.y = ((m >> 4) & 1u) | (((m >> 5) & 1u) << 8) | (((m >> 6) & 1u) << 16) | (((m >> 7) & 1u) << 24);
            v.z = ((m >> 8) & 1u) | (((m >> 9) & 1u) << 8) | (((m >> 10) & 1u) << 16) | (((m >> 11) & 1u) << 24);
            v.w = ((m >> 12) & 1u) | (((m >> 13) & 1u) << 8) | (((m >> 14) & 1u) << 16) | (((m >> 15) & 1u) << 24);
            __stcs(reinterpret_cast<uint4*>(dst), v);
        } else {
            for (int c = 0; c < 16 && jj + c < W; ++c) dst[c] = (m >> c) & 1u;
        }
    }
}

// -------------------------------------------------------------------------------------------------------
// penalty layers (smooth psi only)
// -------------------------------------------------------------------------------------------------------
struct UamRegionRanges2 {
    int begin[UAM_MAX_REGIONS + 1];
};

__global__ void __launch_bounds__(256)
uam_k_rasterize_layers(const UamEdge* __restrict__ edges, const UamShape* __restrict__ shapes,
                       const double* __restrict__ psic, UamRegionRanges2 rr, int n_regions, int H, int W, double x0,
                       double dx, double y0, double dy, double e, const int* __restrict__ coarse_list,
                       const int* __restrict__ coarse_count, int n_super, float* __restrict__ layers) {
    __shared__ int list[UAM_LIST_CAP];
    __shared__ int warp_cnt[8];
    const int tj0 = blockIdx.x * UAM_TILE_W, ti0 = blockIdx.y * UAM_TILE_H;
    const int tj1 = min(tj0 + UAM_TILE_W, W), ti1 = min(ti0 + UAM_TILE_H, H);
    const double xe0 = x0 + tj0 * dx, xe1 = x0 + tj1 * dx, ye0 = y0 + ti0 * dy, ye1 = y0 + ti1 * dy;
    const double xa = fmin(xe0, xe1), xb = fmax(xe0, xe1), ya = fmin(ye0, ye1), yb = fmax(ye0, ye1);
    const int i = ti0 + (threadIdx.x >> 4);
    const int j = tj0 + ((threadIdx.x & 15) << 2);
    const double y = uam_cell_centre(i, y0, dy);
    double x[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) x[c] = uam_cell_centre(j + c, x0, dx);
    const size_t plane = (size_t)H * W;
    const int sup = uam_supertile_of_block();
    for (int r = 0; r < n_regions; ++r) {
        double tot[4] = {0.0, 0.0, 0.0, 0.0};
        // region r's coarse lists start after those of regions 0..r-1: n_super * (shapes before r) entries
        const int* cand = coarse_list + (size_t)n_super * (rr.begin[r] - rr.begin[0]) + (size_t)sup * (rr.begin[r + 1] - rr.begin[r]);
        const int s_end = coarse_count[r * n_super + sup];
        int s_next = 0;
        while (s_next < s_end) {
            int n;
            // psi != 0 needs h_i - e < 0 for every i: cull when some h_i > e on the whole tile
            s_next = uam_cull_shapes(edges, shapes, cand, s_next, s_end, xa, xb, ya, yb, e, list, &n, warp_cnt, psic);
            for (int t = 0; t < n; ++t) {
                const int s = list[t];
                const int4 meta = __ldg(reinterpret_cast<const int4*>(&shapes[s].e0));
                double psi[4] = {1.0, 1.0, 1.0, 1.0};
                for (int ed = meta.x; ed < meta.y; ++ed) {
                    const UamEdge rcd = uam_load_edge(edges + ed);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const double m = fmin(__dsub_rn(uam_h_exact(rcd, x[c], y), e), 0.0);
                        psi[c] = __dmul_rn(psi[c], __dmul_rn(m, m));
                    }
                }
                const double pc = meta.w ? __ldg(psic + s) : 1.0;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (meta.w) {
                        if (psi[c] != 0.0 || pc == 0.0 || pc != pc) tot[c] = __dadd_rn(tot[c], __ddiv_rn(psi[c], pc));
                    } else {
                        tot[c] = __dadd_rn(tot[c], psi[c]);
                    }
                }
            }
            __syncthreads();
        }
        if (i < H) {
            float* row = layers + (size_t)r * plane + (size_t)i * W;
            if (j + 3 < W && (W & 3) == 0 && ((((uintptr_t)layers) & 15) == 0)) {
                __stcs(reinterpret_cast<float4*>(row + j), make_float4((float)tot[0], (float)tot[1], (float)tot[2], (float)tot[3]));
            } else {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (j + c < W) row[j + c] = (float)tot[c];
            }
        }
    }
}

// Scanline form of uam_k_rasterize_layers: same tiles, same ordered shape lists, same per-cell arithmetic -- but before a
// chunk of 16 listed shapes is evaluated, the CTA finds for each (shape, tile row) the interval of cells where psi can be
// non-zero (all h_i - e < 0: uam_row_interval_edge<1>, exact predicate), and a thread evaluates a shape only if its four
// cells touch the row's interval.  A skipped cell has a zero factor in psi, i.e. contributes the exact +0 the per-cell
// kernel adds (the host checks that no product can overflow first), so the bits do not change.  Shapes whose psi(centre)
// is 0 or NaN (0/0 reaches every cell in the reference) are never skipped.
__global__ void __launch_bounds__(256)
uam_k_layers_scan(const UamEdge* __restrict__ edges, const UamShape* __restrict__ shapes, const double* __restrict__ psic,
                  UamRegionRanges2 rr, int n_regions, int H, int W, double x0, double dx, double y0, double dy, double e,
                  const int* __restrict__ coarse_list, const int* __restrict__ coarse_count, int n_super,
                  float* __restrict__ layers) {
    __shared__ int list[UAM_LIST_CAP];
    __shared__ int warp_cnt[8];
    __shared__ unsigned short iv[16][16];               // [shape of the chunk][tile row] = lo | hi << 8 (tile columns)
    const int tj0 = blockIdx.x * UAM_TILE_W, ti0 = blockIdx.y * UAM_TILE_H;
    const int tj1 = min(tj0 + UAM_TILE_W, W), ti1 = min(ti0 + UAM_TILE_H, H);
    const int ncols = tj1 - tj0;
    const double xe0 = x0 + tj0 * dx, xe1 = x0 + tj1 * dx, ye0 = y0 + ti0 * dy, ye1 = y0 + ti1 * dy;
    const double xa = fmin(xe0, xe1), xb = fmax(xe0, xe1), ya = fmin(ye0, ye1), yb = fmax(ye0, ye1);
    const int trow = threadIdx.x >> 4, tq = threadIdx.x & 15;
    const int i = ti0 + trow;
    const int c0 = tq << 2;
    const int j = tj0 + c0;
    const double y = uam_cell_centre(i, y0, dy);
    double x[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) x[c] = uam_cell_centre(j + c, x0, dx);
    const size_t plane = (size_t)H * W;
    const int sup = uam_supertile_of_block();
    for (int r = 0; r < n_regions; ++r) {
        double tot[4] = {0.0, 0.0, 0.0, 0.0};
        const int* cand = coarse_list + (size_t)n_super * (rr.begin[r] - rr.begin[0]) + (size_t)sup * (rr.begin[r + 1] - rr.begin[r]);
        const int s_end = coarse_count[r * n_super + sup];
        int s_next = 0;
        while (s_next < s_end) {
            int n;
            s_next = uam_cull_shapes(edges, shapes, cand, s_next, s_end, xa, xb, ya, yb, e, list, &n, warp_cnt, psic);
            for (int t0 = 0; t0 < n; t0 += 16) {
                {   // interval of (shape t0 + tq, tile row trow)
                    const int t = t0 + tq;
                    int lo = 0, hi = 0;
                    if (t < n && i < H) {
                        const int s = list[t];
                        const int4 meta = __ldg(reinterpret_cast<const int4*>(&shapes[s].e0));
                        hi = ncols;
                        for (int ed = meta.x; ed < meta.y && lo < hi; ++ed) {
                            const UamEdge rcd = uam_load_edge(edges + ed);
                            uam_row_interval_edge<1>(rcd, y, x0, dx, tj0, ncols, e, lo, hi);
                        }
                        if (meta.w) {
                            const double pc = __ldg(psic + s);
                            if (pc == 0.0 || pc != pc) { lo = 0; hi = ncols; }
                        }
                        if (lo >= hi) lo = hi = 0;
                    }
                    iv[tq][trow] = (unsigned short)(lo | (hi << 8));
                }
                __syncthreads();
                const int nt = min(16, n - t0);
                for (int k = 0; k < nt; ++k) {
                    const unsigned v = iv[k][trow];
                    const int lo = (int)(v & 255u), hi = (int)(v >> 8);
                    if (!(c0 < hi && c0 + 4 > lo)) continue;
                    const int s = list[t0 + k];
                    const int4 meta = __ldg(reinterpret_cast<const int4*>(&shapes[s].e0));
                    double psi[4] = {1.0, 1.0, 1.0, 1.0};
                    for (int ed = meta.x; ed < meta.y; ++ed) {
                        const UamEdge rcd = uam_load_edge(edges + ed);
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const double m = fmin(__dsub_rn(uam_h_exact(rcd, x[c], y), e), 0.0);
                            psi[c] = __dmul_rn(psi[c], __dmul_rn(m, m));
                        }
                    }
                    const double pc = meta.w ? __ldg(psic + s) : 1.0;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (meta.w) {
                            if (psi[c] != 0.0 || pc == 0.0 || pc != pc) tot[c] = __dadd_rn(tot[c], __ddiv_rn(psi[c], pc));
                        } else {
                            tot[c] = __dadd_rn(tot[c], psi[c]);
                        }
                    }
                }
                __syncthreads();
            }
            __syncthreads();
        }
        if (i < H) {
            float* row = layers + (size_t)r * plane + (size_t)i * W;
            if (j + 3 < W && (W & 3) == 0 && ((((uintptr_t)layers) & 15) == 0)) {
                __stcs(reinterpret_cast<float4*>(row + j), make_float4((float)tot[0], (float)tot[1], (float)tot[2], (float)tot[3]));
            } else {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (j + c < W) row[j + c] = (float)tot[c];
            }
        }
    }
}

// Row form of the layer rasteriser (the default): one CTA per 256 x 256-cell supertile, one warp per raster row of it at a
// time.  For every candidate shape of the supertile (coarse list, shape order) the warp samples the row at 32 columns
// (first and last included, spacing <= 9): the value of an inequality is monotone along the row (see above), so the cells
// where psi can be non-zero lie strictly between the samples next to the first / last sample that satisfies "h - e < 0" --
// a coarse span per inequality from ONE predicate evaluation per lane, intersected over the shape's inequalities (an ellipse,
// monotone on either side of its centre, also gets the three columns next to its centre checked).  Only the cells of the
// span are evaluated, 32 at a time (lane = column mod 32), with the per-cell arithmetic of uam_k_rasterize_layers, and
// accumulated in shape order in a shared-memory row of doubles; the row is stored once, coalesced.  A cell outside the span
// has a zero factor in psi: the exact +0 the per-cell kernel adds (host-side overflow guard as for the other scan forms).
__device__ __noinline__ void uam_layers_rows_sampled(double (*acc_all)[UAM_SUPER], const UamEdge* __restrict__ edges,
                                                        const UamShape* __restrict__ shapes, const double* __restrict__ psic,
                                                        const UamRegionRanges2& rr, int n_regions, int H, int W, double x0,
                                                        double dx, double y0, double dy, double e,
                                                        const int* __restrict__ coarse_list, const int* __restrict__ coarse_count,
                                                        int n_super, float* __restrict__ layers, int sup, int i0, int rows) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    double* acc = acc_all[warp];
    const int j0 = blockIdx.x * UAM_SUPER;
    const int ncols = min(UAM_SUPER, W - j0);
    const size_t plane = (size_t)H * W;
    // sample columns: 0 .. ncols - 1 inclusive
    const int sj = (lane * (ncols - 1)) / 31;
    const int sj_prev = lane > 0 ? ((lane - 1) * (ncols - 1)) / 31 : -1;
    const int sj_next = lane < 31 ? ((lane + 1) * (ncols - 1)) / 31 : ncols;
    const double xs = uam_cell_centre(j0 + sj, x0, dx);
    for (int row = warp; row < rows; row += nwarps) {
        const int i = i0 + row;
        if (i >= H) break;
        const double y = uam_cell_centre(i, y0, dy);
        for (int r = 0; r < n_regions; ++r) {
            const int* cand = coarse_list + (size_t)n_super * (rr.begin[r] - rr.begin[0]) + (size_t)sup * (rr.begin[r + 1] - rr.begin[r]);
            const int n_cand = coarse_count[r * n_super + sup];
            bool dirty = false;
            for (int c = 0; c < n_cand; ++c) {
                const int s = __ldg(cand + c);
                const int4 meta = __ldg(reinterpret_cast<const int4*>(&shapes[s].e0));
                const double pc = meta.w ? __ldg(psic + s) : 1.0;
                int lo = 0, hi = ncols;
                if (!(meta.w && (pc == 0.0 || pc != pc))) {           // (psi(centre) = 0 / NaN reaches every cell: never skipped)
                    for (int ed = meta.x; ed < meta.y && lo < hi; ++ed) {
                        const UamEdge rcd = uam_load_edge(edges + ed);
                        const unsigned t = __ballot_sync(0xffffffffu, uam_row_pred<1>(rcd, xs, y, e));
                        if (t) {
                            const int f = __ffs(t) - 1, g = 31 - __clz(t);
                            lo = max(lo, __shfl_sync(0xffffffffu, sj_prev, f) + 1);
                            hi = min(hi, __shfl_sync(0xffffffffu, sj_next, g));
                        } else if ((int)rcd.kind == UAM_EDGE_ELLIPSE) {
                            // no sample inside: the ellipse's cells on this row, if any, include the column nearest its centre
                            const double uc = (rcd.p0 - x0) / dx - 0.5 - (double)j0;
                            int jc = (uc >= 0.0 && uc < (double)ncols) ? (int)uc : (uc < 0.0 ? 0 : ncols - 1);
                            const int jt = min(max(jc - 1 + min(lane, 3), 0), ncols - 1);     // lanes 0..3: jc - 1 .. jc + 2
                            const bool in = uam_row_pred<1>(rcd, uam_cell_centre(j0 + jt, x0, dx), y, e);
                            if (__ballot_sync(0xffffffffu, in)) { lo = max(lo, jc - 10); hi = min(hi, jc + 12); }
                            else hi = lo;
                        } else {
                            hi = lo;
                        }
                    }
                }
                if (lo >= hi) continue;
                if (!dirty) {
#pragma unroll
                    for (int k = 0; k < UAM_SUPER / 32; ++k) acc[lane + 32 * k] = 0.0;
                    dirty = true;
                }
                for (int j = (lo & ~31) + lane; j < hi; j += 32) {
                    if (j < lo) continue;
                    const double x = uam_cell_centre(j0 + j, x0, dx);
                    double psi = 1.0;
                    for (int ed = meta.x; ed < meta.y; ++ed) {
                        const UamEdge rcd = uam_load_edge(edges + ed);
                        const double m = fmin(__dsub_rn(uam_h_exact(rcd, x, y), e), 0.0);
                        psi = __dmul_rn(psi, __dmul_rn(m, m));
                    }
                    if (meta.w) {
                        if (psi != 0.0 || pc == 0.0 || pc != pc) acc[j] = __dadd_rn(acc[j], __ddiv_rn(psi, pc));
                    } else {
                        acc[j] = __dadd_rn(acc[j], psi);
                    }
                }
            }
            float* out = layers + (size_t)r * plane + (size_t)i * W + j0;
#pragma unroll
            for (int k = 0; k < UAM_SUPER / 32; ++k) {
                const int j = lane + 32 * k;
                if (j < ncols) __stcs(out + j, dirty ? (float)acc[j] : 0.0f);
            }
        }
    }
}

__global__ void __launch_bounds__(256)
uam_k_layers_rows(const UamEdge* __restrict__ edges, const UamShape* __restrict__ shapes, const double* __restrict__ psic,
                  UamRegionRanges2 rr, int n_regions, int H, int W, double x0, double dx, double y0, double dy, double e,
                  const int* __restrict__ coarse_list, const int* __restrict__ coarse_count, int n_super,
                  float* __restrict__ layers) {
    __shared__ double acc_all[8][UAM_SUPER];
    uam_layers_rows_sampled(acc_all, edges, shapes, psic, rr, n_regions, H, W, x0, dx, y0, dy, e, coarse_list, coarse_count, n_super, layers,
                            blockIdx.y * gridDim.x + blockIdx.x, blockIdx.y * UAM_SUPER, UAM_SUPER);
}

// Cells [lo, hi) of one raster row against a shape of NE straight edges (records at ed[0 .. NE)): the row constants of every
// edge stay in registers, no record loads and no dispatch on the record kind per cell.  Same operations in the same order as
// uam_h_exact + the psi product of the generic loop (the minimum min(v, 0) is taken as v < 0 ? v : 0, which differs from
// fmin only in the sign of a zero that is squared next).  Returns false -- nothing done -- if an edge is not a straight line.
template <int NE>
__device__ __forceinline__ bool uam_layers_cells_lines(const UamEdge* __restrict__ ed, double* __restrict__ acc, int lane, int lo,
                                                       int hi, int j0, double x0, double dx, double y, double e, bool normalised,
                                                       double pc, bool always) {
    double A0[NE], A3[NE], T2[NE], NP4[NE];
    bool lines = true;
#pragma unroll
    for (int k = 0; k < NE; ++k) {
        const UamEdge rcd = uam_load_edge(ed + k);
        lines = lines && (int)rcd.kind == UAM_EDGE_LINE;
        A0[k] = rcd.p0; A3[k] = rcd.p3; T2[k] = __dmul_rn(rcd.p2, __dsub_rn(y, rcd.p1)); NP4[k] = -rcd.p4;
    }
    if (!lines) return false;
    // whole double trips first (two cells per lane, j and j + 32: two independent dependency chains through the fp64 pipe), then
    // the remaining < 64 cells one per lane and trip
    int j = lo + lane;
    for (const int hi2 = lo + ((hi - lo) & ~63); j < hi2; j += 64) {
        const double xa = uam_cell_centre(j0 + j, x0, dx), xb = uam_cell_centre(j0 + j + 32, x0, dx);
        double psa = 1.0, psb = 1.0;
#pragma unroll
        for (int k = 0; k < NE; ++k) {
            // uam_h_exact, line: -sgn * ((By-Ay)*(x-Ax) - (Bx-Ax)*(y-Ay)), the second product hoisted out of the row
            const double va = __dsub_rn(__dmul_rn(NP4[k], __dsub_rn(__dmul_rn(A3[k], __dsub_rn(xa, A0[k])), T2[k])), e);
            const double vb = __dsub_rn(__dmul_rn(NP4[k], __dsub_rn(__dmul_rn(A3[k], __dsub_rn(xb, A0[k])), T2[k])), e);
            const double ma = va < 0.0 ? va : 0.0, mb = vb < 0.0 ? vb : 0.0;
            psa = __dmul_rn(psa, __dmul_rn(ma, ma));
            psb = __dmul_rn(psb, __dmul_rn(mb, mb));
        }
        if (normalised) {
            if (psa != 0.0 || always) acc[j] = __dadd_rn(acc[j], __ddiv_rn(psa, pc));
            if (psb != 0.0 || always) acc[j + 32] = __dadd_rn(acc[j + 32], __ddiv_rn(psb, pc));
        } else {
            acc[j] = __dadd_rn(acc[j], psa);
            acc[j + 32] = __dadd_rn(acc[j + 32], psb);
        }
    }
    for (; j < hi; j += 32) {
        const double x = uam_cell_centre(j0 + j, x0, dx);
        double psi = 1.0;
#pragma unroll
        for (int k = 0; k < NE; ++k) {
            const double v = __dsub_rn(__dmul_rn(NP4[k], __dsub_rn(__dmul_rn(A3[k], __dsub_rn(x, A0[k])), T2[k])), e);
            const double m = v < 0.0 ? v : 0.0;
            psi = __dmul_rn(psi, __dmul_rn(m, m));
        }
        if (normalised) {
            if (psi != 0.0 || always) acc[j] = __dadd_rn(acc[j], __ddiv_rn(psi, pc));
        } else {
            acc[j] = __dadd_rn(acc[j], psi);
        }
    }
    return true;
}

// Interval form of the layer rasteriser (the default).  What the sampled row form above still pays per (row, candidate
// shape) -- one predicate evaluation per lane and inequality, a vote and two shuffles, whether the shape touches the row
// or not (ncu r02: 2.8 ms at 16384^2, issue-bound, the fp64 pipe a quarter busy) -- is done here the way the occupancy
// kernel does it: phase A, THREAD per row of the supertile: for every candidate shape of every region the thread finds the
// row's exact interval of cells with all h_i - e < 0 by bisection with the exact predicate (uam_row_interval_edge<1>, ~9
// evaluations per inequality, 32 rows per warp instruction) and leaves it in shared memory (2 bytes per (row, candidate), up to UAM_IV_CAP x 256 / ROWS candidates);
// phase B, WARP per row: only the cells of the interval are evaluated (they are exactly the cells where psi != 0), lanes =
// columns, same per-cell arithmetic and shape order as every other form, so the same bits.  Shapes of up to 4 straight edges
// (rectangular footprints, triangles) keep their row constants {Ax, By - Ay, (Bx - Ax)(y - Ay), -sgn} in registers: no
// record loads and no dispatch on the record kind per cell.  A supertile with more than UAM_IV_CAP candidates over all
// regions (their intervals would not fit) runs the sampled form.
#define UAM_IV_CAP 60
#define UAM_IV_EMPTY 1u          // lo | (hi - 1) << 8 with hi - 1 < lo
template <int MINB, int ROWS>          // ROWS rows of a supertile per CTA (256 or 128), one thread per row in phase A
__global__ void __launch_bounds__(ROWS, MINB)
uam_k_layers_iv(const UamEdge* __restrict__ edges, const UamShape* __restrict__ shapes, const double* __restrict__ psic,
                const __grid_constant__ UamRegionRanges2 rr, int n_regions, int H, int W, double x0, double dx, double y0, double dy, double e,
                const int* __restrict__ coarse_list, const int* __restrict__ coarse_count, int n_super,
                float* __restrict__ layers) {
    __shared__ double acc_all[ROWS / 32][UAM_SUPER];
    constexpr int CAP = UAM_IV_CAP * (UAM_SUPER / ROWS);               // 60 candidates with whole supertiles, 120 with halves
    __shared__ unsigned short iv[CAP][ROWS];
    __shared__ int s_ncand[UAM_MAX_REGIONS];
    __shared__ const int* s_cand[UAM_MAX_REGIONS];
    const int i0 = blockIdx.y * ROWS;
    const int sup = (i0 / UAM_SUPER) * gridDim.x + blockIdx.x;
    int n_all = 0;
    for (int r = 0; r < n_regions; ++r) n_all += coarse_count[r * n_super + sup];
    if (threadIdx.x < n_regions) {
        const int r = threadIdx.x;
        s_ncand[r] = coarse_count[r * n_super + sup];
        s_cand[r] = coarse_list + (size_t)n_super * (rr.begin[r] - rr.begin[0]) + (size_t)sup * (rr.begin[r + 1] - rr.begin[r]);
    }
    if (n_all > CAP) {                 // (CTA-uniform)
        uam_layers_rows_sampled(acc_all, edges, shapes, psic, rr, n_regions, H, W, x0, dx, y0, dy, e, coarse_list, coarse_count, n_super, layers,
                                sup, i0, ROWS);
        return;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j0 = blockIdx.x * UAM_SUPER;
    const int ncols = min(UAM_SUPER, W - j0);
    const size_t plane = (size_t)H * W;
    // phase A: thread = row
    {
        const int i = i0 + threadIdx.x;
        const double y = uam_cell_centre(i, y0, dy);
        int base = 0;
        for (int r = 0; r < n_regions; ++r) {
            const int* cand = coarse_list + (size_t)n_super * (rr.begin[r] - rr.begin[0]) + (size_t)sup * (rr.begin[r + 1] - rr.begin[r]);
            const int n_cand = coarse_count[r * n_super + sup];
            for (int c = 0; c < n_cand; ++c) {
                const int s = __ldg(cand + c);
                const int4 meta = __ldg(reinterpret_cast<const int4*>(&shapes[s].e0));
                int lo = 0, hi = i < H ? ncols : 0;
                bool every = false;                                   // psi(centre) = 0 / NaN reaches every cell: never skipped
                if (meta.w) {
                    const double pc = __ldg(psic + s);
                    every = pc == 0.0 || pc != pc;
                }
                if (!every) {
                    for (int ed = meta.x; ed < meta.y && lo < hi; ++ed) {
                        const UamEdge rcd = uam_load_edge(edges + ed);
                        uam_row_interval_edge<1>(rcd, y, x0, dx, j0, ncols, e, lo, hi);
                    }
                }
                iv[base + c][threadIdx.x] = (unsigned short)(lo < hi ? (unsigned)lo | ((unsigned)(hi - 1) << 8) : UAM_IV_EMPTY);
            }
            base += n_cand;
        }
    }
    __syncthreads();
    // phase B: warp = row
    double* acc = acc_all[warp];
    const bool vec = (W & 3) == 0 && (ncols & 3) == 0 && ((((uintptr_t)layers) & 15) == 0);
    for (int row = warp; row < ROWS; row += ROWS / 32) {
        const int i = i0 + row;
        if (i >= H) break;
        const double y = uam_cell_centre(i, y0, dy);
        float* out = layers + (size_t)i * W + j0;
        int base = 0;
        for (int r = 0; r < n_regions; ++r, out += plane) {
            const int* cand = s_cand[r];
            const int n_cand = s_ncand[r];
            bool dirty = false;
            for (int c = 0; c < n_cand; ++c) {
                const unsigned v = iv[base + c][row];
                const int lo = (int)(v & 255u), hi = (int)(v >> 8) + 1;
                if (hi <= lo) continue;
                if (!dirty) {
#pragma unroll
                    for (int k = 0; k < UAM_SUPER / 32; ++k) acc[lane + 32 * k] = 0.0;
                    dirty = true;
                }
                __syncwarp();                  // (the lane that owns a column changes from shape to shape: lane = (j - lo) mod 32)
                const int s = __ldg(cand + c);
                const int4 meta = __ldg(reinterpret_cast<const int4*>(&shapes[s].e0));
                const double pc = meta.w ? __ldg(psic + s) : 1.0;
                const bool always = pc == 0.0 || pc != pc;          // (psi / pc is added even when psi == 0: 0/0 = NaN)
                const int ne = meta.y - meta.x;
                if (ne == 4) {
                    if (uam_layers_cells_lines<4>(edges + meta.x, acc, lane, lo, hi, j0, x0, dx, y, e, meta.w != 0, pc, always)) continue;
                } else if (ne == 3) {
                    if (uam_layers_cells_lines<3>(edges + meta.x, acc, lane, lo, hi, j0, x0, dx, y, e, meta.w != 0, pc, always)) continue;
                }
                for (int j = lo + lane; j < hi; j += 32) {
                    const double x = uam_cell_centre(j0 + j, x0, dx);
                    double psi = 1.0;
                    for (int ed = meta.x; ed < meta.y; ++ed) {
                        const UamEdge rcd = uam_load_edge(edges + ed);
                        const double m = fmin(__dsub_rn(uam_h_exact(rcd, x, y), e), 0.0);
                        psi = __dmul_rn(psi, __dmul_rn(m, m));
                    }
                    if (meta.w) {
                        if (psi != 0.0 || always) acc[j] = __dadd_rn(acc[j], __ddiv_rn(psi, pc));
                    } else {
                        acc[j] = __dadd_rn(acc[j], psi);
                    }
                }
            }
            base += n_cand;
            if (!dirty) {                      // (warp-uniform) no shape of the region on this row
                if (vec) {
#pragma unroll
                    for (int k = 0; k < UAM_SUPER / 128; ++k)
                        if (4 * (lane + 32 * k) < ncols) __stcs(reinterpret_cast<float4*>(out) + lane + 32 * k, make_float4(0.0f, 0.0f, 0.0f, 0.0f));
                } else {
#pragma unroll
                    for (int k = 0; k < UAM_SUPER / 32; ++k)
                        if (lane + 32 * k < ncols) __stcs(out + lane + 32 * k, 0.0f);
                }
                continue;
            }
            __syncwarp();
            if (vec) {
#pragma unroll
                for (int k = 0; k < UAM_SUPER / 128; ++k) {
                    const int j = 4 * (lane + 32 * k);
                    if (j < ncols) {
                        const double2 a = *reinterpret_cast<const double2*>(acc + j), b2 = *reinterpret_cast<const double2*>(acc + j + 2);
                        __stcs(reinterpret_cast<float4*>(out) + lane + 32 * k, make_float4((float)a.x, (float)a.y, (float)b2.x, (float)b2.y));
                    }
                }
            } else {
#pragma unroll
                for (int k = 0; k < UAM_SUPER / 32; ++k) {
                    const int j = lane + 32 * k;
                    if (j < ncols) __stcs(out + j, (float)acc[j]);
                }
            }
            __syncwarp();
        }
    }
}

// -------------------------------------------------------------------------------------------------------
// exact EDT (squared Euclidean distance to the nearest occupied cell), two separable phases:
//   phase 1  g(i,j) = distance along column j to the nearest occupied cell.  Banded: per (256-row band, column) find
//            the first/last occupied row, a per-column scan over the band summaries gives every band the nearest
//            occupied row above and below it, a second banded sweep writes g.  (band x column) threads instead of
//            one thread per column.
//   phase 2  d2(i,u) = min_v (u-v)^2 + g(i,v)^2.
//            fast path: one thread per cell searches outwards, delta = 1, 2, ... until delta^2 >= best (no farther
//            column can win) -- O(distance) per cell, the row staged in shared memory.  Exact whenever it stops
//            within UAM_EDT_R columns; otherwise the row is flagged.
//            slow path, flagged rows only (sparse maps): transpose -> per-row lower envelope (Meijster) with an
//            explicit stack in HBM -> transpose back.  All its kernels return at once when no row is flagged.
// Integer arithmetic throughout: bit-exact against scipy's EDT.
// -------------------------------------------------------------------------------------------------------
#define UAM_GINF (1 << 20)   // "no occupied cell in this column" (> any real distance; its square fits int64)
#define UAM_EDT_BAND 256
#define UAM_EDT_R 1024
#define UAM_EDT_CLIP 32768   // fast path: g clipped here (its square is 2^30 = the "none" value of d2)

__global__ void __launch_bounds__(128)
uam_k_edt_band_summary(const uint8_t* __restrict__ occ, int H, int W, int* __restrict__ band_first, int* __restrict__ band_last) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int band = blockIdx.y;
    if (j >= W) return;
    const int i0 = band * UAM_EDT_BAND, i1 = min(i0 + UAM_EDT_BAND, H);
    int first = -1, last = -1;
#pragma unroll 8
    for (int i = i0; i < i1; ++i) {
        if (occ[(size_t)i * W + j]) {
            if (first < 0) first = i;
            last = i;
        }
    }
    band_first[(size_t)band * W + j] = first;
    band_last[(size_t)band * W + j] = last;
}

// per column: nearest occupied row strictly above the band (above[band]) and below it (below[band]); -1 = none
__global__ void __launch_bounds__(128)
uam_k_edt_band_carry(int n_bands, int W, const int* __restrict__ band_first, const int* __restrict__ band_last,
                     int* __restrict__ above, int* __restrict__ below) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= W) return;
    int run = -1;
    for (int b = 0; b < n_bands; ++b) {
        above[(size_t)b * W + j] = run;
        const int l = band_last[(size_t)b * W + j];
        if (l >= 0) run = l;
    }
    run = -1;
    for (int b = n_bands - 1; b >= 0; --b) {
        below[(size_t)b * W + j] = run;
        const int f = band_first[(size_t)b * W + j];
        if (f >= 0) run = f;
    }
}

// g is stored as uint16, clipped at UAM_EDT_CLIP = 2^15 ("no occupied cell in this column": a real distance is below
// 23170, the largest raster side): half the bytes of phase 1's output and of phase 2's input.  COLS = 4: a thread sweeps four
// adjacent columns (4-byte loads of the occupancy, 8-byte stores of g; needs W % 4 == 0), COLS = 1 otherwise.
template <int COLS>
__global__ void __launch_bounds__(128)
uam_k_edt_band_sweep(const uint8_t* __restrict__ occ, int H, int W, const int* __restrict__ above,
                     const int* __restrict__ below, unsigned short* __restrict__ g) {
    const int j = (blockIdx.x * blockDim.x + threadIdx.x) * COLS;
    const int band = blockIdx.y;
    if (j >= W) return;
    const int i0 = band * UAM_EDT_BAND, i1 = min(i0 + UAM_EDT_BAND, H);
    int last[COLS], next[COLS];
#pragma unroll
    for (int c = 0; c < COLS; ++c) last[c] = above[(size_t)band * W + j + c];
#pragma unroll 4
    for (int i = i0; i < i1; ++i) {
        unsigned o4;
        if (COLS == 4) o4 = *reinterpret_cast<const unsigned*>(occ + (size_t)i * W + j);
        else o4 = occ[(size_t)i * W + j];
        unsigned short out[COLS];
#pragma unroll
        for (int c = 0; c < COLS; ++c) {
            if ((o4 >> (8 * c)) & 0xffu) last[c] = i;
            out[c] = (unsigned short)(last[c] >= 0 ? min(i - last[c], UAM_EDT_CLIP) : UAM_EDT_CLIP);
        }
        if (COLS == 4) *reinterpret_cast<uint2*>(g + (size_t)i * W + j) = make_uint2(out[0] | ((unsigned)out[1] << 16), out[2] | ((unsigned)out[3 % COLS] << 16));
        else g[(size_t)i * W + j] = out[0];
    }
#pragma unroll
    for (int c = 0; c < COLS; ++c) next[c] = below[(size_t)band * W + j + c];
#pragma unroll 4
    for (int i = i1 - 1; i >= i0; --i) {
        unsigned short cur[COLS];
        if (COLS == 4) {
            const uint2 v = *reinterpret_cast<const uint2*>(g + (size_t)i * W + j);
            cur[0] = (unsigned short)(v.x & 0xffffu); cur[1 % COLS] = (unsigned short)(v.x >> 16);
            cur[2 % COLS] = (unsigned short)(v.y & 0xffffu); cur[3 % COLS] = (unsigned short)(v.y >> 16);
        } else {
            cur[0] = g[(size_t)i * W + j];
        }
        bool any = false;
#pragma unroll
        for (int c = 0; c < COLS; ++c) {
            if (cur[c] == 0) next[c] = i;
            const int d = next[c] >= 0 ? min(next[c] - i, UAM_EDT_CLIP) : UAM_EDT_CLIP;
            if (d < (int)cur[c]) { cur[c] = (unsigned short)d; any = true; }
        }
        if (any) {
            if (COLS == 4) *reinterpret_cast<uint2*>(g + (size_t)i * W + j) = make_uint2(cur[0] | ((unsigned)cur[1 % COLS] << 16), cur[2 % COLS] | ((unsigned)cur[3 % COLS] << 16));
            else g[(size_t)i * W + j] = cur[0];
        }
    }
}

__device__ __forceinline__ float uam_clearance_of(int d2, float cellf) { return __fsqrt_rn((float)d2) * cellf; }

// phase 2 fast path: block = UAM_EDT_SPAN consecutive cells of one row (4 per thread); the row segment and UAM_EDT_R
// columns on either side are staged in shared memory together with the minimum of g over every aligned group of 8 columns.
// A cell scans its own group, then walks outwards group by group: a group at column distance D whose minimum is m cannot
// hold a better column when D^2 + m^2 >= best, and no farther group can when D^2 >= best.  What makes the skipping bite is
// a tight `best` from the start: the block first solves one ANCHOR cell per group exactly (1/8 of the cells, starting from
// the loose bound g(c)^2), then every other cell starts from its group's anchor through the Lipschitz bound
// d(u) <= d(anchor) + |u - anchor| -- so almost every group on its way is dismissed by one comparison and only the groups
// around the true nearest column are scanned.  Exact (a bound only prunes columns that cannot win).  ncu r02: the
// column-by-column search was issue-bound, 4.5 ms at 16384^2.
#define UAM_EDT_SPAN 4096
#define UAM_EDT_WIN (UAM_EDT_SPAN + 2 * UAM_EDT_R)
// one group of 8 columns against cell c.  sq holds g^2 + k^2 for column 8 q + k (2^30 + k^2 = "no occupied cell in this
// column"): with b = 8 q - c the candidate (b + k)^2 + g^2 is b^2 + (2 b) k + sq -- one multiply-add with a constant k and
// half a three-way minimum per column, b^2 added once at the end
__device__ __forceinline__ void uam_edt_scan_group(const int* sq, int q, int c, int& best) {
    const int4 a = *reinterpret_cast<const int4*>(&sq[q * 8]);
    const int4 b4 = *reinterpret_cast<const int4*>(&sq[q * 8 + 4]);
    const int b = q * 8 - c;                       // column offset of the group's first cell
    const int t = 2 * b;
    int m = min(a.x, t + a.y);
    m = min(min(m, t * 2 + a.z), t * 3 + a.w);
    m = min(min(m, t * 4 + b4.x), t * 5 + b4.y);
    m = min(min(m, t * 6 + b4.z), t * 7 + b4.w);
    best = min(best, b * b + m);
}

// Staging the window and the divergence of per-lane searches are what the earlier versions of this kernel spent their time
// on (ncu r02: 4.7 G warp instructions at 16384^2, issue-bound; counted on config C4: the mean distance is 42 cells, yet a
// warp needs only ~6 group steps and ~3 scans once the search is organised as below -- what is left is fixed cost per warp).
// The span is 4096 cells (halo overhead 1.5 x instead of 3 x); a thread stages whole groups of 8 columns with one 16-byte
// load, writes their squares, the group minimum and the minimum of every 64 columns.  The SEARCH IS WARP-UNIFORM: the 32
// consecutive cells of a warp look at the same group at the same time; a group is scanned by every lane (two broadcast
// 16-byte loads) as soon as one lane can gain from it, skipped with one vote otherwise, and a whole 64-column block is
// skipped when no lane can gain from its minimum.  Each lane starts from its own column, the lanes then exchange their
// radii (r_l = min(r_l, r_k + |l - k|): the Lipschitz bound d(u) <= d(u') + |u - u'|), so one lane that sees a near obstacle
// tightens all 32.  Exact: a bound only prunes columns that cannot win.
#define UAM_EDT_BLK 64
#define UAM_EDT_NONE (1 << 30)
__global__ void __launch_bounds__(256)
uam_k_edt_rows_fast(const unsigned short* __restrict__ g, int H, int W, int* __restrict__ d2, float* __restrict__ clearance,
                    float cellf, uint8_t* __restrict__ row_flag, int* __restrict__ any_flag) {
    __shared__ __align__(16) int sq[UAM_EDT_WIN];                      // g^2
    __shared__ __align__(16) int sm[UAM_EDT_WIN / 8];                                // min of g^2 per group of 8 columns
    __shared__ int sm64[UAM_EDT_WIN / UAM_EDT_BLK];                    // ... per block of 64
    __shared__ int s_next;                                             // next 32-cell chunk of the span
    if (threadIdx.x == 0) s_next = 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.y;
    const int u0 = blockIdx.x * UAM_EDT_SPAN;
    const unsigned short* grow = g + (size_t)i * W;
    const bool vec = (W & 7) == 0 && ((((uintptr_t)g) & 15) == 0);
    for (int q = threadIdx.x; q < UAM_EDT_WIN / 8; q += 256) {
        const int col = u0 - UAM_EDT_R + q * 8;
        uint4 v;
        if (vec && col >= 0 && col + 8 <= W) {
            v = __ldg(reinterpret_cast<const uint4*>(grow + col));
        } else {
            unsigned w[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int c0 = col + 2 * k, c1 = c0 + 1;
                const unsigned a0 = (c0 >= 0 && c0 < W) ? grow[c0] : (unsigned)UAM_EDT_CLIP;
                const unsigned a1 = (c1 >= 0 && c1 < W) ? grow[c1] : (unsigned)UAM_EDT_CLIP;
                w[k] = a0 | (a1 << 16);
            }
            v = make_uint4(w[0], w[1], w[2], w[3]);
        }
        const unsigned w[4] = {v.x, v.y, v.z, v.w};
        int s8[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int g0 = (int)(w[k] & 0xffffu), g1 = (int)(w[k] >> 16);
            s8[2 * k] = g0 * g0;
            s8[2 * k + 1] = g1 * g1;
        }
        *reinterpret_cast<int4*>(&sq[q * 8]) = make_int4(s8[0], s8[1] + 1, s8[2] + 4, s8[3] + 9);
        *reinterpret_cast<int4*>(&sq[q * 8 + 4]) = make_int4(s8[4] + 16, s8[5] + 25, s8[6] + 36, s8[7] + 49);
        int m = min(min(min(s8[0], s8[1]), min(s8[2], s8[3])), min(min(s8[4], s8[5]), min(s8[6], s8[7])));
        sm[q] = m;
        // 8 consecutive groups = one 64-column block: consecutive lanes hold them (q = threadIdx.x + 256 k, 8 | 256)
        m = min(m, __shfl_xor_sync(0xffffffffu, m, 1));
        m = min(m, __shfl_xor_sync(0xffffffffu, m, 2));
        m = min(m, __shfl_xor_sync(0xffffffffu, m, 4));
        if ((lane & 7) == 0) sm64[q >> 3] = m;
    }
    __syncthreads();
    bool unresolved = false;
    int* d2row = d2 + (size_t)i * W + u0;
    float* clrow = clearance ? clearance + (size_t)i * W + u0 : nullptr;
    const int ncell = min(UAM_EDT_SPAN, W - u0);              // cells of this span inside the raster
    // the 32-cell chunks of the span are handed out through a shared-memory counter: the warps inside obstacles are done at once
#pragma unroll 1
    for (;;) {
        int kk = 0;
        if (lane == 0) kk = atomicAdd(&s_next, 1);
        const int off0 = __shfl_sync(0xffffffffu, kk, 0) * 32;     // the warp's first cell (a multiple of 32)
        if (off0 >= ncell) break;                             // (warp-uniform)
        const int c = UAM_EDT_R + off0 + lane;
        const int qa = (UAM_EDT_R + off0) >> 3;               // the warp's four own groups: qa .. qa + 3 (qa is a multiple of 4)
        const bool live = off0 + lane < ncell;
        // start: the own column; radii exchanged between the lanes (r >= the true distance, r + |l - k| bounds the neighbour's)
        int best = sq[c] - (lane & 7) * (lane & 7);           // g^2 of the own column (c = 8 q + (lane & 7))
        if (__all_sync(0xffffffffu, !live || best == 0)) {    // a warp inside an obstacle: nothing to search
            if (live) {
                d2row[off0 + lane] = 0;
                if (clrow) clrow[off0 + lane] = 0.0f;
            }
            continue;
        }
        int r = (int)__fsqrt_ru((float)best) + 1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) r = min(r, __shfl_xor_sync(0xffffffffu, r, o) + o);
        best = live ? min(best, r * r) + 1 : 0;               // strictly above the minimum; a lane past the row's end wants nothing
        // the warp's own 32 columns: groups qa .. qa + 3
        {
            const int4 m = *reinterpret_cast<const int4*>(&sm[qa]);
            const int mm[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int D = max(max(8 * t - lane, lane - (8 * t + 7)), 0);
                if (__any_sync(0xffffffffu, D * D + mm[t] < best)) uam_edt_scan_group(sq, qa + t, c, best);
            }
        }
        // outwards, one BLOCK of 32 columns (4 groups) on either side per step: left block t = groups qa - 4t .. qa - 4t + 3,
        // right block t = groups qa + 4t .. qa + 4t + 3; both stay inside the window for t <= UAM_EDT_R / 32.  A step is one
        // termination vote, two 16-byte loads of group minima and ONE vote whether any lane can gain from any of the 8 groups;
        // only then are the groups looked at one by one, nearest first (ncu r02: the walk group by group spent 35 instructions per
        // 8 columns and side on votes and addresses, 63 % of the kernel).
        bool done = false;
#pragma unroll 1
        for (int t = 1; t <= UAM_EDT_R / 32; ++t) {
            const int DL = lane + 1 + 32 * (t - 1), DR = 32 * t - lane;       // distance to the nearest column of the left / right block
            const int Dm = min(DL, DR);
            if (__all_sync(0xffffffffu, Dm * Dm >= best)) { done = true; break; }    // no farther column can win
            const int qL = qa - 4 * t, qR = qa + 4 * t;
            const int4 mL = *reinterpret_cast<const int4*>(&sm[qL]);
            const int4 mR = *reinterpret_cast<const int4*>(&sm[qR]);
            // group g of the left block is 8 (3 - g) columns farther than DL, group g of the right block 8 g farther than DR
            const int vL3 = DL * DL + mL.w, vL2 = (DL + 8) * (DL + 8) + mL.z, vL1 = (DL + 16) * (DL + 16) + mL.y, vL0 = (DL + 24) * (DL + 24) + mL.x;
            const int vR0 = DR * DR + mR.x, vR1 = (DR + 8) * (DR + 8) + mR.y, vR2 = (DR + 16) * (DR + 16) + mR.z, vR3 = (DR + 24) * (DR + 24) + mR.w;
            const int vmin = min(min(min(vL0, vL1), min(vL2, vL3)), min(min(vR0, vR1), min(vR2, vR3)));
            if (!__any_sync(0xffffffffu, vmin < best)) continue;
            if (__any_sync(0xffffffffu, vL3 < best)) uam_edt_scan_group(sq, qL + 3, c, best);
            if (__any_sync(0xffffffffu, vR0 < best)) uam_edt_scan_group(sq, qR, c, best);
            if (__any_sync(0xffffffffu, vL2 < best)) uam_edt_scan_group(sq, qL + 2, c, best);
            if (__any_sync(0xffffffffu, vR1 < best)) uam_edt_scan_group(sq, qR + 1, c, best);
            if (__any_sync(0xffffffffu, vL1 < best)) uam_edt_scan_group(sq, qL + 1, c, best);
            if (__any_sync(0xffffffffu, vR2 < best)) uam_edt_scan_group(sq, qR + 2, c, best);
            if (__any_sync(0xffffffffu, vL0 < best)) uam_edt_scan_group(sq, qL, c, best);
            if (__any_sync(0xffffffffu, vR3 < best)) uam_edt_scan_group(sq, qR + 3, c, best);
        }
        if (!done) {
            // the window is used up on both sides (UAM_EDT_R columns): unresolved if some lane could still gain from a column
            // beyond it and the raster goes on there
            const int DL = lane + 1 + UAM_EDT_R, DR = UAM_EDT_R + 32 - lane;
            const bool left_open = u0 + off0 - UAM_EDT_R > 0, right_open = u0 + off0 + 32 + UAM_EDT_R < W;
            if ((left_open && __any_sync(0xffffffffu, DL * DL < best)) || (right_open && __any_sync(0xffffffffu, DR * DR < best))) unresolved = true;
        }
        if (live) {
            d2row[off0 + lane] = best;
            if (clrow) clrow[off0 + lane] = uam_clearance_of(best, cellf);
        }
    }
    if (__syncthreads_or(unresolved) && threadIdx.x == 0) {
        row_flag[i] = 1;
        *any_flag = 1;
    }
}

// out[c][r] = in[r][c]   (in: R x C), slow path only: returns at once when no row is flagged.  IN = unsigned short reads
// the clipped g (UAM_EDT_CLIP -> UAM_GINF, "no occupied cell"); out_row_flag (optional) selects which rows of `out`
// (= columns c of `in`) are written.  Grid-stride over 32 x 32 tiles (a few CTAs cost nothing when there is nothing to do).
template <typename IN>
__global__ void __launch_bounds__(256)
uam_k_transpose_i32(const IN* __restrict__ in, int R, int C, int* __restrict__ out, const int* __restrict__ any_flag,
                    const uint8_t* __restrict__ out_row_flag) {
    if (any_flag && *any_flag == 0) return;
    __shared__ int tile[32][33];
    const int tiles_c = (C + 31) / 32, tiles_r = (R + 31) / 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (long long t = blockIdx.x; t < (long long)tiles_c * tiles_r; t += gridDim.x) {
        const int c0 = (int)(t % tiles_c) * 32, r0 = (int)(t / tiles_c) * 32;
        for (int k = ty; k < 32; k += 8) {
            const int r = r0 + k, c = c0 + tx;
            if (r < R && c < C) {
                int v = (int)in[(size_t)r * C + c];
                if (sizeof(IN) == 2 && v >= UAM_EDT_CLIP) v = UAM_GINF;
                tile[k][tx] = v;
            }
        }
        __syncthreads();
        for (int k = ty; k < 32; k += 8) {
            const int c = c0 + k, r = r0 + tx;
            if (r < R && c < C && (!out_row_flag || out_row_flag[c])) out[(size_t)c * R + r] = tile[tx][k];
        }
        __syncthreads();
    }
}

struct __align__(16) UamEdtEntry {
    int s, t;
    long long gsq;
};

// slow path: thread per FLAGGED row i; gT[u][i] = g(i, u).  Stack entry q of row i lives at stack[q * H + i].
__global__ void __launch_bounds__(128)
uam_k_edt_rows(const int* __restrict__ gT, int H, int W, UamEdtEntry* __restrict__ stack, int* __restrict__ dT,
               const int* __restrict__ any_flag, const uint8_t* __restrict__ row_flag) {
    if (*any_flag == 0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= H || !row_flag[i]) return;
    int q = 0;
    UamEdtEntry top;
    {
        const long long g0 = gT[i];
        top.s = 0; top.t = 0; top.gsq = g0 * g0;
        stack[i] = top;
    }
    for (int u = 1; u < W; ++u) {
        const long long gu = gT[(size_t)u * H + i];
        const long long gusq = gu * gu;
        // pop while the parabola of u is below the top's at the top's start t
        while (q >= 0) {
            const long long dt_s = (long long)(top.t - top.s), dt_u = (long long)(top.t - u);
            if (dt_s * dt_s + top.gsq > dt_u * dt_u + gusq) {
                --q;
                if (q >= 0) top = stack[(size_t)q * H + i];
            } else {
                break;
            }
        }
        if (q < 0) {
            q = 0;
            top.s = u; top.t = 0; top.gsq = gusq;
            stack[i] = top;
        } else {
            // Sep(s, u) = (u^2 - s^2 + g(u)^2 - g(s)^2) div (2 (u - s)), numerator >= 0 here
            const long long num = (long long)u * u - (long long)top.s * top.s + gusq - top.gsq;
            const long long w = 1 + num / (2ll * (u - top.s));
            if (w < W) {
                ++q;
                top.s = u; top.t = (int)w; top.gsq = gusq;
                stack[(size_t)q * H + i] = top;
            }
        }
    }
    for (int u = W - 1; u >= 0; --u) {
        const long long d = (long long)(u - top.s);
        const long long v = d * d + top.gsq;
        dT[(size_t)u * H + i] = v >= (1ll << 30) ? (1 << 30) : (int)v;
        if (u == top.t && q > 0) {
            --q;
            top = stack[(size_t)q * H + i];
        }
    }
}

// clearance of the rows the slow path rewrote (the fast path writes the clearance of its own rows itself)
__global__ void __launch_bounds__(256)
uam_k_edt_clearance_rows(const int* __restrict__ d2, int H, int W, float cellf, float* __restrict__ clearance,
                         const int* __restrict__ any_flag, const uint8_t* __restrict__ row_flag) {
    if (*any_flag == 0) return;
    for (int i = blockIdx.x; i < H; i += gridDim.x) {
        if (!row_flag[i]) continue;
        for (int u = threadIdx.x; u < W; u += blockDim.x) clearance[(size_t)i * W + u] = uam_clearance_of(d2[(size_t)i * W + u], cellf);
    }
}

}  // namespace

extern "C" int uam_dem_mask(uam_ctx* ctx, const float* d_image, int64_t n, float threshold, uint8_t* d_mask,
                            void* stream) {
    if (!ctx) return UAM_ERR_INVALID;
    UAM_NVTX("uam.map.dem_mask");
    if (n < 0 || (n > 0 && (!d_image || !d_mask))) return uam_fail(ctx, UAM_ERR_INVALID, "bad argument to uam_dem_mask");
    if (n == 0) return UAM_OK;
    UAM_CUDA(ctx, cudaSetDevice(ctx->device));
    const int eq_mode = threshold == -9999.0f;      // data_manager.py:14-15
    const long long ctas = std::min<long long>((n / 4 + 255) / 256 + 1, (long long)ctx->sm_count * 16);
    uam_k_dem_mask<<<(unsigned)ctas, 256, 0, uam_pick_stream(ctx, stream)>>>(d_image, n, threshold, eq_mode, d_mask);
    UAM_CHECK_LAUNCH(ctx, "uam_k_dem_mask");
    return UAM_OK;
}

// supertile lists of the shapes [s_begin, s_end) into out_list / out_count; through 2048-cell blocks first on large rasters
static int uam_cull_two_level(uam_ctx* ctx, int s_begin, int s_end, double thr, int H, int W, double x0, double dx, double y0,
                              double dy, int* out_list, int* out_count, cudaStream_t st, const double* psic = nullptr) {
    dim3 sgrid((W + UAM_SUPER - 1) / UAM_SUPER, (H + UAM_SUPER - 1) / UAM_SUPER);
    dim3 g0((W + UAM_SUPER0 - 1) / UAM_SUPER0, (H + UAM_SUPER0 - 1) / UAM_SUPER0);
    const int ns = s_end - s_begin;
    if (ns <= 0 || (size_t)g0.x * g0.y < 4) {
        uam_k_cull_coarse<<<sgrid, 256, 0, st>>>(ctx->d_edges, ctx->d_shapes, s_begin, s_end, thr, H, W, x0, dx, y0, dy, UAM_SUPER, nullptr,
                                                  nullptr, 1, 1, out_list, out_count, psic);
        UAM_CHECK_LAUNCH(ctx, "uam_k_cull_coarse");
        return UAM_OK;
    }
    const size_t n0 = (size_t)g0.x * g0.y;
    UAM_TRY(uam_reserve(ctx, &ctx->d_cull0_scratch, &ctx->cull0_scratch_bytes, (n0 * (size_t)ns + n0) * 4));
    int* list0 = (int*)ctx->d_cull0_scratch;
    int* count0 = list0 + n0 * (size_t)ns;
    uam_k_cull_coarse<<<g0, 256, 0, st>>>(ctx->d_edges, ctx->d_shapes, s_begin, s_end, thr, H, W, x0, dx, y0, dy, UAM_SUPER0, nullptr, nullptr, 1,
                                          1, list0, count0, psic);
    UAM_CHECK_LAUNCH(ctx, "uam_k_cull_coarse");
    uam_k_cull_coarse<<<sgrid, 256, 0, st>>>(ctx->d_edges, ctx->d_shapes, s_begin, s_end, thr, H, W, x0, dx, y0, dy, UAM_SUPER, list0, count0,
                                              UAM_SUPER0, (int)g0.x, out_list, out_count, psic);
    UAM_CHECK_LAUNCH(ctx, "uam_k_cull_coarse");
    return UAM_OK;
}

static int uam_check_grid(uam_ctx* ctx, int H, int W, double dx, double dy, const void* out) {
    if (!ctx) return UAM_ERR_INVALID;
    if (H < 1 || W < 1 || !out) return uam_fail(ctx, UAM_ERR_INVALID, "bad raster size / NULL output");
    if (!(dx != 0.0) || !(dy != 0.0)) return uam_fail(ctx, UAM_ERR_INVALID, "cell size must be non-zero");
    if (!ctx->has_shapes) return uam_fail(ctx, UAM_ERR_STATE, "no shape table: call uam_map_set_shapes first");
    return UAM_OK;
}

extern "C" int uam_rasterize_occupancy(uam_ctx* ctx, int H, int W, double x0, double dx, double y0, double dy,
                                       uint8_t* d_occ, void* stream) {
    UAM_TRY(uam_check_grid(ctx, H, W, dx, dy, d_occ));
    UAM_NVTX("uam.map.rasterize_occupancy");
    UAM_CUDA(ctx, cudaSetDevice(ctx->device));
    dim3 grid((W + UAM_TILE_W - 1) / UAM_TILE_W, (H + UAM_TILE_H - 1) / UAM_TILE_H);
    if (grid.y > 65535) return uam_fail(ctx, UAM_ERR_UNSUPPORTED, "H too large");
    cudaStream_t st = uam_pick_stream(ctx, stream);
    dim3 sgrid((W + UAM_SUPER - 1) / UAM_SUPER, (H + UAM_SUPER - 1) / UAM_SUPER);
    const size_t n_super = (size_t)sgrid.x * sgrid.y;
    UAM_TRY(uam_reserve(ctx, &ctx->d_cull_scratch, &ctx->cull_scratch_bytes, (n_super * (size_t)std::max(ctx->n_obs, 1) + n_super) * 4));
    int* clist = (int*)ctx->d_cull_scratch;
    int* ccount = clist + n_super * (size_t)std::max(ctx->n_obs, 1);
    UAM_TRY(uam_cull_two_level(ctx, 0, ctx->n_obs, 1e-14, H, W, x0, dx, y0, dy, clist, ccount, st));
    // scanline form (row intervals by bisection with the exact predicate) unless the records hold non-finite numbers or the
    // per-cell form is asked for (UAM_OPT_RASTERIZER = 0: the round-1 kernel, kept as the cross-check of the tests)
    if (ctx->edges_finite && ctx->rasterizer_scan && std::isfinite(x0) && std::isfinite(dx) && std::isfinite(y0) && std::isfinite(dy)) {
        uam_k_occupancy_scan<<<sgrid, 256, 0, st>>>(ctx->d_edges, ctx->d_shapes, ctx->n_obs, H, W, x0, dx, y0, dy, clist, ccount, d_occ);
        UAM_CHECK_LAUNCH(ctx, "uam_k_occupancy_scan");
        return UAM_OK;
    }
    uam_k_rasterize_occupancy<<<grid, 256, 0, st>>>(ctx->d_edges, ctx->d_shapes, ctx->n_obs, H, W, x0, dx, y0, dy, clist, ccount, d_occ);
    UAM_CHECK_LAUNCH(ctx, "uam_k_rasterize_occupancy");
    return UAM_OK;
}

extern "C" int uam_rasterize_layers(uam_ctx* ctx, int H, int W, double x0, double dx, double y0, double dy,
                                    double enlargement, float* d_layers, void* stream) {
    UAM_TRY(uam_check_grid(ctx, H, W, dx, dy, d_layers));
    UAM_NVTX("uam.map.rasterize_layers");
    if (ctx->n_regions < 1) return uam_fail(ctx, UAM_ERR_STATE, "the map has no regions");
    UAM_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = uam_pick_stream(ctx, stream);
    UamParams prm = {};
    prm.e = enlargement;
    prm.flags = UAM_PENALTY_SMOOTH | UAM_OBSTACLE_SMOOTH;
    UAM_TRY(uam_ensure_shape_norm(ctx, prm, st, false));
    UamRegionRanges2 rr;
    for (int r = 0; r <= ctx->n_regions; ++r) rr.begin[r] = ctx->region_begin[r];
    dim3 grid((W + UAM_TILE_W - 1) / UAM_TILE_W, (H + UAM_TILE_H - 1) / UAM_TILE_H);
    if (grid.y > 65535) return uam_fail(ctx, UAM_ERR_UNSUPPORTED, "H too large");
    dim3 sgrid((W + UAM_SUPER - 1) / UAM_SUPER, (H + UAM_SUPER - 1) / UAM_SUPER);
    const size_t n_super = (size_t)sgrid.x * sgrid.y;
    const int n_reg_shapes = rr.begin[ctx->n_regions] - rr.begin[0];
    UAM_TRY(uam_reserve(ctx, &ctx->d_cull_scratch, &ctx->cull_scratch_bytes,
                        (n_super * (size_t)std::max(n_reg_shapes, 1) + n_super * ctx->n_regions) * 4));
    int* clist = (int*)ctx->d_cull_scratch;
    int* ccount = clist + n_super * (size_t)std::max(n_reg_shapes, 1);
    for (int r = 0; r < ctx->n_regions; ++r) {
        UAM_TRY(uam_cull_two_level(ctx, rr.begin[r], rr.begin[r + 1], enlargement, H, W, x0, dx, y0, dy,
                                   clist + n_super * (size_t)(rr.begin[r] - rr.begin[0]), ccount + (size_t)r * n_super, st, ctx->d_psic));
    }
    // scanline form unless it could change a bit: non-finite numbers, or magnitudes for which a product of squared factors
    // could overflow before its zero factor (inf * 0 = NaN in the reference; a skipped cell would give 0)
    bool scan = ctx->edges_finite && ctx->rasterizer_scan && std::isfinite(x0) && std::isfinite(dx) && std::isfinite(y0) &&
                std::isfinite(dy) && std::isfinite(enlargement);
    if (scan) {
        const double ext = std::max(std::max(std::fabs(x0), std::fabs(x0 + dx * W)), std::max(std::fabs(y0), std::fabs(y0 + dy * H)));
        // |h| <= 4 (M + ext)^3 for every record kind with all numbers below M; a shape has at most max_edges factors m^2
        const double hb = 4.0 * std::pow(ctx->edges_max_abs + ext + std::fabs(enlargement) + 1.0, 3.0);
        scan = 2.0 * ctx->max_edges_per_shape * std::log10(hb) < 300.0;
    }
    if (scan && ctx->rasterizer_scan == 2) {         // the tile form with row intervals by bisection (slower: kept for comparison)
        uam_k_layers_scan<<<grid, 256, 0, st>>>(ctx->d_edges, ctx->d_shapes, ctx->d_psic, rr, ctx->n_regions, H, W, x0, dx,
                                                y0, dy, enlargement, clist, ccount, (int)n_super, d_layers);
        UAM_CHECK_LAUNCH(ctx, "uam_k_layers_scan");
        return UAM_OK;
    }
    if (scan && ctx->rasterizer_scan == 3) {         // the sampled row form (round 2's first scanline kernel: kept for comparison)
        uam_k_layers_rows<<<sgrid, 256, 0, st>>>(ctx->d_edges, ctx->d_shapes, ctx->d_psic, rr, ctx->n_regions, H, W, x0, dx,
                                                 y0, dy, enlargement, clist, ccount, (int)n_super, d_layers);
        UAM_CHECK_LAUNCH(ctx, "uam_k_layers_rows");
        return UAM_OK;
    }
    if (scan) {
        // half supertiles (128 rows, 128 threads) per CTA: twice the CTAs of half the length -- the kernel ends with the last heavy
        // supertile, ncu r02: the SMs were busy 90 % of its time with whole supertiles.  (UAM_LAYERS_ROWS = 256: A/B runs only)
        static const int rows = getenv("UAM_LAYERS_ROWS") ? atoi(getenv("UAM_LAYERS_ROWS")) : 128;
        if (rows == 256) {
            uam_k_layers_iv<2, 256><<<sgrid, 256, 0, st>>>(ctx->d_edges, ctx->d_shapes, ctx->d_psic, rr, ctx->n_regions, H, W, x0, dx,
                                                           y0, dy, enlargement, clist, ccount, (int)n_super, d_layers);
        } else {
            dim3 hgrid(sgrid.x, (H + 127) / 128);
            uam_k_layers_iv<4, 128><<<hgrid, 128, 0, st>>>(ctx->d_edges, ctx->d_shapes, ctx->d_psic, rr, ctx->n_regions, H, W, x0, dx,
                                                           y0, dy, enlargement, clist, ccount, (int)n_super, d_layers);
        }
        UAM_CHECK_LAUNCH(ctx, "uam_k_layers_iv");
        return UAM_OK;
    }
    uam_k_rasterize_layers<<<grid, 256, 0, st>>>(ctx->d_edges, ctx->d_shapes, ctx->d_psic, rr, ctx->n_regions, H, W, x0, dx,
                                                 y0, dy, enlargement, clist, ccount, (int)n_super, d_layers);
    UAM_CHECK_LAUNCH(ctx, "uam_k_rasterize_layers");
    return UAM_OK;
}

extern "C" int uam_edt(uam_ctx* ctx, const uint8_t* d_occ, int H, int W, double cell, int32_t* d_dist2,
                       float* d_clearance, void* stream) {
    if (!ctx) return UAM_ERR_INVALID;
    UAM_NVTX("uam.map.edt");
    if (H < 1 || W < 1 || !d_occ || (!d_dist2 && !d_clearance)) return uam_fail(ctx, UAM_ERR_INVALID, "bad argument to uam_edt");
    if (H > 23170 || W > 23170) return uam_fail(ctx, UAM_ERR_UNSUPPORTED, "raster larger than 23170 per side (d^2 must stay below 2^30)");
    UAM_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = uam_pick_stream(ctx, stream);
    const size_t n = (size_t)H * W;
    const int n_bands = (H + UAM_EDT_BAND - 1) / UAM_EDT_BAND;
    const size_t bw = (size_t)n_bands * W;
    // scratch: gT | dT | [d2] (n i32 each) | g (n u16, padded) | band_first, band_last, above, below (bw i32 each) | any_flag
    //          | row_flag (H) | stack (n x 16 B, slow path only -- allocated always so the call never has to synchronise)
    const size_t need = n * 4 * 3 + ((n * 2 + 255) & ~(size_t)255) + bw * 4 * 4 + 256 + ((size_t)H + 255) + n * sizeof(UamEdtEntry) + 512;
    UAM_TRY(uam_reserve(ctx, &ctx->d_scratch, &ctx->scratch_bytes, need));
    int* gT = (int*)ctx->d_scratch;
    int* dT = gT + n;
    int* d2 = d_dist2 ? d_dist2 : dT + n;
    unsigned short* g = (unsigned short*)(dT + 2 * n);
    int* band_first = (int*)((char*)g + ((n * 2 + 255) & ~(size_t)255));
    int* band_last = band_first + bw;
    int* above = band_last + bw;
    int* below = above + bw;
    int* any_flag = below + bw;
    uint8_t* row_flag = (uint8_t*)(any_flag + 16);
    UamEdtEntry* stack = (UamEdtEntry*)(((uintptr_t)(row_flag + H) + 255) & ~(uintptr_t)255);
    UAM_CUDA(ctx, cudaMemsetAsync(any_flag, 0, 64 + (size_t)H, st));
    // phase 1
    dim3 bgrid((W + 127) / 128, n_bands);
    uam_k_edt_band_summary<<<bgrid, 128, 0, st>>>(d_occ, H, W, band_first, band_last);
    UAM_CHECK_LAUNCH(ctx, "uam_k_edt_band_summary");
    uam_k_edt_band_carry<<<(W + 127) / 128, 128, 0, st>>>(n_bands, W, band_first, band_last, above, below);
    UAM_CHECK_LAUNCH(ctx, "uam_k_edt_band_carry");
    if ((W & 3) == 0 && ((((uintptr_t)d_occ) & 3) == 0)) {
        dim3 bgrid4((W / 4 + 127) / 128, n_bands);
        uam_k_edt_band_sweep<4><<<bgrid4, 128, 0, st>>>(d_occ, H, W, above, below, g);
    } else {
        uam_k_edt_band_sweep<1><<<bgrid, 128, 0, st>>>(d_occ, H, W, above, below, g);
    }
    UAM_CHECK_LAUNCH(ctx, "uam_k_edt_band_sweep");
    // phase 2, fast path (writes d2 and, when asked for, the clearance)
    const float cellf = (float)cell;
    dim3 fgrid((W + UAM_EDT_SPAN - 1) / UAM_EDT_SPAN, H);
    uam_k_edt_rows_fast<<<fgrid, 256, 0, st>>>(g, H, W, d2, d_clearance, cellf, row_flag, any_flag);
    UAM_CHECK_LAUNCH(ctx, "uam_k_edt_rows_fast");
    // phase 2, slow path for flagged rows (every kernel returns immediately when none is flagged)
    const int tgrid = ctx->sm_count * 8;
    uam_k_transpose_i32<unsigned short><<<tgrid, 256, 0, st>>>(g, H, W, gT, any_flag, nullptr);
    UAM_CHECK_LAUNCH(ctx, "uam_k_transpose_i32");
    uam_k_edt_rows<<<(H + 127) / 128, 128, 0, st>>>(gT, H, W, stack, dT, any_flag, row_flag);
    UAM_CHECK_LAUNCH(ctx, "uam_k_edt_rows");
    uam_k_transpose_i32<int><<<tgrid, 256, 0, st>>>(dT, W, H, d2, any_flag, row_flag);
    UAM_CHECK_LAUNCH(ctx, "uam_k_transpose_i32");
    if (d_clearance) {
        uam_k_edt_clearance_rows<<<ctx->sm_count * 4, 256, 0, st>>>(d2, H, W, cellf, d_clearance, any_flag, row_flag);
        UAM_CHECK_LAUNCH(ctx, "uam_k_edt_clearance_rows");
    }
    return UAM_OK;
}
