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
// true if inequality r is > thr everywhere on the tile [xa,xb] x [ya,yb] (with a rounding margin)
__device__ __forceinline__ bool uam_edge_excludes_tile(const UamEdge& r, double xa, double xb, double ya, double yb,
                                                       double thr) {
    const int kind = (int)r.kind;
    double hmin, scale;
    if (kind == UAM_EDGE_ELLIPSE) {
        const double px = fmin(fmax(r.p0, xa), xb), py = fmin(fmax(r.p1, ya), yb);
        hmin = uam_h_exact(r, px, py);
        scale = fabs(hmin) + 2.0;
    } else {
        const double h0 = uam_h_exact(r, xa, ya), h1 = uam_h_exact(r, xb, ya);
        const double h2 = uam_h_exact(r, xa, yb), h3 = uam_h_exact(r, xb, yb);
        hmin = fmin(fmin(h0, h1), fmin(h2, h3));
        if (kind == UAM_EDGE_LINE) {
            const double mx = fmax(fabs(xa - r.p0), fabs(xb - r.p0)), my = fmax(fabs(ya - r.p1), fabs(yb - r.p1));
            scale = fabs(r.p3) * mx + fabs(r.p2) * my;
        } else {
            scale = fmax(fabs(xa), fabs(xb)) + fmax(fabs(ya), fabs(yb)) + fabs(r.p2) + fabs(r.p3);
        }
    }
    return hmin > thr + 1e-9 * scale + 1e-300;
}

// Ordered compaction of the shapes [s_begin, s_end) that survive the tile test into list[0..n) (n <= CAP).
// Returns the next shape index to continue from.  Must be called by all threads of a 256-thread CTA.
__device__ int uam_cull_shapes(const UamEdge* __restrict__ edges, const UamShape* __restrict__ shapes, int s_begin,
                               int s_end, double xa, double xb, double ya, double yb, double thr, int* list, int* n_out,
                               int* warp_cnt) {
    int n = 0;
    int s0 = s_begin;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    while (s0 < s_end && n + (int)blockDim.x <= UAM_LIST_CAP) {
        const int s = s0 + threadIdx.x;
        bool keep = false;
        if (s < s_end) {
            const int4 meta = __ldg(reinterpret_cast<const int4*>(&shapes[s].e0));
            keep = true;
            for (int i = meta.x; i < meta.y && keep; ++i) {
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

__device__ __forceinline__ double uam_cell_centre(int j, double x0, double dx) {
    return __dadd_rn(x0, __dmul_rn(__dadd_rn((double)j, 0.5), dx));     // x0 + (j + 1/2) * dx
}

// -------------------------------------------------------------------------------------------------------
// occupancy
// -------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
uam_k_rasterize_occupancy(const UamEdge* __restrict__ edges, const UamShape* __restrict__ shapes, int n_obs, int H,
                          int W, double x0, double dx, double y0, double dy, uint8_t* __restrict__ occ) {
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
    int s_next = 0;
    while (s_next < n_obs) {
        int n;
        s_next = uam_cull_shapes(edges, shapes, s_next, n_obs, xa, xb, ya, yb, 1e-14, list, &n, warp_cnt);
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
// penalty layers (smooth psi only)
// -------------------------------------------------------------------------------------------------------
struct UamRegionRanges2 {
    int begin[UAM_MAX_REGIONS + 1];
};

__global__ void __launch_bounds__(256)
uam_k_rasterize_layers(const UamEdge* __restrict__ edges, const UamShape* __restrict__ shapes,
                       const double* __restrict__ psic, UamRegionRanges2 rr, int n_regions, int H, int W, double x0,
                       double dx, double y0, double dy, double e, float* __restrict__ layers) {
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
    for (int r = 0; r < n_regions; ++r) {
        double tot[4] = {0.0, 0.0, 0.0, 0.0};
        int s_next = rr.begin[r];
        const int s_end = rr.begin[r + 1];
        while (s_next < s_end) {
            int n;
            // psi != 0 needs h_i - e < 0 for every i: cull when some h_i > e on the whole tile
            s_next = uam_cull_shapes(edges, shapes, s_next, s_end, xa, xb, ya, yb, e, list, &n, warp_cnt);
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

// -------------------------------------------------------------------------------------------------------
// exact EDT: column scan -> transpose -> per-row lower envelope (Meijster) -> transpose back
// -------------------------------------------------------------------------------------------------------
#define UAM_GINF (1 << 20)   // "no occupied cell in this column" (> any real distance; its square fits int64)

// thread per column: g[i][j] = distance (cells) to the nearest occupied cell in column j
__global__ void __launch_bounds__(128)
uam_k_edt_columns(const uint8_t* __restrict__ occ, int H, int W, int* __restrict__ g) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= W) return;
    int d = UAM_GINF;
    for (int i = 0; i < H; ++i) {
        d = occ[(size_t)i * W + j] ? 0 : min(d + 1, UAM_GINF);
        g[(size_t)i * W + j] = d;
    }
    d = UAM_GINF;
    for (int i = H - 1; i >= 0; --i) {
        const int cur = g[(size_t)i * W + j];
        d = cur == 0 ? 0 : min(d + 1, UAM_GINF);
        if (d < cur) g[(size_t)i * W + j] = d;
    }
}

// out[c][r] = in[r][c]   (in: R x C)
__global__ void __launch_bounds__(256)
uam_k_transpose_i32(const int* __restrict__ in, int R, int C, int* __restrict__ out) {
    __shared__ int tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int k = ty; k < 32; k += 8) {
        const int r = r0 + k, c = c0 + tx;
        if (r < R && c < C) tile[k][tx] = in[(size_t)r * C + c];
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
        const int c = c0 + k, r = r0 + tx;
        if (r < R && c < C) out[(size_t)c * R + r] = tile[tx][k];
    }
}

struct __align__(16) UamEdtEntry {
    int s, t;
    long long gsq;
};

// thread per row i; gT[u][i] = g(i, u).  Stack entry q of row i lives at stack[q * H + i].
__global__ void __launch_bounds__(128)
uam_k_edt_rows(const int* __restrict__ gT, int H, int W, UamEdtEntry* __restrict__ stack, int* __restrict__ dT) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= H) return;
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

__global__ void __launch_bounds__(256)
uam_k_edt_clearance(const int* __restrict__ d2, long long n, double cell, float* __restrict__ clearance) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += stride)
        clearance[c] = (float)(sqrt((double)d2[c]) * cell);
}

}  // namespace

extern "C" int uam_dem_mask(uam_ctx* ctx, const float* d_image, int64_t n, float threshold, uint8_t* d_mask,
                            void* stream) {
    if (!ctx) return UAM_ERR_INVALID;
    if (n < 0 || (n > 0 && (!d_image || !d_mask))) return uam_fail(ctx, UAM_ERR_INVALID, "bad argument to uam_dem_mask");
    if (n == 0) return UAM_OK;
    UAM_CUDA(ctx, cudaSetDevice(ctx->device));
    const int eq_mode = threshold == -9999.0f;      // data_manager.py:14-15
    const long long ctas = std::min<long long>((n / 4 + 255) / 256 + 1, (long long)ctx->sm_count * 16);
    uam_k_dem_mask<<<(unsigned)ctas, 256, 0, uam_pick_stream(ctx, stream)>>>(d_image, n, threshold, eq_mode, d_mask);
    UAM_CHECK_LAUNCH(ctx, "uam_k_dem_mask");
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
    UAM_CUDA(ctx, cudaSetDevice(ctx->device));
    dim3 grid((W + UAM_TILE_W - 1) / UAM_TILE_W, (H + UAM_TILE_H - 1) / UAM_TILE_H);
    if (grid.y > 65535) return uam_fail(ctx, UAM_ERR_UNSUPPORTED, "H too large");
    uam_k_rasterize_occupancy<<<grid, 256, 0, uam_pick_stream(ctx, stream)>>>(ctx->d_edges, ctx->d_shapes, ctx->n_obs, H, W,
                                                                               x0, dx, y0, dy, d_occ);
    UAM_CHECK_LAUNCH(ctx, "uam_k_rasterize_occupancy");
    return UAM_OK;
}

extern "C" int uam_rasterize_layers(uam_ctx* ctx, int H, int W, double x0, double dx, double y0, double dy,
                                    double enlargement, float* d_layers, void* stream) {
    UAM_TRY(uam_check_grid(ctx, H, W, dx, dy, d_layers));
    if (ctx->n_regions < 1) return uam_fail(ctx, UAM_ERR_STATE, "the map has no regions");
    UAM_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = uam_pick_stream(ctx, stream);
    UamParams prm = {};
    prm.e = enlargement;
    prm.flags = UAM_PENALTY_SMOOTH | UAM_OBSTACLE_SMOOTH;
    UAM_TRY(uam_ensure_shape_norm(ctx, prm, st));
    UamRegionRanges2 rr;
    for (int r = 0; r <= ctx->n_regions; ++r) rr.begin[r] = ctx->region_begin[r];
    dim3 grid((W + UAM_TILE_W - 1) / UAM_TILE_W, (H + UAM_TILE_H - 1) / UAM_TILE_H);
    if (grid.y > 65535) return uam_fail(ctx, UAM_ERR_UNSUPPORTED, "H too large");
    uam_k_rasterize_layers<<<grid, 256, 0, st>>>(ctx->d_edges, ctx->d_shapes, ctx->d_psic, rr, ctx->n_regions, H, W, x0, dx,
                                                 y0, dy, enlargement, d_layers);
    UAM_CHECK_LAUNCH(ctx, "uam_k_rasterize_layers");
    return UAM_OK;
}

extern "C" int uam_edt(uam_ctx* ctx, const uint8_t* d_occ, int H, int W, double cell, int32_t* d_dist2,
                       float* d_clearance, void* stream) {
    if (!ctx) return UAM_ERR_INVALID;
    if (H < 1 || W < 1 || !d_occ || (!d_dist2 && !d_clearance)) return uam_fail(ctx, UAM_ERR_INVALID, "bad argument to uam_edt");
    if (H > 23170 || W > 23170) return uam_fail(ctx, UAM_ERR_UNSUPPORTED, "raster larger than 23170 per side (d^2 must stay below 2^30)");
    UAM_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = uam_pick_stream(ctx, stream);
    const size_t n = (size_t)H * W;
    // scratch: g (n i32) | gT (n i32) | dT (n i32) | [d2 when the caller wants only clearance] | stack (n x 16 B)
    const size_t need = n * 4 * 4 + n * sizeof(UamEdtEntry) + 256;
    UAM_TRY(uam_reserve(ctx, &ctx->d_scratch, &ctx->scratch_bytes, need));
    int* g = (int*)ctx->d_scratch;
    int* gT = g + n;
    int* dT = gT + n;
    int* d2 = d_dist2 ? d_dist2 : dT + n;
    UamEdtEntry* stack = (UamEdtEntry*)(((uintptr_t)(dT + 2 * n) + 15) & ~(uintptr_t)15);
    uam_k_edt_columns<<<(W + 127) / 128, 128, 0, st>>>(d_occ, H, W, g);
    UAM_CHECK_LAUNCH(ctx, "uam_k_edt_columns");
    dim3 tg((W + 31) / 32, (H + 31) / 32);
    uam_k_transpose_i32<<<tg, 256, 0, st>>>(g, H, W, gT);
    UAM_CHECK_LAUNCH(ctx, "uam_k_transpose_i32");
    uam_k_edt_rows<<<(H + 127) / 128, 128, 0, st>>>(gT, H, W, stack, dT);
    UAM_CHECK_LAUNCH(ctx, "uam_k_edt_rows");
    dim3 tg2((H + 31) / 32, (W + 31) / 32);
    uam_k_transpose_i32<<<tg2, 256, 0, st>>>(dT, W, H, d2);
    UAM_CHECK_LAUNCH(ctx, "uam_k_transpose_i32");
    if (d_clearance) {
        uam_k_edt_clearance<<<ctx->sm_count * 8, 256, 0, st>>>(d2, (long long)n, cell, d_clearance);
        UAM_CHECK_LAUNCH(ctx, "uam_k_edt_clearance");
    }
    return UAM_OK;
}
