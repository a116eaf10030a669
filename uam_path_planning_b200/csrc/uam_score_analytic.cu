// Analytic path scorer: Problem.get_cost / get_nonlincon / Map.collides batched over candidate paths,
// plus the point-query form (get_penalty_function / get_total_penalty_function / collides).
// Everything here is fp64 in the reference's operation order (explicit round-to-nearest intrinsics, no
// FMA contraction), so inequalities, collision flags and the zero pattern of g are bit-exact and costs
// differ from the reference only by the order in which the per-waypoint terms are summed.
#include <algorithm>

#include "uam_internal.cuh"

namespace {

// np.maximum(0.0, v): NaN propagates
__device__ __forceinline__ double uam_relu_nan(double v) { return (v > 0.0 || v != v) ? v : 0.0; }

__device__ __forceinline__ double uam_norm2(double dx, double dy) {
    return sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}

// one term of a region sum: total += psi_s(x)/psi_s(center_s)     problem.py:76-80 (without the weight)
__device__ __forceinline__ void uam_add_shape_term(const UamEdge* __restrict__ edges, const double* __restrict__ psic,
                                                   int s, const int4 meta, double x, double y, bool smooth, double e,
                                                   bool early_exit, double& total) {
    const double psi = uam_psi(edges, meta.x, meta.y, x, y, smooth, e, nullptr, early_exit);
    if (meta.w) {
        const double pc = __ldg(psic + s);
        // x + 0/pc == x unless pc == 0 (0/0 = NaN must propagate like the reference)
        if (psi != 0.0 || pc == 0.0 || pc != pc) total = __dadd_rn(total, __ddiv_rn(psi, pc));
    } else {
        total = __dadd_rn(total, psi);
    }
}

// sum over the region's shapes of psi(x)/psi(center)     problem.py:72-80 (without the weight)
__device__ __forceinline__ double uam_region_total(const UamEdge* __restrict__ edges, const UamShape* __restrict__ shapes,
                                                   const double* __restrict__ psic, int s0, int s1, double x, double y,
                                                   bool smooth, double e, bool early_exit) {
    double total = 0.0;
    for (int s = s0; s < s1; ++s) {
        const int4 meta = __ldg(reinterpret_cast<const int4*>(&shapes[s].e0));   // e0, e1, region, has_center
        uam_add_shape_term(edges, psic, s, meta, x, y, smooth, e, early_exit, total);
    }
    return total;
}

// The same sum over the candidates items[i..i1) that belong to shapes [s0, s1) (ascending ids; the shape grid's cell
// list): the shapes left out contribute exact zeros.  Advances i past the range.
__device__ __forceinline__ double uam_region_total_listed(const UamEdge* __restrict__ edges, const UamShape* __restrict__ shapes,
                                                          const double* __restrict__ psic, const int* __restrict__ items,
                                                          int& i, int i1, int s1, double x, double y, bool smooth, double e) {
    double total = 0.0;
    for (; i < i1; ++i) {
        const int s = __ldg(items + i);
        if (s >= s1) break;
        const int4 meta = __ldg(reinterpret_cast<const int4*>(&shapes[s].e0));
        uam_add_shape_term(edges, psic, s, meta, x, y, smooth, e, true, total);
    }
    return total;
}

struct UamRegionRanges {
    int begin[UAM_MAX_REGIONS + 1];
};

// One warp per path; lanes stride the N+2 waypoints.
__global__ void __launch_bounds__(UAM_CTA_THREADS)
uam_k_score_analytic(const double2* __restrict__ z, long long B, int N, UamParams prm, UamRegionRanges rr, UamShapeGrid sg,
                     const UamEdge* __restrict__ edges, const UamShape* __restrict__ shapes,
                     const double* __restrict__ psic, int n_obs, double* __restrict__ cost,
                     uint8_t* __restrict__ collide, double* __restrict__ g) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * UAM_WARPS_PER_CTA + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * UAM_WARPS_PER_CTA;
    const int W = N + 2;
    const bool pen_smooth = (prm.flags & UAM_PENALTY_SMOOTH) != 0;
    const bool obs_smooth = (prm.flags & UAM_OBSTACLE_SMOOTH) != 0;
    const bool fast = (prm.flags & UAM_INTERNAL_FINITE_EDGES) != 0;
    const bool len_smooth = (prm.flags & UAM_LENGTH_SMOOTH) != 0;
    const bool mr_smooth = (prm.flags & UAM_MAXRATIO_SMOOTH) != 0;
    const double mr = mr_smooth ? __dmul_rn(prm.maxratio, prm.maxratio) : prm.maxratio;   // problem.py:95-96
    const int glen = 3 * N + n_obs * W;
    const double dN = (double)N;

    for (long long path = warp0; path < B; path += nwarps) {
        const double2* zp = z + path * W;
        double* gp = g ? g + path * (long long)glen : nullptr;
        double pen_sum = 0.0, len_sum = 0.0;
        bool col = false;
        for (int j = lane; j < W; j += 32) {
            const double2 p = zp[j];
            // cell of the shape grid (-1: no grid, or the point is outside it / not finite -> every shape is evaluated)
            const int cell = uam_shape_grid_cell(sg, p.x, p.y);
            int li = 0, l1 = 0;
            if (cell >= 0) { li = __ldg(sg.start + cell); l1 = __ldg(sg.start + cell + 1); }
            // ---- hard obstacles: collision (map.py:41-43) and g's obstacle block (problem.py:109-112) ----
            if (cell < 0 || (gp && !sg.obs_values)) {
                for (int o = 0; o < n_obs; ++o) {
                    const int4 meta = __ldg(reinterpret_cast<const int4*>(&shapes[o].e0));
                    bool inside;
                    const double psi = uam_psi(edges, meta.x, meta.y, p.x, p.y, obs_smooth, 0.0, &inside, fast);
                    col = col || inside;
                    if (gp) gp[3 * N + o * W + j] = psi;
                }
            } else if (gp) {
                for (int o = 0; o < n_obs; ++o) {
                    double psi = 0.0;          // an obstacle that is not in the cell's list: psi = 0, not inside
                    if (li < l1 && __ldg(sg.items + li) == o) {
                        const int4 meta = __ldg(reinterpret_cast<const int4*>(&shapes[o].e0));
                        bool inside;
                        psi = uam_psi(edges, meta.x, meta.y, p.x, p.y, true, 0.0, &inside, true);
                        col = col || inside;
                        ++li;
                    }
                    gp[3 * N + o * W + j] = psi;
                }
            } else {
                for (; li < l1; ++li) {
                    const int o = __ldg(sg.items + li);
                    if (o >= n_obs) break;
                    const int4 meta = __ldg(reinterpret_cast<const int4*>(&shapes[o].e0));
                    bool inside;
                    uam_psi(edges, meta.x, meta.y, p.x, p.y, true, 0.0, &inside, true);     // `contains` only
                    col = col || inside;
                }
            }
            // ---- weighted region penalties at z_j, regions in insertion order (problem.py:49-56) ----
            double P = 0.0;
            if (cell < 0) {
                for (int r = 0; r < prm.n_regions; ++r) {
                    const double tot = uam_region_total(edges, shapes, psic, rr.begin[r], rr.begin[r + 1], p.x, p.y,
                                                        pen_smooth, prm.e, fast);
                    P = __dadd_rn(P, __dmul_rn(prm.w[r], tot));
                }
            } else {
                // only the regions with a listed shape: w_r * 0 = +-0 leaves P unchanged (finite weights: uam_pick_shape_grid)
                while (li < l1 && __ldg(sg.items + li) < n_obs) ++li;
                while (li < l1) {
                    const int r = __ldg(&shapes[__ldg(sg.items + li)].region);
                    const double tot = uam_region_total_listed(edges, shapes, psic, sg.items, li, l1, rr.begin[r + 1], p.x, p.y,
                                                               pen_smooth, prm.e);
                    P = __dadd_rn(P, __dmul_rn(prm.w[r], tot));
                }
            }
            pen_sum += __ddiv_rn(P, dN);                                   // problem.py:43
            // ---- segment pair k = j: length term (problem.py:130-146) and ratio/angle block (:100-107) ----
            if (j < N) {
                const double2 q = zp[j + 1], r2 = zp[j + 2];
                const double ax = __dsub_rn(q.x, p.x), ay = __dsub_rn(q.y, p.y);
                const double bx = __dsub_rn(r2.x, q.x), by = __dsub_rn(r2.y, q.y);
                const double na = uam_norm2(ax, ay), nb = uam_norm2(bx, by);
                len_sum += len_smooth ? __dmul_rn(na, na) : na;
                if (gp) {
                    const double a = mr_smooth ? __dmul_rn(na, na) : na;
                    const double b = mr_smooth ? __dmul_rn(nb, nb) : nb;
                    const double dot = __dadd_rn(__dmul_rn(ax, bx), __dmul_rn(ay, by));
                    const double cos_t = __ddiv_rn(dot, __dmul_rn(a, b));
                    gp[3 * j + 0] = uam_relu_nan(__dsub_rn(b, __dmul_rn(mr, a)));
                    gp[3 * j + 1] = uam_relu_nan(__dsub_rn(__ddiv_rn(a, mr), b));
                    gp[3 * j + 2] = uam_relu_nan(__dsub_rn(prm.mincos, cos_t));
                }
            }
            if (j == 0 && !(prm.flags & UAM_OWN_START)) {
                // first pair of length_of: (map.x_start, z_0)   -- the reference passes z_ that already holds
                // the start, so this term is |z_0 - map.x_start| and the last segment z_N -> goal is absent
                const double n0 = uam_norm2(__dsub_rn(p.x, prm.ms_x), __dsub_rn(p.y, prm.ms_y));
                len_sum += len_smooth ? __dmul_rn(n0, n0) : n0;
            }
        }
        pen_sum = uam_warp_sum(pen_sum);
        len_sum = uam_warp_sum(len_sum);
        col = __any_sync(0xffffffffu, col);
        if (lane == 0) {
            if (cost) cost[path] = (double)(N + 1) * len_sum + pen_sum;    // problem.py:41-44
            if (collide) collide[path] = col ? 1 : 0;
        }
    }
}

// One thread per query point.
__global__ void __launch_bounds__(256)
uam_k_eval_points(const double2* __restrict__ x, long long M, UamParams prm, UamRegionRanges rr, UamShapeGrid sg,
                  const UamEdge* __restrict__ edges, const UamShape* __restrict__ shapes,
                  const double* __restrict__ psic, int n_obs, double* __restrict__ region_pen,
                  double* __restrict__ obst_pen, uint8_t* __restrict__ collide) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const bool pen_smooth = (prm.flags & UAM_PENALTY_SMOOTH) != 0;
    const bool obs_smooth = (prm.flags & UAM_OBSTACLE_SMOOTH) != 0;
    const bool fast = (prm.flags & UAM_INTERNAL_FINITE_EDGES) != 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += stride) {
        const double2 p = x[i];
        const int cell = uam_shape_grid_cell(sg, p.x, p.y);
        if (cell >= 0) {
            // candidates of the cell only (ascending: obstacles, then region by region); the rest are exact zeros
            const int l0 = __ldg(sg.start + cell), l1 = __ldg(sg.start + cell + 1);
            int li = l0;
            if (obst_pen)
                obst_pen[i] = sg.obs_values ? uam_region_total_listed(edges, shapes, psic, sg.items, li, l1, n_obs, p.x, p.y, true, prm.e)
                                            : uam_region_total(edges, shapes, psic, 0, n_obs, p.x, p.y, obs_smooth, prm.e, fast);
            if (collide) {
                bool col = false;
                for (li = l0; li < l1; ++li) {
                    const int o = __ldg(sg.items + li);
                    if (o >= n_obs) break;
                    const int4 meta = __ldg(reinterpret_cast<const int4*>(&shapes[o].e0));
                    bool inside;
                    uam_psi(edges, meta.x, meta.y, p.x, p.y, true, 0.0, &inside, true);
                    col = col || inside;
                }
                collide[i] = col ? 1 : 0;
            }
            if (region_pen) {
                li = l0;
                while (li < l1 && __ldg(sg.items + li) < n_obs) ++li;
                for (int r = 0; r < prm.n_regions; ++r) {
                    const double tot = uam_region_total_listed(edges, shapes, psic, sg.items, li, l1, rr.begin[r + 1], p.x, p.y,
                                                               pen_smooth, prm.e);
                    region_pen[i * prm.n_regions + r] = __dmul_rn(prm.w[r], tot);
                }
            }
            continue;
        }
        if (region_pen) {
            for (int r = 0; r < prm.n_regions; ++r) {
                const double tot = uam_region_total(edges, shapes, psic, rr.begin[r], rr.begin[r + 1], p.x, p.y,
                                                    pen_smooth, prm.e, fast);
                region_pen[i * prm.n_regions + r] = __dmul_rn(prm.w[r], tot);
            }
        }
        if (obst_pen)   // get_penalty_function(None): w = 1, obstacle_smooth, params['enlargement']
            obst_pen[i] = uam_region_total(edges, shapes, psic, 0, n_obs, p.x, p.y, obs_smooth, prm.e, fast);
        if (collide) {
            bool col = false;
            for (int o = 0; o < n_obs; ++o) {
                const int4 meta = __ldg(reinterpret_cast<const int4*>(&shapes[o].e0));
                bool inside;
                uam_psi(edges, meta.x, meta.y, p.x, p.y, true, 0.0, &inside, fast);
                col = col || inside;
            }
            collide[i] = col ? 1 : 0;
        }
    }
}


// Problem.length_of(x, smooth)   problem.py:130-146:  y = [map.x_start; x; map.x_goal] (M + 2 points),
// out = sum of nrm(y_{k+1} - y_k) over the FIRST N+1 pairs only.  One warp per row.
__global__ void __launch_bounds__(UAM_CTA_THREADS)
uam_k_length_of(const double2* __restrict__ x, long long B, int M, int N, double msx, double msy, double mgx,
                double mgy, int smooth, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * UAM_WARPS_PER_CTA + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * UAM_WARPS_PER_CTA;
    for (long long row = warp0; row < B; row += nwarps) {
        const double2* xp = x + row * M;
        double acc = 0.0;
        for (int k = lane; k <= N; k += 32) {
            const double2 a = k == 0 ? make_double2(msx, msy) : xp[k - 1];
            const double2 b = k == M ? make_double2(mgx, mgy) : xp[k];
            const double d = uam_norm2(__dsub_rn(b.x, a.x), __dsub_rn(b.y, a.y));
            acc += smooth ? __dmul_rn(d, d) : d;
        }
        acc = uam_warp_sum(acc);
        if (lane == 0) out[row] = acc;
    }
}

// h_i(x_m) for n_rec raw inequality records at M points (Function.__call__, function.py:119-120): out[i*M + m]
__global__ void __launch_bounds__(256)
uam_k_eval_inequalities(const UamEdge* __restrict__ recs, int n_rec, const double2* __restrict__ x, long long M,
                        double* __restrict__ out) {
    const long long total = (long long)n_rec * M;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int i = (int)(t / M);
        const long long m = t - (long long)i * M;
        const double2 p = x[m];
        out[t] = uam_h_exact(uam_load_edge(recs + i), p.x, p.y);
    }
}

// need_norm = false: the call only asks for `contains` (no psi value is divided by psi(centre)), so the psi(centre)
// table and the shape grid are left as they are (a grid built for any enlargement / flags culls `contains` correctly)
int uam_analytic_prepare(uam_ctx* ctx, const double* h_p, int n_p, int flags, cudaStream_t st, UamParams* prm,
                         UamRegionRanges* rr, bool need_norm = true) {
    if (!ctx->has_shapes) return uam_fail(ctx, UAM_ERR_STATE, "no shape table: call uam_map_set_shapes first");
    UAM_TRY(uam_make_params(ctx, h_p, n_p, flags, prm));
    if (prm->n_regions != ctx->n_regions)
        return uam_fail(ctx, UAM_ERR_INVALID, "p carries %d region weights, the map has %d regions", prm->n_regions,
                        ctx->n_regions);
    UAM_CUDA(ctx, cudaSetDevice(ctx->device));
    if (need_norm) UAM_TRY(uam_ensure_shape_norm(ctx, *prm, st));
    for (int r = 0; r <= ctx->n_regions; ++r) rr->begin[r] = ctx->region_begin[r];
    if (ctx->edges_finite) prm->flags |= UAM_INTERNAL_FINITE_EDGES;
    return UAM_OK;
}

}  // namespace

extern "C" int uam_score_paths_analytic(uam_ctx* ctx, const double* d_z, int64_t B, int N, const double* h_p, int n_p,
                                        int flags, double* d_cost, uint8_t* d_collide, double* d_g, void* stream) {
    if (!ctx) return UAM_ERR_INVALID;
    UAM_NVTX("uam.analytic.score");
    if (B < 0 || N < 1) return uam_fail(ctx, UAM_ERR_INVALID, "need B >= 0 and N >= 1 (got B=%lld N=%d)", (long long)B, N);
    if (B > 0 && !d_z) return uam_fail(ctx, UAM_ERR_INVALID, "paths pointer is NULL");
    cudaStream_t st = uam_pick_stream(ctx, stream);
    UamParams prm;
    UamRegionRanges rr;
    UAM_TRY(uam_analytic_prepare(ctx, h_p, n_p, flags, st, &prm, &rr));
    if (B == 0) return UAM_OK;
    const long long ctas = std::min<long long>((B + UAM_WARPS_PER_CTA - 1) / UAM_WARPS_PER_CTA, (long long)ctx->sm_count * 8);
    uam_k_score_analytic<<<(unsigned)ctas, UAM_CTA_THREADS, 0, st>>>(
        reinterpret_cast<const double2*>(d_z), B, N, prm, rr, uam_pick_shape_grid(ctx, prm), ctx->d_edges, ctx->d_shapes, ctx->d_psic,
        ctx->n_obs, d_cost, d_collide, d_g);
    UAM_CHECK_LAUNCH(ctx, "uam_k_score_analytic");
    return UAM_OK;
}

extern "C" int uam_analytic_g_len(const uam_ctx* ctx, int N, int64_t* len) {
    if (!ctx || !len) return UAM_ERR_INVALID;
    *len = 3 * (int64_t)N + (int64_t)ctx->n_obs * (N + 2);
    return UAM_OK;
}

extern "C" int uam_score_paths_analytic_host(uam_ctx* ctx, const double* h_z, int64_t B, int N, const double* h_p,
                                             int n_p, int flags, double* h_cost, uint8_t* h_collide, double* h_g) {
    if (!ctx) return UAM_ERR_INVALID;
    if (B < 0 || N < 1) return uam_fail(ctx, UAM_ERR_INVALID, "need B >= 0 and N >= 1 (got B=%lld N=%d)", (long long)B, N);
    if (B == 0) return UAM_OK;
    if (!h_z) return uam_fail(ctx, UAM_ERR_INVALID, "paths pointer is NULL");
    UAM_CUDA(ctx, cudaSetDevice(ctx->device));
    UAM_CUDA(ctx, cudaDeviceSynchronize());
    const size_t row = (size_t)2 * (N + 2) * sizeof(double);
    const size_t glen = 3 * (size_t)N + (size_t)ctx->n_obs * (N + 2);
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(B, (int64_t)((256u << 20) / (row + (h_g ? glen * 8 : 0) + 9))));
    cudaStream_t st = ctx->stream;
    for (int64_t b0 = 0; b0 < B; b0 += chunk) {
        const int64_t nb = std::min(chunk, B - b0);
        const size_t out_bytes = (size_t)nb * 16 + (h_g ? (size_t)nb * glen * 8 : 0);
        UAM_TRY(uam_reserve(ctx, &ctx->d_stage_in[0], &ctx->stage_in_bytes[0], (size_t)nb * row));
        UAM_TRY(uam_reserve(ctx, &ctx->d_stage_out[0], &ctx->stage_out_bytes[0], out_bytes));
        double* d_cost = (double*)ctx->d_stage_out[0];
        uint8_t* d_col = (uint8_t*)(d_cost + nb);
        double* d_g = h_g ? d_cost + 2 * nb : nullptr;     // collide bytes live in [nb*8, nb*16)
        UAM_CUDA(ctx, cudaMemcpyAsync(ctx->d_stage_in[0], (const char*)h_z + (size_t)b0 * row, (size_t)nb * row,
                                      cudaMemcpyHostToDevice, st));
        UAM_TRY(uam_score_paths_analytic(ctx, (const double*)ctx->d_stage_in[0], nb, N, h_p, n_p, flags, d_cost, d_col, d_g,
                                         st));
        if (h_cost) UAM_CUDA(ctx, cudaMemcpyAsync(h_cost + b0, d_cost, (size_t)nb * 8, cudaMemcpyDeviceToHost, st));
        if (h_collide) UAM_CUDA(ctx, cudaMemcpyAsync(h_collide + b0, d_col, (size_t)nb, cudaMemcpyDeviceToHost, st));
        if (h_g) UAM_CUDA(ctx, cudaMemcpyAsync(h_g + (size_t)b0 * glen, d_g, (size_t)nb * glen * 8, cudaMemcpyDeviceToHost, st));
        UAM_CUDA(ctx, cudaStreamSynchronize(st));
    }
    return UAM_OK;
}

extern "C" int uam_eval_points(uam_ctx* ctx, const double* d_x, int64_t M, const double* h_p, int n_p, int flags,
                               double* d_region_pen, double* d_obst_pen, uint8_t* d_collide, void* stream) {
    if (!ctx) return UAM_ERR_INVALID;
    if (M < 0) return uam_fail(ctx, UAM_ERR_INVALID, "M < 0");
    if (M > 0 && !d_x) return uam_fail(ctx, UAM_ERR_INVALID, "points pointer is NULL");
    cudaStream_t st = uam_pick_stream(ctx, stream);
    UamParams prm;
    UamRegionRanges rr;
    const bool only_contains = !d_region_pen && !d_obst_pen;
    UAM_TRY(uam_analytic_prepare(ctx, h_p, n_p, flags, st, &prm, &rr, !only_contains));
    if (M == 0) return UAM_OK;
    const UamShapeGrid sg = only_contains ? (ctx->shape_grid_opt ? ctx->shape_grid : UamShapeGrid{}) : uam_pick_shape_grid(ctx, prm);
    const long long ctas = std::min<long long>((M + 255) / 256, (long long)ctx->sm_count * 8);
    uam_k_eval_points<<<(unsigned)ctas, 256, 0, st>>>(reinterpret_cast<const double2*>(d_x), M, prm, rr, sg, ctx->d_edges,
                                                       ctx->d_shapes, ctx->d_psic, ctx->n_obs, d_region_pen, d_obst_pen,
                                                       d_collide);
    UAM_CHECK_LAUNCH(ctx, "uam_k_eval_points");
    return UAM_OK;
}

extern "C" int uam_eval_points_host(uam_ctx* ctx, const double* h_x, int64_t M, const double* h_p, int n_p, int flags,
                                    double* h_region_pen, double* h_obst_pen, uint8_t* h_collide) {
    if (!ctx) return UAM_ERR_INVALID;
    if (M < 0) return uam_fail(ctx, UAM_ERR_INVALID, "M < 0");
    if (M == 0) return UAM_OK;
    if (!h_x) return uam_fail(ctx, UAM_ERR_INVALID, "points pointer is NULL");
    UAM_CUDA(ctx, cudaSetDevice(ctx->device));
    UAM_CUDA(ctx, cudaDeviceSynchronize());
    const int R = ctx->n_regions;
    cudaStream_t st = ctx->stream;
    UAM_TRY(uam_reserve(ctx, &ctx->d_stage_in[0], &ctx->stage_in_bytes[0], (size_t)M * 16));
    UAM_TRY(uam_reserve(ctx, &ctx->d_stage_out[0], &ctx->stage_out_bytes[0], (size_t)M * (8 * (size_t)R + 8 + 8)));
    double* d_rp = (double*)ctx->d_stage_out[0];
    double* d_op = d_rp + (size_t)M * R;
    uint8_t* d_col = (uint8_t*)(d_op + M);
    UAM_CUDA(ctx, cudaMemcpyAsync(ctx->d_stage_in[0], h_x, (size_t)M * 16, cudaMemcpyHostToDevice, st));
    UAM_TRY(uam_eval_points(ctx, (const double*)ctx->d_stage_in[0], M, h_p, n_p, flags, h_region_pen ? d_rp : nullptr,
                            h_obst_pen ? d_op : nullptr, h_collide ? d_col : nullptr, st));
    if (h_region_pen) UAM_CUDA(ctx, cudaMemcpyAsync(h_region_pen, d_rp, (size_t)M * R * 8, cudaMemcpyDeviceToHost, st));
    if (h_obst_pen) UAM_CUDA(ctx, cudaMemcpyAsync(h_obst_pen, d_op, (size_t)M * 8, cudaMemcpyDeviceToHost, st));
    if (h_collide) UAM_CUDA(ctx, cudaMemcpyAsync(h_collide, d_col, (size_t)M, cudaMemcpyDeviceToHost, st));
    UAM_CUDA(ctx, cudaStreamSynchronize(st));
    return UAM_OK;
}

// h_ends = [map.x_start (2), map.x_goal (2)]
extern "C" int uam_length_of(uam_ctx* ctx, const double* d_x, int64_t B, int M, int N, const double* h_ends,
                             int smooth, double* d_out, void* stream) {
    if (!ctx) return UAM_ERR_INVALID;
    if (B < 0 || M < 0 || N < 0 || N > M) return uam_fail(ctx, UAM_ERR_INVALID, "length_of: need 0 <= N <= M points (N=%d M=%d)", N, M);
    if (!h_ends || !d_out || (B > 0 && M > 0 && !d_x)) return uam_fail(ctx, UAM_ERR_INVALID, "length_of: NULL pointer");
    if (B == 0) return UAM_OK;
    UAM_CUDA(ctx, cudaSetDevice(ctx->device));
    const long long ctas = std::min<long long>((B + UAM_WARPS_PER_CTA - 1) / UAM_WARPS_PER_CTA, (long long)ctx->sm_count * 8);
    uam_k_length_of<<<(unsigned)ctas, UAM_CTA_THREADS, 0, uam_pick_stream(ctx, stream)>>>(
        reinterpret_cast<const double2*>(d_x), B, M, N, h_ends[0], h_ends[1], h_ends[2], h_ends[3], smooth, d_out);
    UAM_CHECK_LAUNCH(ctx, "uam_k_length_of");
    return UAM_OK;
}

extern "C" int uam_length_of_host(uam_ctx* ctx, const double* h_x, int64_t B, int M, int N, const double* h_ends,
                                  int smooth, double* h_out) {
    if (!ctx) return UAM_ERR_INVALID;
    if (B <= 0) return B == 0 ? UAM_OK : uam_fail(ctx, UAM_ERR_INVALID, "B < 0");
    if (!h_out) return uam_fail(ctx, UAM_ERR_INVALID, "length_of: NULL output");
    UAM_CUDA(ctx, cudaSetDevice(ctx->device));
    UAM_CUDA(ctx, cudaDeviceSynchronize());
    cudaStream_t st = ctx->stream;
    const size_t in_bytes = (size_t)B * M * 16;
    UAM_TRY(uam_reserve(ctx, &ctx->d_stage_in[0], &ctx->stage_in_bytes[0], in_bytes + 16));
    UAM_TRY(uam_reserve(ctx, &ctx->d_stage_out[0], &ctx->stage_out_bytes[0], (size_t)B * 8));
    if (in_bytes) UAM_CUDA(ctx, cudaMemcpyAsync(ctx->d_stage_in[0], h_x, in_bytes, cudaMemcpyHostToDevice, st));
    UAM_TRY(uam_length_of(ctx, (const double*)ctx->d_stage_in[0], B, M, N, h_ends, smooth, (double*)ctx->d_stage_out[0], st));
    UAM_CUDA(ctx, cudaMemcpyAsync(h_out, ctx->d_stage_out[0], (size_t)B * 8, cudaMemcpyDeviceToHost, st));
    UAM_CUDA(ctx, cudaStreamSynchronize(st));
    return UAM_OK;
}

extern "C" int uam_eval_inequalities_host(uam_ctx* ctx, const double* h_records, int n_rec, const double* h_x, int64_t M,
                                          double* h_out) {
    if (!ctx) return UAM_ERR_INVALID;
    if (n_rec < 0 || M < 0) return uam_fail(ctx, UAM_ERR_INVALID, "negative size");
    if (n_rec == 0 || M == 0) return UAM_OK;
    if (!h_records || !h_x || !h_out) return uam_fail(ctx, UAM_ERR_INVALID, "NULL pointer");
    for (int i = 0; i < n_rec; ++i) {
        const int k = (int)h_records[8 * (size_t)i];
        if (k != UAM_EDGE_LINE && k != UAM_EDGE_ELLIPSE && k != UAM_EDGE_BOX)
            return uam_fail(ctx, UAM_ERR_INVALID, "inequality %d: unknown kind %d", i, k);
    }
    UAM_CUDA(ctx, cudaSetDevice(ctx->device));
    UAM_CUDA(ctx, cudaDeviceSynchronize());
    cudaStream_t st = ctx->stream;
    const size_t rb = (size_t)n_rec * sizeof(UamEdge), xb = (size_t)M * 16, ob = (size_t)n_rec * M * 8;
    UAM_TRY(uam_reserve(ctx, &ctx->d_stage_in[0], &ctx->stage_in_bytes[0], rb + xb));
    UAM_TRY(uam_reserve(ctx, &ctx->d_stage_out[0], &ctx->stage_out_bytes[0], ob));
    UamEdge* d_rec = (UamEdge*)ctx->d_stage_in[0];
    double2* d_x = (double2*)((char*)ctx->d_stage_in[0] + rb);
    UAM_CUDA(ctx, cudaMemcpyAsync(d_rec, h_records, rb, cudaMemcpyHostToDevice, st));
    UAM_CUDA(ctx, cudaMemcpyAsync(d_x, h_x, xb, cudaMemcpyHostToDevice, st));
    const long long total = (long long)n_rec * M;
    const long long ctas = std::min<long long>((total + 255) / 256, (long long)ctx->sm_count * 8);
    uam_k_eval_inequalities<<<(unsigned)ctas, 256, 0, st>>>(d_rec, n_rec, d_x, M, (double*)ctx->d_stage_out[0]);
    UAM_CHECK_LAUNCH(ctx, "uam_k_eval_inequalities");
    UAM_CUDA(ctx, cudaMemcpyAsync(h_out, ctx->d_stage_out[0], ob, cudaMemcpyDeviceToHost, st));
    UAM_CUDA(ctx, cudaStreamSynchronize(st));
    return UAM_OK;
}
