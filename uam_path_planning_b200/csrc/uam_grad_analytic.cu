// Gradient of Problem.get_cost (path_generation/problem.py:38-44) with respect to every waypoint, batched.
// The reference obtains this derivative from CasADi's algorithmic differentiation inside the OpEn solver
// (solver.py:82-101); here it is written out analytically (SURVEY.md 8f item 1 -- the step after scoring):
//   cost = (N+1) L(z) + (1/N) sum_j P(z_j)
//   L    = nrm(z_0 - m_s) + sum_{k=0}^{N-1} nrm(z_{k+1} - z_k)      (the last segment is absent, as in get_cost)
//   P(x) = sum_r w_r sum_s psi_s(x; e) / psi_s(c_s; e),   psi = prod_i m_i^2,  m_i = min(h_i - e, 0)
//   grad psi = psi * sum_i 2 grad h_i / m_i  where every m_i < 0, and 0 as soon as one m_i = 0
// fp64, one warp per path, lanes stride the waypoints.  Smooth penalties only (the non-smooth form is NaN in the
// reference, quirk Q4).
#include <algorithm>

#include "uam_internal.cuh"

namespace {

struct UamRegionRanges3 {
    int begin[UAM_MAX_REGIONS + 1];
};

__device__ __forceinline__ void uam_grad_h(const UamEdge& r, double x, double y, double& gx, double& gy) {
    const int kind = (int)r.kind;
    if (kind == UAM_EDGE_LINE) {            // h = -sgn ((By-Ay)(x-Ax) - (Bx-Ax)(y-Ay))
        gx = -r.p4 * r.p3;
        gy = r.p4 * r.p2;
    } else if (kind == UAM_EDGE_ELLIPSE) {  // h = ((x-cx)/r1)^2 + ((y-cy)/r2)^2 - 1
        gx = 2.0 * ((x - r.p0) / r.p2) / r.p2;
        gy = 2.0 * ((y - r.p1) / r.p3) / r.p3;
    } else {                                // h = +-(x_d - c) - r
        const double s = r.p1 > 0 ? 1.0 : -1.0;
        gx = ((int)r.p0 == 0) ? s : 0.0;
        gy = ((int)r.p0 == 0) ? 0.0 : s;
    }
}

// psi and its gradient at (x, y)
__device__ __forceinline__ double uam_psi_grad(const UamEdge* __restrict__ edges, int e0, int e1, double x, double y,
                                               double e, double& gx, double& gy) {
    double psi = 1.0, sx = 0.0, sy = 0.0;
    for (int i = e0; i < e1; ++i) {
        const UamEdge r = uam_load_edge(edges + i);
        const double m = fmin(__dsub_rn(uam_h_exact(r, x, y), e), 0.0);
        psi = __dmul_rn(psi, __dmul_rn(m, m));
        if (psi == 0.0) { gx = 0.0; gy = 0.0; return 0.0; }
        double hx, hy;
        uam_grad_h(r, x, y, hx, hy);
        sx += 2.0 * hx / m;
        sy += 2.0 * hy / m;
    }
    gx = psi * sx;
    gy = psi * sy;
    return psi;
}

__global__ void __launch_bounds__(UAM_CTA_THREADS)
uam_k_grad_analytic(const double2* __restrict__ z, long long B, int N, UamParams prm, UamRegionRanges3 rr, UamShapeGrid sg,
                    const UamEdge* __restrict__ edges, const UamShape* __restrict__ shapes,
                    const double* __restrict__ psic, double* __restrict__ cost, double2* __restrict__ grad) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * UAM_WARPS_PER_CTA + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * UAM_WARPS_PER_CTA;
    const int W = N + 2;
    const bool len_smooth = (prm.flags & UAM_LENGTH_SMOOTH) != 0;
    const bool own_start = (prm.flags & UAM_OWN_START) != 0;
    const double dN = (double)N, dN1 = (double)(N + 1);
    for (long long path = warp0; path < B; path += nwarps) {
        const double2* zp = z + path * W;
        double pen_sum = 0.0, len_sum = 0.0;
        for (int j = lane; j < W; j += 32) {
            const double2 p = zp[j];
            // ---- penalty term and its gradient at z_j ----
            double P = 0.0, Gx = 0.0, Gy = 0.0;
            auto add_shape = [&](int s, double& tot, double& tx, double& ty) {
                const int4 meta = __ldg(reinterpret_cast<const int4*>(&shapes[s].e0));
                double gx, gy;
                const double psi = uam_psi_grad(edges, meta.x, meta.y, p.x, p.y, prm.e, gx, gy);
                const double pc = meta.w ? __ldg(psic + s) : 1.0;
                if (psi != 0.0 || pc == 0.0 || pc != pc) {
                    tot += psi / pc;
                    tx += gx / pc;
                    ty += gy / pc;
                }
            };
            const int cell = uam_shape_grid_cell(sg, p.x, p.y);
            if (cell < 0) {
                for (int r = 0; r < prm.n_regions; ++r) {
                    double tot = 0.0, tx = 0.0, ty = 0.0;
                    for (int s = rr.begin[r]; s < rr.begin[r + 1]; ++s) add_shape(s, tot, tx, ty);
                    P += prm.w[r] * tot;
                    Gx += prm.w[r] * tx;
                    Gy += prm.w[r] * ty;
                }
            } else {
                // shape grid: only the cell's candidates, region by region (the shapes left out have psi = 0 and grad = 0)
                int li = __ldg(sg.start + cell);
                const int l1 = __ldg(sg.start + cell + 1);
                while (li < l1 && __ldg(sg.items + li) < rr.begin[0]) ++li;          // hard obstacles are not in the cost
                while (li < l1) {
                    const int r = __ldg(&shapes[__ldg(sg.items + li)].region);
                    double tot = 0.0, tx = 0.0, ty = 0.0;
                    for (; li < l1 && __ldg(sg.items + li) < rr.begin[r + 1]; ++li) add_shape(__ldg(sg.items + li), tot, tx, ty);
                    P += prm.w[r] * tot;
                    Gx += prm.w[r] * tx;
                    Gy += prm.w[r] * ty;
                }
            }
            pen_sum += P / dN;
            Gx /= dN;
            Gy /= dN;
            // ---- length term: pairs (m_s, z_0), (z_0, z_1), ..., (z_{N-1}, z_N) ----
            double Lx = 0.0, Ly = 0.0;
            if (j <= N) {
                // pair ending at z_j: (z_{j-1}, z_j) for j >= 1, (m_s, z_0) for j == 0
                if (j >= 1 || !own_start) {
                    const double2 a = j >= 1 ? zp[j - 1] : make_double2(prm.ms_x, prm.ms_y);
                    const double dx = p.x - a.x, dy = p.y - a.y;
                    if (len_smooth) { Lx += 2.0 * dx; Ly += 2.0 * dy; }
                    else { const double n = sqrt(dx * dx + dy * dy); if (n > 0.0) { Lx += dx / n; Ly += dy / n; } }   // |d| = 0: zero subgradient
                }
                // pair starting at z_j: (z_j, z_{j+1}) for j <= N-1
                if (j <= N - 1) {
                    const double2 q = zp[j + 1];
                    const double dx = q.x - p.x, dy = q.y - p.y;
                    const double n = sqrt(dx * dx + dy * dy);
                    if (len_smooth) { Lx -= 2.0 * dx; Ly -= 2.0 * dy; len_sum += n * n; }
                    else { if (n > 0.0) { Lx -= dx / n; Ly -= dy / n; } len_sum += n; }
                }
                if (j == 0 && !own_start) {
                    const double dx = p.x - prm.ms_x, dy = p.y - prm.ms_y;
                    const double n = sqrt(dx * dx + dy * dy);
                    len_sum += len_smooth ? n * n : n;
                }
            }
            grad[path * W + j] = make_double2(dN1 * Lx + Gx, dN1 * Ly + Gy);
        }
        pen_sum = uam_warp_sum(pen_sum);
        len_sum = uam_warp_sum(len_sum);
        if (lane == 0 && cost) cost[path] = dN1 * len_sum + pen_sum;
    }
}

}  // namespace

extern "C" int uam_grad_paths_analytic(uam_ctx* ctx, const double* d_z, int64_t B, int N, const double* h_p, int n_p,
                                       int flags, double* d_cost, double* d_grad, void* stream) {
    if (!ctx) return UAM_ERR_INVALID;
    if (B < 0 || N < 1) return uam_fail(ctx, UAM_ERR_INVALID, "need B >= 0 and N >= 1 (got B=%lld N=%d)", (long long)B, N);
    if (B > 0 && (!d_z || !d_grad)) return uam_fail(ctx, UAM_ERR_INVALID, "NULL pointer");
    if (!(flags & UAM_PENALTY_SMOOTH)) return uam_fail(ctx, UAM_ERR_UNSUPPORTED, "gradient needs penalty_smooth (the non-smooth penalty is NaN in the reference)");
    if (!ctx->has_shapes) return uam_fail(ctx, UAM_ERR_STATE, "no shape table: call uam_map_set_shapes first");
    UamParams prm;
    UAM_TRY(uam_make_params(ctx, h_p, n_p, flags, &prm));
    if (prm.n_regions != ctx->n_regions)
        return uam_fail(ctx, UAM_ERR_INVALID, "p carries %d region weights, the map has %d regions", prm.n_regions, ctx->n_regions);
    UAM_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = uam_pick_stream(ctx, stream);
    UAM_TRY(uam_ensure_shape_norm(ctx, prm, st));
    if (B == 0) return UAM_OK;
    UamRegionRanges3 rr;
    for (int r = 0; r <= ctx->n_regions; ++r) rr.begin[r] = ctx->region_begin[r];
    const long long ctas = std::min<long long>((B + UAM_WARPS_PER_CTA - 1) / UAM_WARPS_PER_CTA, (long long)ctx->sm_count * 8);
    uam_k_grad_analytic<<<(unsigned)ctas, UAM_CTA_THREADS, 0, st>>>(reinterpret_cast<const double2*>(d_z), B, N, prm, rr, uam_pick_shape_grid(ctx, prm), ctx->d_edges,
                                                                    ctx->d_shapes, ctx->d_psic, d_cost, reinterpret_cast<double2*>(d_grad));
    UAM_CHECK_LAUNCH(ctx, "uam_k_grad_analytic");
    return UAM_OK;
}
