// Batched candidate generator: Solver.create_x_init (path_generation/solver.py:103-136) for B displacements at
// once, written straight into the scorer's path layout [start, x_1 .. x_N, goal] on the device -- so that a sweep of
// the arc family costs 8 bytes of upload per candidate instead of 16 (N + 2).
//   d == 0 : N interior points of linspace(start, goal, N + 2)
//   d != 0 : a = |goal - start| / 2, b = d a, alpha = atan2(start - goal), beta = 2 atan(2ab / (a^2 - b^2)),
//            rho = (a^2 + b^2) / (2b), t = linspace((pi - beta)/2, (pi + beta)/2, N + 2)[1:-1],
//            point = R(alpha) [rho cos t ; (b^2 - a^2)/(2b) + rho sin t] + (start + goal)/2
// fp64 throughout, numpy's operation order; sin/cos/atan are CUDA's (<= 2 ulp from glibc's), so waypoints agree with
// the reference to ~1e-15 relative, not bit for bit.
#include <algorithm>

#include "uam_internal.cuh"

namespace {

__global__ void __launch_bounds__(UAM_CTA_THREADS)
uam_k_make_arc_paths(double sx, double sy, double gx, double gy, int N, const double* __restrict__ disp, long long B,
                     double2* __restrict__ z) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * UAM_WARPS_PER_CTA + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * UAM_WARPS_PER_CTA;
    const int Wp = N + 2;
    const double pi = 3.141592653589793;
    for (long long path = warp0; path < B; path += nwarps) {
        const double d = disp[path];
        double2* zp = z + path * Wp;
        if (d == 0.0) {
            // np.linspace(a, b, n): a + arange(n) * ((b - a) / (n - 1)), last element set to b
            const double stepx = __ddiv_rn(__dsub_rn(gx, sx), (double)(Wp - 1));
            const double stepy = __ddiv_rn(__dsub_rn(gy, sy), (double)(Wp - 1));
            for (int j = lane; j < Wp; j += 32) {
                double2 p;
                p.x = __dadd_rn(sx, __dmul_rn((double)j, stepx));
                p.y = __dadd_rn(sy, __dmul_rn((double)j, stepy));
                if (j == 0) p = make_double2(sx, sy);
                if (j == Wp - 1) p = make_double2(gx, gy);
                zp[j] = p;
            }
            continue;
        }
        const double vx = __dsub_rn(sx, gx), vy = __dsub_rn(sy, gy);
        const double a = __ddiv_rn(sqrt(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy))), 2.0);
        const double b = __dmul_rn(d, a);
        const double alpha = atan2(vy, vx);
        const double ca = cos(alpha), sa = sin(alpha);
        const double a2 = __dmul_rn(a, a), b2 = __dmul_rn(b, b);
        const double beta = __dmul_rn(2.0, atan(__ddiv_rn(__dmul_rn(__dmul_rn(2.0, a), b), __dsub_rn(a2, b2))));
        const double radius = __ddiv_rn(__dadd_rn(a2, b2), __dmul_rn(2.0, b));
        const double off = __ddiv_rn(__dsub_rn(b2, a2), __dmul_rn(2.0, b));
        const double t0 = __ddiv_rn(__dsub_rn(pi, beta), 2.0), t1 = __ddiv_rn(__dadd_rn(pi, beta), 2.0);
        const double step = __ddiv_rn(__dsub_rn(t1, t0), (double)(Wp - 1));
        const double cx = __ddiv_rn(__dadd_rn(gx, sx), 2.0), cy = __ddiv_rn(__dadd_rn(gy, sy), 2.0);
        for (int j = lane; j < Wp; j += 32) {
            double2 p;
            if (j == 0) {
                p = make_double2(sx, sy);
            } else if (j == Wp - 1) {
                p = make_double2(gx, gy);
            } else {
                const double t = __dadd_rn(t0, __dmul_rn((double)j, step));
                const double ex = __dmul_rn(radius, cos(t));
                const double ey = __dadd_rn(off, __dmul_rn(radius, sin(t)));
                // R @ [ex; ey] + C
                p.x = __dadd_rn(__dadd_rn(__dmul_rn(ca, ex), __dmul_rn(-sa, ey)), cx);
                p.y = __dadd_rn(__dadd_rn(__dmul_rn(sa, ex), __dmul_rn(ca, ey)), cy);
            }
            zp[j] = p;
        }
    }
}

}  // namespace

extern "C" int uam_make_arc_paths(uam_ctx* ctx, const double* h_ends, int N, const double* d_displacement, int64_t B,
                                  double* d_z, void* stream) {
    if (!ctx) return UAM_ERR_INVALID;
    if (!h_ends || N < 1 || B < 0 || (B > 0 && (!d_displacement || !d_z)))
        return uam_fail(ctx, UAM_ERR_INVALID, "bad argument to uam_make_arc_paths");
    if (B == 0) return UAM_OK;
    UAM_CUDA(ctx, cudaSetDevice(ctx->device));
    const long long ctas = std::min<long long>((B + UAM_WARPS_PER_CTA - 1) / UAM_WARPS_PER_CTA, (long long)ctx->sm_count * 8);
    uam_k_make_arc_paths<<<(unsigned)ctas, UAM_CTA_THREADS, 0, uam_pick_stream(ctx, stream)>>>(
        h_ends[0], h_ends[1], h_ends[2], h_ends[3], N, d_displacement, B, reinterpret_cast<double2*>(d_z));
    UAM_CHECK_LAUNCH(ctx, "uam_k_make_arc_paths");
    return UAM_OK;
}
