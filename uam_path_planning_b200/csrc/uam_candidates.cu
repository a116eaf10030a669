// Batched candidate generator: Solver.create_x_init (path_generation/solver.py:103-136) for B displacements at
// once, written straight into the scorer's path layout [start, x_1 .. x_N, goal] on the device -- so that a sweep of
// the arc family costs 8 bytes of upload per candidate instead of 16 (N + 2).
//   d == 0 : N interior points of linspace(start, goal, N + 2)
//   d != 0 : a = |goal - start| / 2, b = d a, alpha = atan2(start - goal), beta = 2 atan(2ab / (a^2 - b^2)),
//            rho = (a^2 + b^2) / (2b), t = linspace((pi - beta)/2, (pi + beta)/2, N + 2)[1:-1],
//            point = R(alpha) [rho cos t ; (b^2 - a^2)/(2b) + rho sin t] + (start + goal)/2
// uam_make_candidates is the general form: every candidate brings its own (start, goal, displacement) -- a batch of
// independent start/goal queries, 40 bytes of upload each -- and the N interior waypoints may be jittered by
// N(0, sigma^2) per coordinate from a counter-based generator (splitmix64 of (seed, path, waypoint) -> two uniforms ->
// Box-Muller), so a candidate is reproducible from its index alone, whatever the launch geometry or the number of GPUs.
// fp64 throughout, numpy's operation order; sin/cos/atan are CUDA's (<= 2 ulp from glibc's), so waypoints agree with
// the reference to ~1e-15 relative, not bit for bit.
#include <algorithm>

#include "uam_internal.cuh"

namespace {

__device__ __forceinline__ unsigned long long uam_splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
// two independent N(0,1) values for (seed, path, waypoint): u1 in (0, 1], u2 in [0, 1) from the top 53 bits of two hashes
__device__ __forceinline__ double2 uam_normal_pair(unsigned long long seed, unsigned long long path, int j, int Wp) {
    const unsigned long long ctr = path * (unsigned long long)Wp + (unsigned long long)j;
    const unsigned long long h1 = uam_splitmix64(seed ^ uam_splitmix64(2ull * ctr));
    const unsigned long long h2 = uam_splitmix64(seed ^ uam_splitmix64(2ull * ctr + 1ull));
    const double u1 = __dmul_rn((double)((h1 >> 11) + 1ull), 1.1102230246251565e-16);     // 2^-53
    const double u2 = __dmul_rn((double)(h2 >> 11), 1.1102230246251565e-16);
    const double r = sqrt(__dmul_rn(-2.0, log(u1)));
    const double a = __dmul_rn(6.283185307179586, u2);
    return make_double2(__dmul_rn(r, cos(a)), __dmul_rn(r, sin(a)));
}

// cand != nullptr: row b = {xs, ys, xg, yg, d}; else the ends are (sx, sy, gx, gy) for every path and d = disp[b]
__global__ void __launch_bounds__(UAM_CTA_THREADS)
uam_k_make_arc_paths(double sx, double sy, double gx, double gy, int N, const double* __restrict__ disp,
                     const double* __restrict__ cand, double sigma, unsigned long long seed, unsigned long long index0,
                     long long B, double2* __restrict__ z) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * UAM_WARPS_PER_CTA + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * UAM_WARPS_PER_CTA;
    const int Wp = N + 2;
    const double pi = 3.141592653589793;
    for (long long path = warp0; path < B; path += nwarps) {
        double d;
        if (cand) {
            const double* c = cand + 5 * path;
            sx = c[0]; sy = c[1]; gx = c[2]; gy = c[3]; d = c[4];
        } else {
            d = disp[path];
        }
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
                if (sigma != 0.0 && j > 0 && j < Wp - 1) {
                    const double2 n = uam_normal_pair(seed, index0 + (unsigned long long)path, j, Wp);
                    p.x = __dadd_rn(p.x, __dmul_rn(sigma, n.x));
                    p.y = __dadd_rn(p.y, __dmul_rn(sigma, n.y));
                }
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
                if (sigma != 0.0) {
                    const double2 n = uam_normal_pair(seed, index0 + (unsigned long long)path, j, Wp);
                    p.x = __dadd_rn(p.x, __dmul_rn(sigma, n.x));
                    p.y = __dadd_rn(p.y, __dmul_rn(sigma, n.y));
                }
            }
            zp[j] = p;
        }
    }
}

}  // namespace

int uam_make_candidates_launch(uam_ctx* ctx, const double* d_cand, const double* h_ends, const double* d_disp, int N, int64_t B,
                               double jitter_sigma, uint64_t seed, uint64_t index0, double* d_z, cudaStream_t st) {
    const long long ctas = std::min<long long>((B + UAM_WARPS_PER_CTA - 1) / UAM_WARPS_PER_CTA, (long long)ctx->sm_count * 8);
    const double e0 = h_ends ? h_ends[0] : 0.0, e1 = h_ends ? h_ends[1] : 0.0, e2 = h_ends ? h_ends[2] : 0.0, e3 = h_ends ? h_ends[3] : 0.0;
    uam_k_make_arc_paths<<<(unsigned)ctas, UAM_CTA_THREADS, 0, st>>>(e0, e1, e2, e3, N, d_disp, d_cand, jitter_sigma,
                                                                     (unsigned long long)seed, (unsigned long long)index0, B,
                                                                     reinterpret_cast<double2*>(d_z));
    UAM_CHECK_LAUNCH(ctx, "uam_k_make_arc_paths");
    return UAM_OK;
}

extern "C" int uam_make_arc_paths(uam_ctx* ctx, const double* h_ends, int N, const double* d_displacement, int64_t B,
                                  double* d_z, void* stream) {
    if (!ctx) return UAM_ERR_INVALID;
    if (!h_ends || N < 1 || B < 0 || (B > 0 && (!d_displacement || !d_z)))
        return uam_fail(ctx, UAM_ERR_INVALID, "bad argument to uam_make_arc_paths");
    if (B == 0) return UAM_OK;
    UAM_CUDA(ctx, cudaSetDevice(ctx->device));
    return uam_make_candidates_launch(ctx, nullptr, h_ends, d_displacement, N, B, 0.0, 0, 0, d_z, uam_pick_stream(ctx, stream));
}

extern "C" int uam_make_candidates(uam_ctx* ctx, const double* d_cand, int64_t B, int N, double jitter_sigma, uint64_t seed,
                                   uint64_t index0, double* d_z, void* stream) {
    if (!ctx) return UAM_ERR_INVALID;
    if (N < 1 || B < 0 || (B > 0 && (!d_cand || !d_z)) || !(jitter_sigma >= 0.0))
        return uam_fail(ctx, UAM_ERR_INVALID, "bad argument to uam_make_candidates");
    if (B == 0) return UAM_OK;
    UAM_CUDA(ctx, cudaSetDevice(ctx->device));
    return uam_make_candidates_launch(ctx, d_cand, nullptr, nullptr, N, B, jitter_sigma, seed, index0, d_z, uam_pick_stream(ctx, stream));
}
