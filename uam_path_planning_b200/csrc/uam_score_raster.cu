// Raster path scorer: the reference's cost functional (problem.py:38-44,130-146) with the analytic penalty
// replaced by a bilinear lookup in the cost rasters, and Map.collides replaced by nearest-cell occupancy.
//
// HBM layout: one texel per cell, layers interleaved with the occupancy flag (float2 for L = 1, float4 for
// L = 2..3), row-major (H, W): one bilinear tap = two 2-texel runs (2 x 16 B or 2 x 32 B) instead of
// 4 x (L + 1) scattered words.
//
// One warp per candidate path.
//   waypoint mode  (samples_per_cell == 0): lanes stride the N+2 waypoints -- the reference's sampling.
//   integral mode  (samples_per_cell  > 0): every segment k gets S_k = max(1, ceil(|dz_k|_cells * spc))
//       left-endpoint samples; the path's samples are flattened and lanes stride the flat sample index, so 32
//       consecutive samples of one polyline (a compact footprint in the raster) are fetched together.
// World -> pixel coordinates, S_k and the sample positions are fp64 in a fixed operation order (identical bits
// to the oracle: same cells, same occupancy lookups, same sample counts); bilinear weights, texel arithmetic and
// the penalty sum are fp32; the length term is fp64.  Per-path sums finish with warp-shuffle reductions.
#include <algorithm>

#include "uam_internal.cuh"

#define UAM_MAX_SAMPLES_PER_SEGMENT 1048576.0   // cap on S_k (a segment never spans more cells than this)

namespace {

struct UamRasterParams {
    double x0, dx, y0, dy;
    double ms_x, ms_y;
    double spc;
    int H, W;
    int tiles_x;          // tiled layout: tiles per tile-row
    float w0, w1, w2;
    int flags;
};

template <int TF> struct UamTexel;
template <> struct UamTexel<2> { typedef float2 T; };
template <> struct UamTexel<4> { typedef float4 T; };

// Texel address.  LAYOUT 0: row-major (H, W).  LAYOUT 1: tiled so that one 128-byte line is a compact 2-D block
// and every 32-byte sector a 2 x 1 (float4) / 2 x 2 (float2) block -- a polyline crossing the raster in any
// direction then touches ~1/3 fewer lines per warp instruction than with 8-texel-wide row-major lines, and the
// 64-byte DRAM fetch granule is a 2 x 2 (float4) / 4 x 2 (float2) block instead of a 4 x 1 / 8 x 1 strip.
//   float4: line = 4 wide x 2 tall:  ((i>>1) * tiles_x + (j>>2)) * 8  + ((j>>1)&1)*4 + (i&1)*2 + (j&1)
//   float2: line = 4 wide x 4 tall:  ((i>>2) * tiles_x + (j>>2)) * 16 + ((i>>1)&1)*8 + ((j>>1)&1)*4 + (i&1)*2 + (j&1)
template <int TF, int LAYOUT>
__device__ __forceinline__ size_t uam_tex_index(int i, int j, int W, int tiles_x) {
    if (LAYOUT == 0) return (size_t)i * W + j;
    if (TF == 4) return ((size_t)(i >> 1) * tiles_x + (j >> 2)) * 8 + (((j >> 1) & 1) << 2) + ((i & 1) << 1) + (j & 1);
    return ((size_t)(i >> 2) * tiles_x + (j >> 2)) * 16 + (((i >> 1) & 1) << 3) + (((j >> 1) & 1) << 2) + ((i & 1) << 1) + (j & 1);
}

// pixel coordinate of a world coordinate: (x - x0)/dx - 1/2 with a true division (oracle: pixel_coords)
__device__ __forceinline__ double uam_pix(double x, double x0, double dx) {
    return __dsub_rn(__ddiv_rn(__dsub_rn(x, x0), dx), 0.5);
}

__device__ __forceinline__ double uam_norm2r(double dx, double dy) {
    return sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
}

__device__ __forceinline__ float uam_lerp2(float t00, float t01, float t10, float t11, float fx, float fy) {
    const float top = t00 + fx * (t01 - t00);
    const float bot = t10 + fx * (t11 - t10);
    return top + fy * (bot - top);
}

// clamp pixel coordinates, split into cell + fp32 fraction (identical bits to the oracle's sample_uv)
__device__ __forceinline__ void uam_cell_frac(const UamRasterParams& rp, double u, double v, int& i0, int& j0, float& fx,
                                              float& fy) {
    u = fmin(fmax(u, 0.0), (double)(rp.W - 1));
    v = fmin(fmax(v, 0.0), (double)(rp.H - 1));
    j0 = min((int)u, rp.W - 2);
    i0 = min((int)v, rp.H - 2);
    fx = (float)__dsub_rn(u, (double)j0);
    fy = (float)__dsub_rn(v, (double)i0);
}

// Weighted bilinear penalty + nearest-cell occupancy at pixel coordinates (u, v): one lane fetches all 4 texels.
template <int TF, int LAYOUT>
__device__ __forceinline__ void uam_sample(const typename UamTexel<TF>::T* __restrict__ tex, const UamRasterParams& rp,
                                           double u, double v, float& pen, bool& occ) {
    int i0, j0;
    float fx, fy;
    uam_cell_frac(rp, u, v, i0, j0, fx, fy);
    const bool right = fx >= 0.5f, down = fy >= 0.5f;
    if constexpr (TF == 2) {
        const float2 a = __ldg(reinterpret_cast<const float2*>(tex) + uam_tex_index<TF, LAYOUT>(i0, j0, rp.W, rp.tiles_x));
        const float2 b = __ldg(reinterpret_cast<const float2*>(tex) + uam_tex_index<TF, LAYOUT>(i0, j0 + 1, rp.W, rp.tiles_x));
        const float2 c = __ldg(reinterpret_cast<const float2*>(tex) + uam_tex_index<TF, LAYOUT>(i0 + 1, j0, rp.W, rp.tiles_x));
        const float2 d = __ldg(reinterpret_cast<const float2*>(tex) + uam_tex_index<TF, LAYOUT>(i0 + 1, j0 + 1, rp.W, rp.tiles_x));
        pen = rp.w0 * uam_lerp2(a.x, b.x, c.x, d.x, fx, fy);
        const float o = down ? (right ? d.y : c.y) : (right ? b.y : a.y);
        occ = o != 0.0f;
    } else {
        const float4* t4 = reinterpret_cast<const float4*>(tex);
        const float4 a = __ldg(t4 + uam_tex_index<TF, LAYOUT>(i0, j0, rp.W, rp.tiles_x));
        const float4 b = __ldg(t4 + uam_tex_index<TF, LAYOUT>(i0, j0 + 1, rp.W, rp.tiles_x));
        const float4 c = __ldg(t4 + uam_tex_index<TF, LAYOUT>(i0 + 1, j0, rp.W, rp.tiles_x));
        const float4 d = __ldg(t4 + uam_tex_index<TF, LAYOUT>(i0 + 1, j0 + 1, rp.W, rp.tiles_x));
        pen = rp.w0 * uam_lerp2(a.x, b.x, c.x, d.x, fx, fy) + rp.w1 * uam_lerp2(a.y, b.y, c.y, d.y, fx, fy) +
              rp.w2 * uam_lerp2(a.z, b.z, c.z, d.z, fx, fy);
        const float o = down ? (right ? d.w : c.w) : (right ? b.w : a.w);
        occ = o != 0.0f;
    }
}

// Lane-pair form: the two lanes of a pair work on the SAME sample; lane `side` (0 = left, 1 = right) fetches the
// texel column j0 + side (rows i0 and i0 + 1), lerps it in y, and the pair exchanges the column results with one
// shuffle per layer.  The two lanes' loads of a row sit in the same 32-byte sector / 128-byte line, so each warp
// load instruction touches half as many lines as when every lane fetches its own 2 x 2 footprint.
// Must be called by all 32 lanes (shuffles); `active` masks lanes past the end.  pen is valid on both lanes.
template <int TF, int LAYOUT>
__device__ __forceinline__ void uam_sample_pair(const typename UamTexel<TF>::T* __restrict__ tex,
                                                const UamRasterParams& rp, double u, double v, int side, bool active,
                                                float& pen, bool& occ) {
    int i0 = 0, j0 = 0;
    float fx = 0.0f, fy = 0.0f;
    float c0 = 0.0f, c1 = 0.0f, c2 = 0.0f, o = 0.0f;
    if (active) {
        uam_cell_frac(rp, u, v, i0, j0, fx, fy);
        const bool down = fy >= 0.5f;
        if constexpr (TF == 2) {
            const float2 a = __ldg(reinterpret_cast<const float2*>(tex) + uam_tex_index<TF, LAYOUT>(i0, j0 + side, rp.W, rp.tiles_x));
            const float2 b = __ldg(reinterpret_cast<const float2*>(tex) + uam_tex_index<TF, LAYOUT>(i0 + 1, j0 + side, rp.W, rp.tiles_x));
            c0 = a.x + fy * (b.x - a.x);
            o = down ? b.y : a.y;
        } else {
            const float4* t4 = reinterpret_cast<const float4*>(tex);
            const float4 a = __ldg(t4 + uam_tex_index<TF, LAYOUT>(i0, j0 + side, rp.W, rp.tiles_x));
            const float4 b = __ldg(t4 + uam_tex_index<TF, LAYOUT>(i0 + 1, j0 + side, rp.W, rp.tiles_x));
            c0 = a.x + fy * (b.x - a.x);
            c1 = a.y + fy * (b.y - a.y);
            c2 = a.z + fy * (b.z - a.z);
            o = down ? b.w : a.w;
        }
    }
    const float p0 = __shfl_xor_sync(0xffffffffu, c0, 1);
    // after the exchange: (left, right) = side ? (p, c) : (c, p)
    float val = rp.w0 * (side ? p0 + fx * (c0 - p0) : c0 + fx * (p0 - c0));
    if constexpr (TF == 4) {
        const float p1 = __shfl_xor_sync(0xffffffffu, c1, 1);
        const float p2 = __shfl_xor_sync(0xffffffffu, c2, 1);
        val += rp.w1 * (side ? p1 + fx * (c1 - p1) : c1 + fx * (p1 - c1));
        val += rp.w2 * (side ? p2 + fx * (c2 - p2) : c2 + fx * (p2 - c2));
    }
    pen = val;
    // nearest cell column = j0 + (fx >= 0.5): only the lane holding that column reports occupancy
    occ = active && ((fx >= 0.5f) == (side != 0)) && (o != 0.0f);
}

// ---- shared pieces of the path kernels -----------------------------------------------------------------------
// length term of one waypoint j (reference quirk: segments 0..N-1 only + |z_0 - map.x_start|)
__device__ __forceinline__ double uam_len_term(const double2* __restrict__ zp, int j, int N, const double2 p,
                                               const UamRasterParams& rp) {
    const bool len_smooth = (rp.flags & UAM_LENGTH_SMOOTH) != 0;
    double acc = 0.0;
    if (j < N) {
        const double2 q = zp[j + 1];
        const double d = uam_norm2r(__dsub_rn(q.x, p.x), __dsub_rn(q.y, p.y));
        acc += len_smooth ? __dmul_rn(d, d) : d;
    }
    if (j == 0 && !(rp.flags & UAM_OWN_START)) {
        const double d = uam_norm2r(__dsub_rn(p.x, rp.ms_x), __dsub_rn(p.y, rp.ms_y));
        acc += len_smooth ? __dmul_rn(d, d) : d;
    }
    return acc;
}

// ---- waypoint mode ----------------------------------------------------------------------------------------
template <int TF, int LAYOUT>
__global__ void __launch_bounds__(UAM_CTA_THREADS)
uam_k_score_raster_wp(const double2* __restrict__ z, long long B, int Wp, UamRasterParams rp,
                      const typename UamTexel<TF>::T* __restrict__ tex, float* __restrict__ cost,
                      uint8_t* __restrict__ collide, long long* __restrict__ nsamp) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * UAM_WARPS_PER_CTA + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * UAM_WARPS_PER_CTA;
    const int N = Wp - 2;
    for (long long path = warp0; path < B; path += nwarps) {
        const double2* zp = z + path * Wp;
        float pen_sum = 0.0f;
        double len_sum = 0.0;
        bool col = false;
        for (int j = lane; j < Wp; j += 32) {
            const double2 p = zp[j];
            float pen;
            bool occ;
            uam_sample<TF, LAYOUT>(tex, rp, uam_pix(p.x, rp.x0, rp.dx), uam_pix(p.y, rp.y0, rp.dy), pen, occ);
            pen_sum += pen;
            col = col || occ;
            len_sum += uam_len_term(zp, j, N, p, rp);
        }
        pen_sum = uam_warp_sum(pen_sum);
        len_sum = uam_warp_sum(len_sum);
        col = __any_sync(0xffffffffu, col);
        if (lane == 0) {
            if (cost) cost[path] = (float)((double)(N + 1) * len_sum + (double)pen_sum / (double)N);
            if (collide) collide[path] = col ? 1 : 0;
            if (nsamp) nsamp[path] = Wp;
        }
    }
}

// ---- integral mode ----------------------------------------------------------------------------------------
// Per-warp shared memory: U[Wp] V[Wp] SU[Wp] SV[Wp] (double), P[Wp+1] (long long), IS[Wp] (float).
__host__ __device__ inline size_t uam_int_warp_smem(int Wp) {
    size_t b = (size_t)Wp * 8 * 4 + (size_t)(Wp + 1) * 8 + (size_t)Wp * 4;
    return (b + 15) & ~(size_t)15;
}

// PAIR = 0: one lane per sample (4 texel loads per lane).  PAIR = 1: two lanes per sample (uam_sample_pair).
template <int TF, int LAYOUT, int PAIR>
__global__ void __launch_bounds__(UAM_CTA_THREADS)
uam_k_score_raster_int(const double2* __restrict__ z, long long B, int Wp, UamRasterParams rp,
                       const typename UamTexel<TF>::T* __restrict__ tex, float* __restrict__ cost,
                       uint8_t* __restrict__ collide, long long* __restrict__ nsamp) {
    extern __shared__ __align__(16) unsigned char uam_smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wpc = blockDim.x >> 5;
    unsigned char* base = uam_smem + (size_t)warp * uam_int_warp_smem(Wp);
    double* sU = reinterpret_cast<double*>(base);
    double* sV = sU + Wp;
    double* sSU = sV + Wp;
    double* sSV = sSU + Wp;
    long long* sP = reinterpret_cast<long long*>(sSV + Wp);
    float* sIS = reinterpret_cast<float*>(sP + Wp + 1);

    const long long warp0 = (long long)blockIdx.x * wpc + warp;
    const long long nwarps = (long long)gridDim.x * wpc;
    const int N = Wp - 2;

    for (long long path = warp0; path < B; path += nwarps) {
        const double2* zp = z + path * Wp;
        double len_sum = 0.0;
        // A: pixel coordinates of the waypoints
        for (int j = lane; j < Wp; j += 32) {
            const double2 p = zp[j];
            sU[j] = uam_pix(p.x, rp.x0, rp.dx);
            sV[j] = uam_pix(p.y, rp.y0, rp.dy);
            len_sum += uam_len_term(zp, j, N, p, rp);
        }
        __syncwarp();
        // B: per-segment sample count, step and exclusive prefix; pseudo-segment Wp-1 = the goal waypoint
        long long carry = 0;
        for (int b0 = 0; b0 < Wp; b0 += 32) {
            const int k = b0 + lane;
            long long S = 0;
            if (k < Wp - 1) {
                const double dU = __dsub_rn(sU[k + 1], sU[k]), dV = __dsub_rn(sV[k + 1], sV[k]);
                double Sd = fmax(1.0, ceil(__dmul_rn(uam_norm2r(dU, dV), rp.spc)));
                if (!(Sd <= UAM_MAX_SAMPLES_PER_SEGMENT)) Sd = UAM_MAX_SAMPLES_PER_SEGMENT;
                S = (long long)Sd;
                sSU[k] = __ddiv_rn(dU, Sd);
                sSV[k] = __ddiv_rn(dV, Sd);
                sIS[k] = (float)(1.0 / Sd);
            } else if (k == Wp - 1) {
                S = 1;
                sSU[k] = 0.0;
                sSV[k] = 0.0;
                sIS[k] = 1.0f;
            }
            long long incl = S;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (k < Wp) sP[k] = carry + incl - S;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) sP[Wp] = carry;
        __syncwarp();
        // C: flat sample loop
        const long long T = carry;
        float acc = 0.0f;
        bool col = false;
        int k = 0;
        long long p0 = 0, p1 = sP[1];
        double kU = sU[0], kV = sV[0], kSU = sSU[0], kSV = sSV[0];
        float kIS = sIS[0];
        if (PAIR) {
            const int side = lane & 1;
            for (long long tb = 0; tb < T; tb += 16) {
                const long long t = tb + (lane >> 1);
                const bool active = t < T;
                double u = 0.0, v = 0.0;
                if (active) {
                    if (t >= p1) {
                        do { ++k; p1 = sP[k + 1]; } while (t >= p1);
                        p0 = sP[k];
                        kU = sU[k]; kV = sV[k]; kSU = sSU[k]; kSV = sSV[k]; kIS = sIS[k];
                    }
                    const double s = (double)(t - p0);
                    u = __dadd_rn(kU, __dmul_rn(s, kSU));
                    v = __dadd_rn(kV, __dmul_rn(s, kSV));
                }
                float pen;
                bool occ;
                uam_sample_pair<TF, LAYOUT>(tex, rp, u, v, side, active, pen, occ);
                if (active && side == 0) acc += pen * kIS;
                col = col || occ;
            }
        } else {
            for (long long t = lane; t < T; t += 32) {
                if (t >= p1) {
                    do { ++k; p1 = sP[k + 1]; } while (t >= p1);
                    p0 = sP[k];
                    kU = sU[k]; kV = sV[k]; kSU = sSU[k]; kSV = sSV[k]; kIS = sIS[k];
                }
                const double s = (double)(t - p0);
                const double u = __dadd_rn(kU, __dmul_rn(s, kSU));
                const double v = __dadd_rn(kV, __dmul_rn(s, kSV));
                float pen;
                bool occ;
                uam_sample<TF, LAYOUT>(tex, rp, u, v, pen, occ);
                acc += pen * kIS;
                col = col || occ;
            }
        }
        acc = uam_warp_sum(acc);
        len_sum = uam_warp_sum(len_sum);
        col = __any_sync(0xffffffffu, col);
        if (lane == 0) {
            if (cost) cost[path] = (float)((double)(N + 1) * len_sum + (double)acc / (double)N);
            if (collide) collide[path] = col ? 1 : 0;
            if (nsamp) nsamp[path] = T;
        }
        __syncwarp();
    }
}

int uam_raster_prepare(uam_ctx* ctx, int64_t B, int N, const double* h_p, int n_p, int flags, double spc,
                       UamRasterParams* rp) {
    if (B < 0 || N < 1) return uam_fail(ctx, UAM_ERR_INVALID, "need B >= 0 and N >= 1 (got B=%lld N=%d)", (long long)B, N);
    if (!ctx->has_raster) return uam_fail(ctx, UAM_ERR_STATE, "no raster: call uam_map_set_raster first");
    if (!(spc >= 0.0) || spc > 64.0) return uam_fail(ctx, UAM_ERR_INVALID, "samples_per_cell must be in [0, 64]");
    UamParams prm;
    UAM_TRY(uam_make_params(ctx, h_p, n_p, flags, &prm));
    if (prm.n_regions != ctx->geo.L)
        return uam_fail(ctx, UAM_ERR_INVALID, "p carries %d layer weights, the raster has %d layers", prm.n_regions, ctx->geo.L);
    rp->x0 = ctx->geo.x0; rp->dx = ctx->geo.dx; rp->y0 = ctx->geo.y0; rp->dy = ctx->geo.dy;
    rp->ms_x = prm.ms_x; rp->ms_y = prm.ms_y;
    rp->spc = spc;
    rp->H = ctx->geo.H; rp->W = ctx->geo.W;
    rp->tiles_x = ctx->geo.tiles_x;
    rp->w0 = (float)prm.w[0];
    rp->w1 = ctx->geo.L > 1 ? (float)prm.w[1] : 0.0f;
    rp->w2 = ctx->geo.L > 2 ? (float)prm.w[2] : 0.0f;
    rp->flags = flags;
    return UAM_OK;
}

template <int TF, int LAYOUT>
int uam_raster_launch_t(uam_ctx* ctx, const double2* z, int64_t B, int Wp, const UamRasterParams& rp, float* d_cost,
                        uint8_t* d_collide, long long* d_nsamp, cudaStream_t st) {
    typedef typename UamTexel<TF>::T T;
    const T* tex = (const T*)ctx->d_tex;
    if (rp.spc == 0.0) {
        const long long ctas = std::min<long long>((B + UAM_WARPS_PER_CTA - 1) / UAM_WARPS_PER_CTA, (long long)ctx->sm_count * 16);
        uam_k_score_raster_wp<TF, LAYOUT><<<(unsigned)ctas, UAM_CTA_THREADS, 0, st>>>(z, B, Wp, rp, tex, d_cost, d_collide, d_nsamp);
        UAM_CHECK_LAUNCH(ctx, "uam_k_score_raster_wp");
        return UAM_OK;
    }
    const size_t per_warp = uam_int_warp_smem(Wp);
    const size_t budget = 200 * 1024;
    if (per_warp > budget) return uam_fail(ctx, UAM_ERR_UNSUPPORTED, "N = %d waypoints per path is too many for integral mode", Wp - 2);
    const int wpc = (int)std::max<size_t>(1, std::min<size_t>(UAM_WARPS_PER_CTA, budget / per_warp));
    const size_t smem = per_warp * wpc;
    const long long ctas = std::min<long long>((B + wpc - 1) / wpc, (long long)ctx->sm_count * 16);
    if (ctx->int_variant == 0) {
        if (smem > 48 * 1024) UAM_CUDA(ctx, cudaFuncSetAttribute(uam_k_score_raster_int<TF, LAYOUT, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        uam_k_score_raster_int<TF, LAYOUT, 0><<<(unsigned)ctas, wpc * 32, smem, st>>>(z, B, Wp, rp, tex, d_cost, d_collide, d_nsamp);
    } else {
        if (smem > 48 * 1024) UAM_CUDA(ctx, cudaFuncSetAttribute(uam_k_score_raster_int<TF, LAYOUT, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        uam_k_score_raster_int<TF, LAYOUT, 1><<<(unsigned)ctas, wpc * 32, smem, st>>>(z, B, Wp, rp, tex, d_cost, d_collide, d_nsamp);
    }
    UAM_CHECK_LAUNCH(ctx, "uam_k_score_raster_int");
    return UAM_OK;
}

int uam_raster_launch(uam_ctx* ctx, const double* d_z, int64_t B, int N, const UamRasterParams& rp, float* d_cost,
                      uint8_t* d_collide, long long* d_nsamp, cudaStream_t st) {
    const int Wp = N + 2;
    const double2* z = reinterpret_cast<const double2*>(d_z);
    const int tf = ctx->geo.texel_floats, lay = ctx->geo.layout;
    if (tf == 2) return lay ? uam_raster_launch_t<2, 1>(ctx, z, B, Wp, rp, d_cost, d_collide, d_nsamp, st)
                            : uam_raster_launch_t<2, 0>(ctx, z, B, Wp, rp, d_cost, d_collide, d_nsamp, st);
    return lay ? uam_raster_launch_t<4, 1>(ctx, z, B, Wp, rp, d_cost, d_collide, d_nsamp, st)
               : uam_raster_launch_t<4, 0>(ctx, z, B, Wp, rp, d_cost, d_collide, d_nsamp, st);
}

// ---- best candidate: min over b of (float bits of cost << 32 | global index) ----------------------------------
template <typename CT>
__global__ void __launch_bounds__(256)
uam_k_best(const CT* __restrict__ cost, long long B, unsigned long long offset, unsigned long long* __restrict__ key) {
    unsigned long long best = ~0ull;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += stride) {
        const float c = (float)cost[b];
        const unsigned long long k = ((unsigned long long)__float_as_uint(c) << 32) | ((offset + (unsigned long long)b) & 0xffffffffull);
        best = k < best ? k : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o);
        best = t < best ? t : best;
    }
    __shared__ unsigned long long s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < (int)(blockDim.x >> 5); ++i) best = s[i] < best ? s[i] : best;
        if (best != ~0ull) atomicMin(key, best);
    }
}

__global__ void uam_k_set_u64(unsigned long long* p, unsigned long long v) { *p = v; }

}  // namespace

extern "C" int uam_score_paths_raster(uam_ctx* ctx, const double* d_z, int64_t B, int N, const double* h_p, int n_p,
                                      int flags, double samples_per_cell, float* d_cost, uint8_t* d_collide,
                                      int64_t* d_nsamples, void* stream) {
    if (!ctx) return UAM_ERR_INVALID;
    UamRasterParams rp;
    UAM_TRY(uam_raster_prepare(ctx, B, N, h_p, n_p, flags, samples_per_cell, &rp));
    if (B == 0) return UAM_OK;
    if (!d_z) return uam_fail(ctx, UAM_ERR_INVALID, "paths pointer is NULL");
    UAM_CUDA(ctx, cudaSetDevice(ctx->device));
    return uam_raster_launch(ctx, d_z, B, N, rp, d_cost, d_collide, (long long*)d_nsamples, uam_pick_stream(ctx, stream));
}

// Host buffers in, host buffers out: the batch is cut into chunks that flow through UAM_HOST_PIPE_DEPTH
// streams (H2D copy of chunk c+1 overlaps the scoring of chunk c and the D2H copy of chunk c-1).
extern "C" int uam_score_paths_raster_host(uam_ctx* ctx, const double* h_z, int64_t B, int N, const double* h_p,
                                           int n_p, int flags, double samples_per_cell, float* h_cost,
                                           uint8_t* h_collide) {
    if (!ctx) return UAM_ERR_INVALID;
    UamRasterParams rp;
    UAM_TRY(uam_raster_prepare(ctx, B, N, h_p, n_p, flags, samples_per_cell, &rp));
    if (B == 0) return UAM_OK;
    if (!h_z) return uam_fail(ctx, UAM_ERR_INVALID, "paths pointer is NULL");
    UAM_CUDA(ctx, cudaSetDevice(ctx->device));
    UAM_CUDA(ctx, cudaDeviceSynchronize());
    const size_t row = (size_t)2 * (N + 2) * sizeof(double);
    const int64_t chunk = std::max<int64_t>(1024, std::min<int64_t>((B + 2 * UAM_HOST_PIPE_DEPTH - 1) / (2 * UAM_HOST_PIPE_DEPTH),
                                                                      (int64_t)((64u << 20) / row)));
    int c = 0;
    for (int64_t b0 = 0; b0 < B; b0 += chunk, ++c) {
        const int s = c % UAM_HOST_PIPE_DEPTH;
        const int64_t nb = std::min(chunk, B - b0);
        cudaStream_t st = ctx->pipe_stream[s];
        UAM_TRY(uam_reserve(ctx, &ctx->d_stage_in[s], &ctx->stage_in_bytes[s], (size_t)chunk * row));
        UAM_TRY(uam_reserve(ctx, &ctx->d_stage_out[s], &ctx->stage_out_bytes[s], (size_t)chunk * 8));
        float* d_cost = (float*)ctx->d_stage_out[s];
        uint8_t* d_col = (uint8_t*)(d_cost + chunk);
        UAM_CUDA(ctx, cudaMemcpyAsync(ctx->d_stage_in[s], (const char*)h_z + (size_t)b0 * row, (size_t)nb * row,
                                      cudaMemcpyHostToDevice, st));
        UAM_TRY(uam_raster_launch(ctx, (const double*)ctx->d_stage_in[s], nb, N, rp, d_cost, d_col, nullptr, st));
        if (h_cost) UAM_CUDA(ctx, cudaMemcpyAsync(h_cost + b0, d_cost, (size_t)nb * 4, cudaMemcpyDeviceToHost, st));
        if (h_collide) UAM_CUDA(ctx, cudaMemcpyAsync(h_collide + b0, d_col, (size_t)nb, cudaMemcpyDeviceToHost, st));
    }
    for (int s = 0; s < UAM_HOST_PIPE_DEPTH; ++s) UAM_CUDA(ctx, cudaStreamSynchronize(ctx->pipe_stream[s]));
    return UAM_OK;
}

extern "C" int uam_best(uam_ctx* ctx, const void* d_cost, int cost_is_f64, int64_t B, int64_t global_offset,
                        uint64_t* d_key, int reset, void* stream) {
    if (!ctx) return UAM_ERR_INVALID;
    if (!d_key || B < 0 || (B > 0 && !d_cost)) return uam_fail(ctx, UAM_ERR_INVALID, "bad argument to uam_best");
    if (global_offset < 0 || global_offset + B > 0xffffffffll)
        return uam_fail(ctx, UAM_ERR_UNSUPPORTED, "global path index must fit 32 bits");
    UAM_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = uam_pick_stream(ctx, stream);
    if (reset) {
        uam_k_set_u64<<<1, 1, 0, st>>>((unsigned long long*)d_key, ~0ull);
        UAM_CHECK_LAUNCH(ctx, "uam_k_set_u64");
    }
    if (B == 0) return UAM_OK;
    const long long ctas = std::min<long long>((B + 255) / 256, (long long)ctx->sm_count * 4);
    if (cost_is_f64)
        uam_k_best<double><<<(unsigned)ctas, 256, 0, st>>>((const double*)d_cost, B, (unsigned long long)global_offset, (unsigned long long*)d_key);
    else
        uam_k_best<float><<<(unsigned)ctas, 256, 0, st>>>((const float*)d_cost, B, (unsigned long long)global_offset, (unsigned long long*)d_key);
    UAM_CHECK_LAUNCH(ctx, "uam_k_best");
    return UAM_OK;
}
