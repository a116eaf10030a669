// Batched cost-to-go sweeps on an 8-connected grid (build-defined extension: the reference has no grid search;
// BASELINE.json config 5).  Q independent single-source problems share every launch.
//
// Edge u -> v costs step(u,v) * (cost[u] + cost[v]) with step = 2 (axis) / 3 (diagonal): all integer, so shortest
// distances are unique numbers and a label-correcting relaxation in ANY order converges to the same bits as
// Dijkstra.  Frontier-parallel relaxation, organised by 32 x 32 tiles:
//   round:  uam_k_grid_compact  collects the active (query, tile) pairs and clears their flags;
//           uam_k_grid_relax    one CTA per active pair loads the tile + 1-cell halo of dist/cost into shared memory,
//                               relaxes it to its local fixed point, writes the interior back and, if any cell of its
//                               boundary ring dropped, flags the 8 neighbouring tiles for the next round.
//   rounds repeat until no tile is active (the active count is read back every few rounds).
//   uam_k_grid_parent then picks each cell's predecessor: argmin over the 8 neighbours in a fixed slot order with a
//   strict '<' -- deterministic, identical to the oracle's post-pass.
#include <algorithm>

#include "uam_internal.cuh"

namespace {

#define GT 32                          // tile side
#define GH (GT + 2)                    // with halo
#define UAM_GRID_INF (1ll << 62)

__device__ __constant__ int c_di[8] = {-1, -1, -1, 0, 0, 1, 1, 1};
__device__ __constant__ int c_dj[8] = {-1, 0, 1, -1, 1, -1, 0, 1};
__device__ __constant__ int c_st[8] = {3, 2, 3, 2, 2, 3, 2, 3};

__global__ void __launch_bounds__(256)
uam_k_grid_init(long long* __restrict__ dist, const uint8_t* __restrict__ blocked, const int* __restrict__ sources, int Q,
                int H, int W, int tiles_x, int tiles_per_q, uint8_t* __restrict__ flags) {
    const size_t cells = (size_t)H * W;
    const size_t total = cells * Q;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) dist[t] = UAM_GRID_INF;
    // all threads of all CTAs must have finished the fill before the sources are written: done by a second launch
}

__global__ void uam_k_grid_seed(long long* __restrict__ dist, const uint8_t* __restrict__ blocked,
                                const int* __restrict__ sources, int Q, int H, int W, int tiles_x, int tiles_per_q,
                                uint8_t* __restrict__ flags) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    const int si = sources[2 * q], sj = sources[2 * q + 1];
    if (si < 0 || si >= H || sj < 0 || sj >= W) return;
    const size_t c = (size_t)si * W + sj;
    if (blocked && blocked[c]) return;
    dist[(size_t)q * H * W + c] = 0;
    flags[(size_t)q * tiles_per_q + (size_t)(si / GT) * tiles_x + sj / GT] = 1;
}

__global__ void __launch_bounds__(256)
uam_k_grid_compact(uint8_t* __restrict__ flags, size_t n, unsigned* __restrict__ list, unsigned* __restrict__ count) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
        if (flags[t]) {
            flags[t] = 0;
            list[atomicAdd(count, 1u)] = (unsigned)t;
        }
    }
}

__global__ void __launch_bounds__(256)
uam_k_grid_relax(const uint16_t* __restrict__ cost, const uint8_t* __restrict__ blocked, int H, int W, int tiles_x,
                 int tiles_y, const unsigned* __restrict__ list, const unsigned* __restrict__ count,
                 long long* __restrict__ dist, uint8_t* __restrict__ flags_next) {
    __shared__ long long sd[GH][GH + 1];
    __shared__ int sc[GH][GH + 1];           // cell cost, -1 = blocked / outside
    const int tiles_per_q = tiles_x * tiles_y;
    const unsigned n = *count;
    for (unsigned w = blockIdx.x; w < n; w += gridDim.x) {
        const unsigned ent = list[w];
        const int q = ent / tiles_per_q;
        const int tile = ent - q * tiles_per_q;
        const int ti = tile / tiles_x, tj = tile - ti * tiles_x;
        const int i0 = ti * GT - 1, j0 = tj * GT - 1;
        long long* dq = dist + (size_t)q * H * W;
        __syncthreads();
        for (int t = threadIdx.x; t < GH * GH; t += blockDim.x) {
            const int li = t / GH, lj = t - li * GH;
            const int i = i0 + li, j = j0 + lj;
            long long d = UAM_GRID_INF;
            int c = -1;
            if (i >= 0 && i < H && j >= 0 && j < W) {
                const size_t g = (size_t)i * W + j;
                if (!(blocked && blocked[g])) {
                    c = cost[g];
                    d = dq[g];
                }
            }
            sd[li][lj] = d;
            sc[li][lj] = c;
        }
        __syncthreads();
        // each thread owns 4 cells of the interior: rows (threadIdx.x >> 5) + 8k, column threadIdx.x & 31
        const int lj = (threadIdx.x & 31) + 1;
        const int lr = threadIdx.x >> 5;
        long long before[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) before[k] = sd[lr + 8 * k + 1][lj];
        bool any_change = false;
        int changed;
        do {
            changed = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int li = lr + 8 * k + 1;
                const int cv = sc[li][lj];
                if (cv < 0) continue;
                long long best = sd[li][lj];
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    const int ni = li + c_di[s], nj = lj + c_dj[s];
                    const int cn = sc[ni][nj];
                    if (cn < 0) continue;
                    const long long cand = sd[ni][nj] + (long long)(c_st[s] * (cn + cv));
                    best = cand < best ? cand : best;
                }
                if (best < sd[li][lj]) {
                    sd[li][lj] = best;
                    changed = 1;
                }
            }
            any_change = any_change || changed;
            changed = __syncthreads_or(changed);
        } while (changed);
        // write back + did the boundary ring change?
        int ring = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int li = lr + 8 * k + 1;
            const long long now = sd[li][lj];
            if (now < before[k]) {
                const int i = i0 + li, j = j0 + lj;
                dq[(size_t)i * W + j] = now;
                if (li == 1 || li == GT || lj == 1 || lj == GT) ring = 1;
            }
        }
        ring = __syncthreads_or(ring);
        if (ring && threadIdx.x < 8) {
            const int ni = ti + c_di[threadIdx.x], nj = tj + c_dj[threadIdx.x];
            if (ni >= 0 && ni < tiles_y && nj >= 0 && nj < tiles_x)
                flags_next[(size_t)q * tiles_per_q + (size_t)ni * tiles_x + nj] = 1;
        }
    }
}

__global__ void __launch_bounds__(256)
uam_k_grid_parent(const uint16_t* __restrict__ cost, const uint8_t* __restrict__ blocked, int H, int W, int Q,
                  const long long* __restrict__ dist, const int* __restrict__ sources, int* __restrict__ parent) {
    const size_t cells = (size_t)H * W;
    const size_t total = cells * Q;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int q = (int)(t / cells);
        const size_t v = t - (size_t)q * cells;
        const long long* dq = dist + (size_t)q * cells;
        int p = -1;
        if (dq[v] < UAM_GRID_INF) {
            const int vi = (int)(v / W), vj = (int)(v - (size_t)vi * W);
            if (vi == sources[2 * q] && vj == sources[2 * q + 1]) {
                p = (int)v;
            } else {
                const int cv = cost[v];
                long long best = UAM_GRID_INF;
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    const int ui = vi + c_di[s], uj = vj + c_dj[s];
                    if (ui < 0 || ui >= H || uj < 0 || uj >= W) continue;
                    const size_t u = (size_t)ui * W + uj;
                    const long long du = dq[u];
                    if (du >= UAM_GRID_INF) continue;
                    const long long nd = du + (long long)(c_st[s] * ((int)cost[u] + cv));
                    if (nd < best) { best = nd; p = (int)u; }
                }
            }
        }
        parent[t] = p;
    }
}

}  // namespace

extern "C" int uam_grid_search(uam_ctx* ctx, const uint16_t* d_cost, const uint8_t* d_blocked, int H, int W,
                               const int32_t* d_sources, int Q, int64_t* d_dist, int32_t* d_parent, void* stream) {
    if (!ctx) return UAM_ERR_INVALID;
    if (H < 1 || W < 1 || Q < 0 || !d_cost || !d_dist || (Q > 0 && !d_sources))
        return uam_fail(ctx, UAM_ERR_INVALID, "bad argument to uam_grid_search");
    if ((size_t)H * W >= 0x7fffffffull) return uam_fail(ctx, UAM_ERR_UNSUPPORTED, "grid has 2^31 or more cells");
    if (Q == 0) return UAM_OK;
    UAM_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = uam_pick_stream(ctx, stream);
    const int tiles_x = (W + GT - 1) / GT, tiles_y = (H + GT - 1) / GT;
    const size_t tiles_per_q = (size_t)tiles_x * tiles_y;
    const size_t n_flags = tiles_per_q * Q;
    if (n_flags >= 0xffffffffull) return uam_fail(ctx, UAM_ERR_UNSUPPORTED, "too many (query, tile) pairs");
    // scratch: flags A | flags B | list (u32) | count (u32 x 2)
    const size_t fbytes = (n_flags + 255) & ~(size_t)255;
    UAM_TRY(uam_reserve(ctx, &ctx->d_scratch, &ctx->scratch_bytes, 2 * fbytes + n_flags * 4 + 256));
    uint8_t* flags_a = (uint8_t*)ctx->d_scratch;
    uint8_t* flags_b = flags_a + fbytes;
    unsigned* list = (unsigned*)(flags_b + fbytes);
    unsigned* count = list + n_flags;
    UAM_CUDA(ctx, cudaMemsetAsync(flags_a, 0, 2 * fbytes, st));
    const int grid_fill = ctx->sm_count * 16;
    uam_k_grid_init<<<grid_fill, 256, 0, st>>>((long long*)d_dist, d_blocked, d_sources, Q, H, W, tiles_x, (int)tiles_per_q, flags_a);
    UAM_CHECK_LAUNCH(ctx, "uam_k_grid_init");
    uam_k_grid_seed<<<(Q + 127) / 128, 128, 0, st>>>((long long*)d_dist, d_blocked, d_sources, Q, H, W, tiles_x, (int)tiles_per_q, flags_a);
    UAM_CHECK_LAUNCH(ctx, "uam_k_grid_seed");
    uint8_t* cur = flags_a;
    uint8_t* nxt = flags_b;
    const int grid_relax = ctx->sm_count * 8;
    const long long max_rounds = 64ll * ((long long)tiles_x + tiles_y) * GT + 1024;     // far above any real front count
    unsigned h_count = 1;
    for (long long round = 0; round < max_rounds && h_count; ++round) {
        UAM_CUDA(ctx, cudaMemsetAsync(count, 0, 4, st));
        uam_k_grid_compact<<<ctx->sm_count * 4, 256, 0, st>>>(cur, n_flags, list, count);
        UAM_CHECK_LAUNCH(ctx, "uam_k_grid_compact");
        uam_k_grid_relax<<<grid_relax, 256, 0, st>>>(d_cost, d_blocked, H, W, tiles_x, tiles_y, list, count, (long long*)d_dist, nxt);
        UAM_CHECK_LAUNCH(ctx, "uam_k_grid_relax");
        std::swap(cur, nxt);
        if ((round & 7) == 7) {     // termination check every 8 rounds: rounds with an empty list are no-ops
            UAM_CUDA(ctx, cudaMemcpyAsync(&h_count, count, 4, cudaMemcpyDeviceToHost, st));
            UAM_CUDA(ctx, cudaStreamSynchronize(st));
        }
    }
    if (h_count) return uam_fail(ctx, UAM_ERR_STATE, "grid search did not converge");
    if (d_parent) {
        uam_k_grid_parent<<<grid_fill, 256, 0, st>>>(d_cost, d_blocked, H, W, Q, (const long long*)d_dist, d_sources, d_parent);
        UAM_CHECK_LAUNCH(ctx, "uam_k_grid_parent");
    }
    return UAM_OK;
}
