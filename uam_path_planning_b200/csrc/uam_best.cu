// Best candidate of a batch: min over b of uam_best_key(cost[b], global index) (uam_internal.cuh) -- the running min of
// path_generation/main.py:162-180 -- and its exchange between the ranks of one box.
//
// Single GPU: per-thread / warp / CTA min, one atomicMin per CTA, the last CTA publishes the result (uam_best_tail_cta).
// Several GPUs: the same tail also IS the all-reduce.  Every rank owns a symmetric block (UamPeerBlock: cudaMalloc,
// exported with CUDA IPC, mapped by every peer process: uam_peer_export / uam_peer_attach); the last CTA stores its
// rank's key into its column of every peer's block over NVLink, raises an epoch-tagged flag behind a system-scope fence,
// waits for the flags of all ranks and takes the min.  No NCCL kernel, no extra launch: the collective rides in the last
// kernel of the scoring step (uam_k_reduce_paths in uam_score_raster.cu calls the same tail).
#include <algorithm>
#include <cstring>

#include "uam_internal.cuh"

namespace {

template <typename CT>
__global__ void __launch_bounds__(256)
uam_k_best(const CT* __restrict__ cost, long long B, UamBestTail tl) {
    unsigned long long best = UAM_KEY_EMPTY;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += stride) {
        const unsigned long long k = uam_best_key((float)cost[b], tl.offset + (unsigned long long)b);
        best = k < best ? k : best;
    }
    uam_best_tail_cta(tl, uam_cta_min_key(best));
}

__global__ void uam_k_set_u64(unsigned long long* p, unsigned long long v) { *p = v; }

__global__ void uam_k_best_init(unsigned long long* local, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) local[i] = (i % 4 == 0) ? UAM_KEY_EMPTY : 0ull;
}

int uam_best_ensure(uam_ctx* ctx) {
    if (ctx->d_best_local) return UAM_OK;
    const int n = 4 * (1 + UAM_HOST_PIPE_DEPTH);       // {min, done, status, pad} per slot
    UAM_CUDA(ctx, cudaMalloc(&ctx->d_best_local, sizeof(unsigned long long) * n));
    uam_k_best_init<<<1, 32, 0, nullptr>>>(ctx->d_best_local, n);
    UAM_CHECK_LAUNCH(ctx, "uam_k_best_init");
    UAM_CUDA(ctx, cudaStreamSynchronize(nullptr));
    return UAM_OK;
}

}  // namespace

int uam_best_tail(uam_ctx* ctx, int slot, unsigned long long offset, unsigned long long* d_out, bool with_peers, bool combine,
                  UamBestTail* tl) {
    UAM_TRY(uam_best_ensure(ctx));
    memset(tl, 0, sizeof *tl);
    tl->local = ctx->d_best_local + 4 * slot;
    tl->status = reinterpret_cast<unsigned*>(ctx->d_best_local + 4 * slot + 2);
    tl->out = d_out;
    tl->offset = offset;
    tl->combine = combine ? 1 : 0;
    tl->world = 1;
    if (with_peers && ctx->peer_world > 1) {
        tl->world = ctx->peer_world;
        tl->rank = ctx->peer_rank;
        tl->epoch = ++ctx->peer_epoch;
        for (int r = 0; r < ctx->peer_world; ++r) tl->peer[r] = ctx->peer_ptr[r];
    }
    return UAM_OK;
}

int uam_best_launch(uam_ctx* ctx, const void* d_cost, int cost_is_f64, int64_t B, const UamBestTail& tl, cudaStream_t st) {
    const long long ctas = std::max<long long>(1, std::min<long long>((B + 255) / 256, (long long)ctx->sm_count * 4));
    if (cost_is_f64) uam_k_best<double><<<(unsigned)ctas, 256, 0, st>>>((const double*)d_cost, B, tl);
    else uam_k_best<float><<<(unsigned)ctas, 256, 0, st>>>((const float*)d_cost, B, tl);
    UAM_CHECK_LAUNCH(ctx, "uam_k_best");
    return UAM_OK;
}

static int uam_best_check(uam_ctx* ctx, const void* d_cost, int64_t B, int64_t global_offset, const void* d_key) {
    if (!d_key || B < 0 || (B > 0 && !d_cost)) return uam_fail(ctx, UAM_ERR_INVALID, "bad argument to uam_best");
    if (global_offset < 0 || global_offset + B > 0x7fffffffll)
        return uam_fail(ctx, UAM_ERR_UNSUPPORTED, "global path index must fit 31 bits");
    return UAM_OK;
}

extern "C" int uam_best(uam_ctx* ctx, const void* d_cost, int cost_is_f64, int64_t B, int64_t global_offset,
                        uint64_t* d_key, int reset, void* stream) {
    if (!ctx) return UAM_ERR_INVALID;
    UAM_TRY(uam_best_check(ctx, d_cost, B, global_offset, d_key));
    UAM_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = uam_pick_stream(ctx, stream);
    if (reset) {
        uam_k_set_u64<<<1, 1, 0, st>>>((unsigned long long*)d_key, UAM_KEY_EMPTY);
        UAM_CHECK_LAUNCH(ctx, "uam_k_set_u64");
    }
    if (B == 0) return UAM_OK;
    UamBestTail tl;
    UAM_TRY(uam_best_tail(ctx, 0, (unsigned long long)global_offset, (unsigned long long*)d_key, false, true, &tl));
    return uam_best_launch(ctx, d_cost, cost_is_f64, B, tl, st);
}

extern "C" int uam_best_allreduce(uam_ctx* ctx, const void* d_cost, int cost_is_f64, int64_t B, int64_t global_offset,
                                  uint64_t* d_key, void* stream) {
    if (!ctx) return UAM_ERR_INVALID;
    UAM_TRY(uam_best_check(ctx, d_cost, B, global_offset, d_key));
    UAM_CUDA(ctx, cudaSetDevice(ctx->device));
    UamBestTail tl;
    UAM_TRY(uam_best_tail(ctx, 0, (unsigned long long)global_offset, (unsigned long long*)d_key, true, false, &tl));
    return uam_best_launch(ctx, d_cost, cost_is_f64, B, tl, uam_pick_stream(ctx, stream));   // B == 0: one CTA, empty key
}

// ---- peer group --------------------------------------------------------------------------------------------------
extern "C" int uam_peer_export(uam_ctx* ctx, void* h_handle64) {
    if (!ctx || !h_handle64) return UAM_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    UAM_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->d_peer_own) {
        UAM_CUDA(ctx, cudaMalloc(&ctx->d_peer_own, sizeof(UamPeerBlock)));
        UAM_CUDA(ctx, cudaMemset(ctx->d_peer_own, 0, sizeof(UamPeerBlock)));
        UAM_CUDA(ctx, cudaDeviceSynchronize());
    }
    cudaIpcMemHandle_t h;
    UAM_CUDA(ctx, cudaIpcGetMemHandle(&h, ctx->d_peer_own));
    memcpy(h_handle64, &h, 64);
    return UAM_OK;
}

extern "C" int uam_peer_attach(uam_ctx* ctx, int rank, int world, const void* h_handles) {
    if (!ctx) return UAM_ERR_INVALID;
    if (world < 1 || world > UAM_MAX_PEERS || rank < 0 || rank >= world || !h_handles)
        return uam_fail(ctx, UAM_ERR_INVALID, "bad peer group: rank %d of %d (at most %d ranks)", rank, world, UAM_MAX_PEERS);
    if (!ctx->d_peer_own) return uam_fail(ctx, UAM_ERR_STATE, "call uam_peer_export first");
    if (ctx->peer_world) return uam_fail(ctx, UAM_ERR_STATE, "a peer group is already attached to this ctx");
    UAM_CUDA(ctx, cudaSetDevice(ctx->device));
    for (int r = 0; r < world; ++r) {
        if (r == rank) { ctx->peer_ptr[r] = ctx->d_peer_own; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)h_handles + 64 * (size_t)r, 64);
        void* p = nullptr;
        UAM_CUDA(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->peer_ptr[r] = (UamPeerBlock*)p;
    }
    ctx->peer_world = world;
    ctx->peer_rank = rank;
    ctx->peer_epoch = 0;
    return UAM_OK;
}

extern "C" int uam_peer_status(uam_ctx* ctx, int* timed_out) {
    if (!ctx || !timed_out) return UAM_ERR_INVALID;
    *timed_out = 0;
    if (!ctx->d_best_local) return UAM_OK;
    UAM_CUDA(ctx, cudaSetDevice(ctx->device));
    unsigned long long h[4 * (1 + UAM_HOST_PIPE_DEPTH)];
    UAM_CUDA(ctx, cudaMemcpy(h, ctx->d_best_local, sizeof h, cudaMemcpyDeviceToHost));
    for (int s = 0; s <= UAM_HOST_PIPE_DEPTH; ++s) *timed_out |= (h[4 * s + 2] != 0);
    return UAM_OK;
}
