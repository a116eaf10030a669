// Best candidate of a batch: min over b of uam_best_key(cost[b], global index) (uam_internal.cuh) -- the device half of the
// running min of path_generation/main.py:162-180; the 8-byte key is min-reduced across ranks by the caller (NCCL).
#include <algorithm>

#include "uam_internal.cuh"

namespace {

template <typename CT>
__global__ void __launch_bounds__(256)
uam_k_best(const CT* __restrict__ cost, long long B, unsigned long long offset, unsigned long long* __restrict__ key) {
    unsigned long long best = UAM_KEY_EMPTY;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += stride) {
        const unsigned long long k = uam_best_key((float)cost[b], offset + (unsigned long long)b);
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
        if (best != UAM_KEY_EMPTY) atomicMin(key, best);
    }
}

__global__ void uam_k_set_u64(unsigned long long* p, unsigned long long v) { *p = v; }

}  // namespace

extern "C" int uam_best(uam_ctx* ctx, const void* d_cost, int cost_is_f64, int64_t B, int64_t global_offset,
                        uint64_t* d_key, int reset, void* stream) {
    if (!ctx) return UAM_ERR_INVALID;
    if (!d_key || B < 0 || (B > 0 && !d_cost)) return uam_fail(ctx, UAM_ERR_INVALID, "bad argument to uam_best");
    if (global_offset < 0 || global_offset + B > 0x7fffffffll)
        return uam_fail(ctx, UAM_ERR_UNSUPPORTED, "global path index must fit 31 bits");
    UAM_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = uam_pick_stream(ctx, stream);
    if (reset) {
        uam_k_set_u64<<<1, 1, 0, st>>>((unsigned long long*)d_key, UAM_KEY_EMPTY);
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
