// Per-object top-k by (score desc, index asc).  ABI: include/zs.h.
// Generalises `scores.argmax()` (python/ossid/scripts/online_learning.py:466-467): numpy's
// argmax returns the FIRST maximum, so ties break towards the lower index; that rule also
// makes the result independent of how hypotheses are sharded over GPUs.
#include "zs_common.cuh"

namespace {

// Monotone map float -> uint32 (larger float = larger key); NaN sorts below -inf.
__device__ __forceinline__ uint32_t f2key(float f) {
    if (f != f) return 0u;
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// Selection by k passes of a block-wide arg-max over keys strictly below the previous winner.
// key = (f2key(score) << 32) | (0xffffffff - index): unique per element, max = best.
// One CTA per segment (object): segment g covers scores[seg[4g] .. seg[4g]+seg[4g+1]), reports index + seg[4g+2].
// seg == nullptr: a single segment [0, n) with index_base (the zs_topk entry point).
__global__ void __launch_bounds__(1024)
zs_k_topk(const float* __restrict__ scores, int n, int k, int index_base, const int32_t* __restrict__ index_map,
          const int32_t* __restrict__ seg, float* __restrict__ s_out, int32_t* __restrict__ i_out) {
    if (seg) {
        const int32_t* e = seg + 4 * blockIdx.x;
        scores += e[0];
        if (index_map) index_map += e[0];
        n = e[1];
        index_base = e[2];
        s_out += (size_t)blockIdx.x * k;
        i_out += (size_t)blockIdx.x * k;
    }
    __shared__ unsigned long long s_best[32];
    __shared__ unsigned long long s_prev;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_prev = ~0ull;
    __syncthreads();
    for (int r = 0; r < k; ++r) {
        const unsigned long long prev = s_prev;
        unsigned long long best = 0ull;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const unsigned long long key = ((unsigned long long)f2key(__ldg(scores + i)) << 32) | (0xffffffffu - (uint32_t)i);
            if (key < prev && key > best) best = key;
        }
        for (int d = 16; d; d >>= 1) {
            const unsigned long long t = __shfl_xor_sync(0xffffffffu, best, d);
            best = t > best ? t : best;
        }
        if (lane == 0) s_best[wid] = best;
        __syncthreads();
        if (wid == 0) {
            best = s_best[lane];
            for (int d = 16; d; d >>= 1) {
                const unsigned long long t = __shfl_xor_sync(0xffffffffu, best, d);
                best = t > best ? t : best;
            }
            if (lane == 0) {
                if (r < n) {
                    const int idx = (int)(0xffffffffu - (uint32_t)(best & 0xffffffffull));
                    s_out[r] = scores[idx];
                    i_out[r] = (index_map ? index_map[idx] : idx) + index_base;
                } else {
                    s_out[r] = -INFINITY;
                    i_out[r] = -1;
                }
                s_prev = best;
            }
        }
        __syncthreads();
    }
}

}  // namespace

extern "C" int zs_topk(zs_ctx* ctx, const float* scores, int n, int k, int index_base, const int32_t* index_map,
                       float* s_out, int32_t* i_out, void* stream) {
    if (!ctx) return ZS_ERR_INVALID;
    if (n < 0 || k <= 0 || k > ZS_MAX_TOPK || !s_out || !i_out || (n > 0 && !scores))
        return zs_fail(ctx, ZS_ERR_INVALID, "zs_topk n %d k %d", n, k);
    ZS_CUDA(ctx, cudaSetDevice(ctx->device));
    zs_k_topk<<<1, 1024, 0, (cudaStream_t)stream>>>(scores, n, k, index_base, index_map, nullptr, s_out, i_out);
    ZS_LAUNCHED(ctx);
    return ZS_OK;
}

extern "C" int zs_topk_segments(zs_ctx* ctx, const float* scores, const int32_t* segments, int n_segments, int k,
                                const int32_t* index_map, float* s_out, int32_t* i_out, void* stream) {
    if (!ctx) return ZS_ERR_INVALID;
    if (n_segments == 0) return ZS_OK;
    if (n_segments < 0 || k <= 0 || k > ZS_MAX_TOPK || !s_out || !i_out || !segments)
        return zs_fail(ctx, ZS_ERR_INVALID, "zs_topk_segments n_segments %d k %d", n_segments, k);
    ZS_CUDA(ctx, cudaSetDevice(ctx->device));
    zs_k_topk<<<n_segments, 1024, 0, (cudaStream_t)stream>>>(scores, 0, k, 0, index_map, segments, s_out, i_out);
    ZS_LAUNCHED(ctx);
    return ZS_OK;
}
