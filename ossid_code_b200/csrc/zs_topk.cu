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
// An entry whose index_map value is negative is an empty slot and is skipped; a NaN score sorts below every real
// score and is reported as -inf ("NaN never wins"); slots with no candidate left are (-inf, -1).
__global__ void __launch_bounds__(1024)
zs_k_topk(const float* __restrict__ scores, int n, int k, int index_base, const int32_t* __restrict__ index_map,
          const int32_t* __restrict__ seg, float* __restrict__ s_out, int32_t* __restrict__ i_out) {
    if (seg) {
        const int32_t* e = seg + 4 * blockIdx.x;
        scores += e[0];
        if (index_map) index_map += e[0] + e[3];
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
            if (index_map && __ldg(index_map + i) < 0) continue;             // empty slot
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
                if (best != 0ull) {                                          // keys of real candidates are never 0 (i < 2^31)
                    const int idx = (int)(0xffffffffu - (uint32_t)(best & 0xffffffffull));
                    const float sc = scores[idx];
                    s_out[r] = sc == sc ? sc : -INFINITY;
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

// Merge of the all-gathered per-rank candidate records (one warp per object).  Record of one rank, `rec_ints` int32:
//   [0, n_obj*k)            score bits of its k candidates per object, ordered by (score desc, index asc)
//   [n_obj*k, 2*n_obj*k)    their global hypothesis indices (-1 = empty slot)
//   [2*n_obj*k, +2*n_obj)   per object {hypotheses that passed the free-space pre-filter on this rank, violation
//                           count of its never-empty fallback}
//   [that rounded up to a multiple of 4, +12*n_obj*k)  (optional) the candidates' poses
// Never-empty rule made global (oracle violation_filter): if any rank kept a hypothesis of the object, the fallback
// candidates of ranks that kept none are dropped; if no rank kept any, only the first minimum-violation fallback
// survives (ranks own ascending index ranges, so lowest rank = lowest index on equal counts).
// Order: (score desc, global index asc); NaN sorts last and is reported as -inf.
__global__ void __launch_bounds__(32)
zs_k_merge_topk(const int32_t* __restrict__ gathered, int world, int rec_ints, int n_obj, int k,
                float* __restrict__ s_out, int32_t* __restrict__ i_out, float* __restrict__ p_out) {
    const int o = blockIdx.x, lane = threadIdx.x;
    const int n = world * k;
    bool any_kept = false;
    int fb_rank = -1;
    unsigned long long fb_best = ~0ull;
    for (int w = 0; w < world; ++w) {
        const int32_t* rec = gathered + (size_t)w * rec_ints;
        const int kept = rec[2 * n_obj * k + 2 * o], viol = rec[2 * n_obj * k + 2 * o + 1];
        any_kept |= kept > 0;
        if (kept <= 0 && rec[n_obj * k + o * k] >= 0) {               // this rank contributes a fallback candidate
            const unsigned long long key = ((unsigned long long)(uint32_t)viol << 32) | (uint32_t)w;
            if (key < fb_best) { fb_best = key; fb_rank = w; }
        }
    }
    unsigned long long prev = ~0ull;
    for (int r = 0; r < k; ++r) {
        unsigned long long best = 0ull;
        int best_pos = -1;
        for (int c = lane; c < n; c += 32) {
            const int w = c / k, j = c - w * k;
            const int32_t* rec = gathered + (size_t)w * rec_ints;
            const int idx = rec[n_obj * k + o * k + j];
            if (idx < 0) continue;
            const bool rank_kept = rec[2 * n_obj * k + 2 * o] > 0;
            if (any_kept ? !rank_kept : (w != fb_rank)) continue;
            const unsigned long long key = ((unsigned long long)f2key(__int_as_float(rec[o * k + j])) << 32) |
                                           (0xffffffffu - (uint32_t)idx);
            if (key < prev && key > best) { best = key; best_pos = c; }
        }
        for (int d = 16; d; d >>= 1) {
            const unsigned long long t = __shfl_xor_sync(0xffffffffu, best, d);
            const int tp = __shfl_xor_sync(0xffffffffu, best_pos, d);
            if (t > best) { best = t; best_pos = tp; }
        }
        const int w_best = best != 0ull ? best_pos / k : 0, j_best = best != 0ull ? best_pos - w_best * k : 0;
        if (lane == 0) {
            if (best != 0ull) {
                const float sc = __int_as_float(gathered[(size_t)w_best * rec_ints + o * k + j_best]);
                s_out[(size_t)o * k + r] = sc == sc ? sc : -INFINITY;
                i_out[(size_t)o * k + r] = (int)(0xffffffffu - (uint32_t)(best & 0xffffffffull));
            } else {
                s_out[(size_t)o * k + r] = -INFINITY;
                i_out[(size_t)o * k + r] = -1;
            }
        }
        if (p_out && lane < 12)                                          // the winner's pose travels with it
            p_out[((size_t)o * k + r) * 12 + lane] = best != 0ull
                ? __int_as_float(gathered[(size_t)w_best * rec_ints + ((2 * n_obj * k + 2 * n_obj + 3) & ~3) + (o * k + j_best) * 12 + lane]) : 0.f;
        prev = best;                       // 0 when nothing is left: every later pass finds nothing either
    }
}

}  // namespace

extern "C" int zs_merge_topk(zs_ctx* ctx, const int32_t* gathered, int world, int rec_ints, int n_obj, int k,
                             float* s_out, int32_t* i_out, float* poses_out, void* stream) {
    if (!ctx) return ZS_ERR_INVALID;
    if (n_obj == 0) return ZS_OK;
    if (world <= 0 || n_obj < 0 || k <= 0 || k > ZS_MAX_TOPK || !gathered || !s_out || !i_out ||
        rec_ints < (poses_out ? ((2 * n_obj * k + 2 * n_obj + 3) & ~3) + 12 * n_obj * k : 2 * n_obj * k + 2 * n_obj))
        return zs_fail(ctx, ZS_ERR_INVALID, "zs_merge_topk world %d n_obj %d k %d rec_ints %d", world, n_obj, k, rec_ints);
    ZS_CUDA(ctx, cudaSetDevice(ctx->device));
    zs_k_merge_topk<<<n_obj, 32, 0, (cudaStream_t)stream>>>(gathered, world, rec_ints, n_obj, k, s_out, i_out, poses_out);
    ZS_LAUNCHED(ctx);
    return ZS_OK;
}

namespace {
__global__ void zs_k_gather_poses(const float* __restrict__ poses, const int32_t* __restrict__ idx,
                                  const int32_t* __restrict__ seg, int n_seg, int k, float* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_seg * k * 12) return;
    const int c = t / 12, e = t - c * 12, g = c / k;
    const int i = idx[c];
    out[t] = i >= 0 ? poses[((size_t)seg[4 * g] + (size_t)(i - seg[4 * g + 1])) * 12 + e] : 0.f;
}
}  // namespace

extern "C" int zs_gather_poses(zs_ctx* ctx, const float* poses, const int32_t* idx, const int32_t* segments, int n_seg, int k,
                               float* out, void* stream) {
    if (!ctx) return ZS_ERR_INVALID;
    if (n_seg == 0) return ZS_OK;
    if (n_seg < 0 || k <= 0 || k > ZS_MAX_TOPK || !poses || !idx || !segments || !out)
        return zs_fail(ctx, ZS_ERR_INVALID, "zs_gather_poses n_seg %d k %d", n_seg, k);
    ZS_CUDA(ctx, cudaSetDevice(ctx->device));
    zs_k_gather_poses<<<(n_seg * k * 12 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(poses, idx, segments, n_seg, k, out);
    ZS_LAUNCHED(ctx);
    return ZS_OK;
}

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
