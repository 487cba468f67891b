// fp32 scorer on CUDA cores: the 1e-4 parity path for the shared MLP, plus the fp32 head
// (pooled 1024 -> 512 -> 256 -> 1) that both precisions use, and zs_score's dispatch.
// ABI: include/zs.h.  Reference call: model({"point_x": ...}), python/ossid/utils/zephyr_utils.py:34.
//
// Register tiling shared by every layer: a CTA of 256 threads computes a tile of 128 rows
// (model points, or hypotheses in the head) x 16*CT output channels; thread (cg = tid/16,
// pg = tid%16) owns CT channels x 8 rows.  Inputs are staged transposed in shared memory
// ([k][128 rows], conflict-free float4 reads); weights are read transposed ([k][co]) from
// global memory as warp-broadcast float4 loads that stay in L1.
#include "zs_common.cuh"

namespace {

constexpr int kRows = 128;      // rows per tile
constexpr int kThreadsMlp = 256;

struct f32_weights {
    const float *W1t, *b1, *W2t, *b2, *W3t, *b3;   // W*t: [K][CO]
};

template <int CT>
__device__ __forceinline__ void mm_tile(const float* __restrict__ in_t, int K, const float* __restrict__ Wt,
                                        int ldw, float (&acc)[CT][8], int pg) {
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        const float4 a0 = *reinterpret_cast<const float4*>(in_t + k * kRows + pg * 8);
        const float4 a1 = *reinterpret_cast<const float4*>(in_t + k * kRows + pg * 8 + 4);
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        float w[CT];
#pragma unroll
        for (int c = 0; c < CT; c += 4) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(Wt + (size_t)k * ldw + c));
            w[c] = t.x; w[c + 1] = t.y; w[c + 2] = t.z; w[c + 3] = t.w;
        }
#pragma unroll
        for (int c = 0; c < CT; ++c)
#pragma unroll
            for (int p = 0; p < 8; ++p) acc[c][p] = fmaf(w[c], a[p], acc[c][p]);
    }
}

template <int CT>
__device__ __forceinline__ void zero_acc(float (&acc)[CT][8]) {
#pragma unroll
    for (int c = 0; c < CT; ++c)
#pragma unroll
        for (int p = 0; p < 8; ++p) acc[c][p] = 0.f;
}

// out_t[co][row] = relu(acc + b[co])
template <int CT>
__device__ __forceinline__ void store_relu(float* __restrict__ out_t, const float* __restrict__ b, int co0,
                                           const float (&acc)[CT][8], int pg) {
#pragma unroll
    for (int c = 0; c < CT; ++c) {
        const float bb = __ldg(b + co0 + c);
        float4 o0, o1;
        o0.x = fmaxf(acc[c][0] + bb, 0.f); o0.y = fmaxf(acc[c][1] + bb, 0.f);
        o0.z = fmaxf(acc[c][2] + bb, 0.f); o0.w = fmaxf(acc[c][3] + bb, 0.f);
        o1.x = fmaxf(acc[c][4] + bb, 0.f); o1.y = fmaxf(acc[c][5] + bb, 0.f);
        o1.z = fmaxf(acc[c][6] + bb, 0.f); o1.w = fmaxf(acc[c][7] + bb, 0.f);
        *reinterpret_cast<float4*>(out_t + (co0 + c) * kRows + pg * 8) = o0;
        *reinterpret_cast<float4*>(out_t + (co0 + c) * kRows + pg * 8 + 4) = o1;
    }
}

// Shared MLP 8 -> 64 -> 128 -> 1024 + max over points; one CTA per hypothesis (grid stride).
__global__ void __launch_bounds__(kThreadsMlp, 1)
zs_k_mlp_f32(const float* __restrict__ feat, int n, int N, f32_weights w, float* __restrict__ pooled) {
    extern __shared__ __align__(16) float sm[];
    float* xt = sm;                    // [8][128]
    float* h1t = xt + 8 * kRows;       // [64][128]
    float* h2t = h1t + 64 * kRows;     // [128][128]
    float* pool = h2t + 128 * kRows;   // [1024]
    const int tid = threadIdx.x, pg = tid & 15, cg = tid >> 4;
    for (int h = blockIdx.x; h < n; h += gridDim.x) {
        for (int i = tid; i < 1024; i += kThreadsMlp) pool[i] = 0.f;   // ReLU output >= 0, so 0 is the identity of max
        for (int t0 = 0; t0 < N; t0 += kRows) {
            // features of 128 points -> xt[c][pt]
            for (int i = tid; i < kRows * 2; i += kThreadsMlp) {
                const int pt = i >> 1, half = i & 1;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (t0 + pt < N) v = __ldg(reinterpret_cast<const float4*>(feat + ((size_t)h * N + t0 + pt) * 8) + half);
                xt[(half * 4 + 0) * kRows + pt] = v.x; xt[(half * 4 + 1) * kRows + pt] = v.y;
                xt[(half * 4 + 2) * kRows + pt] = v.z; xt[(half * 4 + 3) * kRows + pt] = v.w;
            }
            __syncthreads();
            {
                float acc[4][8];
                zero_acc<4>(acc);
                mm_tile<4>(xt, 8, w.W1t + cg * 4, 64, acc, pg);
                store_relu<4>(h1t, w.b1, cg * 4, acc, pg);
            }
            __syncthreads();
            {
                float acc[8][8];
                zero_acc<8>(acc);
                mm_tile<8>(h1t, 64, w.W2t + cg * 8, 128, acc, pg);
                store_relu<8>(h2t, w.b2, cg * 8, acc, pg);
            }
            __syncthreads();
            for (int cb = 0; cb < 8; ++cb) {
                const int co0 = cb * 128 + cg * 8;
                float acc[8][8];
                zero_acc<8>(acc);
                mm_tile<8>(h2t, 128, w.W3t + co0, 1024, acc, pg);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float bb = __ldg(w.b3 + co0 + c);
                    float m = 0.f;
#pragma unroll
                    for (int p = 0; p < 8; ++p)
                        if (t0 + pg * 8 + p < N) m = fmaxf(m, acc[c][p] + bb);
                    // 16 consecutive lanes share cg: reduce over pg
                    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
                    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
                    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
                    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
                    if (pg == 0) pool[co0 + c] = fmaxf(pool[co0 + c], m);   // this thread is the only writer of co0+c
                }
            }
            // next tile's xt / h1t writes cannot race with h2t readers; the barrier after the xt load orders the rest
        }
        __syncthreads();
        for (int i = tid; i < 1024; i += kThreadsMlp) pooled[(size_t)h * 1024 + i] = pool[i];
        __syncthreads();
    }
}

// out[n][CO] = act(in[n][K] . Wt[K][CO] + b); grid (ceil(n/128), CO/128).
template <bool kRelu>
__global__ void __launch_bounds__(kThreadsMlp)
zs_k_fc(const float* __restrict__ in, const float* __restrict__ Wt, const float* __restrict__ b,
        float* __restrict__ out, int n, int K, int CO) {
    constexpr int KC = 32;
    __shared__ __align__(16) float in_t[KC * kRows];
    const int tid = threadIdx.x, pg = tid & 15, cg = tid >> 4;
    const int r0 = blockIdx.x * kRows, co0 = blockIdx.y * 128 + cg * 8;
    float acc[8][8];
    zero_acc<8>(acc);
    for (int k0 = 0; k0 < K; k0 += KC) {
        __syncthreads();
        for (int i = tid; i < kRows * (KC / 4); i += kThreadsMlp) {
            const int r = i / (KC / 4), q = i % (KC / 4);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r0 + r < n) v = __ldg(reinterpret_cast<const float4*>(in + (size_t)(r0 + r) * K + k0) + q);
            in_t[(q * 4 + 0) * kRows + r] = v.x; in_t[(q * 4 + 1) * kRows + r] = v.y;
            in_t[(q * 4 + 2) * kRows + r] = v.z; in_t[(q * 4 + 3) * kRows + r] = v.w;
        }
        __syncthreads();
        mm_tile<8>(in_t, KC, Wt + (size_t)k0 * CO + co0, CO, acc, pg);
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float bb = __ldg(b + co0 + c);
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            const int r = r0 + pg * 8 + p;
            float v = acc[c][p] + bb;
            if (kRelu) v = fmaxf(v, 0.f);
            if (r < n) out[(size_t)r * CO + co0 + c] = v;
        }
    }
}

// The same for a handful of rows (the fp32 re-rank of the top-k candidates: n = k x objects, a few hundred): a
// 128-row tile would leave all but a few SMs idle, so a CTA takes 8 rows x 64 output channels and its four warp pairs
// split K (partial sums reduced through shared memory).  grid (ceil(n/8), CO/64), 256 threads.
constexpr int kSmallRows = 8, kSmallCo = 64;
template <bool kRelu>
__global__ void __launch_bounds__(256)
zs_k_fc_small(const float* __restrict__ in, const float* __restrict__ Wt, const float* __restrict__ b,
              float* __restrict__ out, int n, int K, int CO) {
    extern __shared__ __align__(16) float sm_small[];
    float* in_t = sm_small;                                  // [K][8 rows]
    float* part = sm_small + (size_t)K * kSmallRows;         // [4 k-slices][8 rows][64 channels]
    const int tid = threadIdx.x, c = tid & 63, ks = tid >> 6;
    const int r0 = blockIdx.x * kSmallRows, co = blockIdx.y * kSmallCo + c;
    for (int i = tid; i < kSmallRows * K; i += 256) {
        const int r = i / K, k = i - r * K;
        in_t[k * kSmallRows + r] = (r0 + r < n) ? __ldg(in + (size_t)(r0 + r) * K + k) : 0.f;
    }
    __syncthreads();
    float acc[kSmallRows];
#pragma unroll
    for (int r = 0; r < kSmallRows; ++r) acc[r] = 0.f;
    const int kq = K / 4;
    // the weight column is the only global read of the loop: keep eight loads in flight (the loop is bound by their L2
    // latency, not by the 8 FMAs per weight)
    for (int k0 = ks * kq; k0 < (ks + 1) * kq; k0 += 8) {
        float w[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) w[u] = __ldg(Wt + (size_t)(k0 + u) * CO + co);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float4 a0 = *reinterpret_cast<const float4*>(in_t + (k0 + u) * kSmallRows);
            const float4 a1 = *reinterpret_cast<const float4*>(in_t + (k0 + u) * kSmallRows + 4);
            acc[0] = fmaf(w[u], a0.x, acc[0]); acc[1] = fmaf(w[u], a0.y, acc[1]); acc[2] = fmaf(w[u], a0.z, acc[2]); acc[3] = fmaf(w[u], a0.w, acc[3]);
            acc[4] = fmaf(w[u], a1.x, acc[4]); acc[5] = fmaf(w[u], a1.y, acc[5]); acc[6] = fmaf(w[u], a1.z, acc[6]); acc[7] = fmaf(w[u], a1.w, acc[7]);
        }
    }
#pragma unroll
    for (int r = 0; r < kSmallRows; ++r) part[(ks * kSmallRows + r) * kSmallCo + c] = acc[r];
    __syncthreads();
    for (int i = tid; i < kSmallRows * kSmallCo; i += 256) {
        const int r = i >> 6, cc = i & 63;
        float v = ((part[(0 * kSmallRows + r) * kSmallCo + cc] + part[(1 * kSmallRows + r) * kSmallCo + cc]) +
                   (part[(2 * kSmallRows + r) * kSmallCo + cc] + part[(3 * kSmallRows + r) * kSmallCo + cc])) +
                  __ldg(b + blockIdx.y * kSmallCo + cc);
        if (kRelu) v = fmaxf(v, 0.f);
        if (r0 + r < n) out[(size_t)(r0 + r) * CO + blockIdx.y * kSmallCo + cc] = v;
    }
}

// scores[h] = g2[h] . F3 + c3 ; one warp per hypothesis.
__global__ void zs_k_fc_out(const float* __restrict__ g2, const float* __restrict__ F3, const float* __restrict__ c3,
                            float* __restrict__ scores, int n) {
    const int lane = threadIdx.x & 31;
    const int h = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (h >= n) return;
    float s = 0.f;
    for (int k = lane * 4; k < 256; k += 128) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(g2 + (size_t)h * 256 + k));
        const float4 f = __ldg(reinterpret_cast<const float4*>(F3 + k));
        s += a.x * f.x + a.y * f.y + a.z * f.z + a.w * f.w;
    }
    for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (lane == 0) scores[h] = s + __ldg(c3);
}

// Wt[k][co] = W[co][k]
__global__ void zs_k_transpose(const float* __restrict__ W, float* __restrict__ Wt, int CO, int K) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < CO * K) {
        const int co = i / K, k = i % K;
        Wt[(size_t)k * CO + co] = W[i];
    }
}

constexpr int kOffW1t = 0, kOffW2t = kOffW1t + 8 * 64, kOffW3t = kOffW2t + 64 * 128,
              kOffF1t = kOffW3t + 128 * 1024, kOffF2t = kOffF1t + 1024 * 512, kTransFloats = kOffF2t + 512 * 256;
constexpr int kScoreChunk = ZS_SCORE_CHUNK;

}  // namespace

int zs_f32_prepare_weights(zs_ctx* ctx, int slot, cudaStream_t st) {
    zs_weights& w = ctx->w[slot];
    if (!w.f32t && cudaMalloc(&w.f32t, kTransFloats * sizeof(float)) != cudaSuccess) {
        cudaGetLastError();
        return zs_fail(ctx, ZS_ERR_NOMEM, "transposed weights");
    }
    const struct { int src, dst, co, k; } jobs[] = {
        {ZS_OFF_W1, kOffW1t, 64, 8}, {ZS_OFF_W2, kOffW2t, 128, 64}, {ZS_OFF_W3, kOffW3t, 1024, 128},
        {ZS_OFF_F1, kOffF1t, 512, 1024}, {ZS_OFF_F2, kOffF2t, 256, 512}};
    for (const auto& j : jobs) {
        zs_k_transpose<<<(j.co * j.k + 255) / 256, 256, 0, st>>>(w.f32 + j.src, w.f32t + j.dst, j.co, j.k);
        ZS_LAUNCHED(ctx);
    }
    return ZS_OK;
}

// pooled [m][1024] -> scores [m]; g1 [m][512], g2 [m][256] scratch.
static int head_impl(zs_ctx* ctx, int slot, const float* pooled, int n, float* scores, float* g1, float* g2,
                     int precision, cudaStream_t st) {
    if (precision == ZS_BF16) return zs_head_tc(ctx, slot, pooled, n, scores, g1, false, nullptr, st);   // tensor cores (tf32)
    if (precision == ZS_BF16_SPLIT && n > 1024)                      // fp32-accurate on the tensor cores (3-term tf32)
        return zs_head_tc(ctx, slot, pooled, n, scores, g1, true, g2 + (size_t)n * 256, st);
    const zs_weights& w = ctx->w[slot];
    if (n <= 1024) {         // a handful of rows (the re-rank): small tiles so that the whole GPU takes part
        const int tiles = (n + kSmallRows - 1) / kSmallRows;
        const size_t sm1 = (size_t)(1024 * kSmallRows + 4 * kSmallRows * kSmallCo) * sizeof(float);
        const size_t sm2 = (size_t)(512 * kSmallRows + 4 * kSmallRows * kSmallCo) * sizeof(float);
        zs_k_fc_small<true><<<dim3(tiles, 512 / kSmallCo), 256, sm1, st>>>(pooled, w.f32t + kOffF1t, w.f32 + ZS_OFF_C1, g1, n, 1024, 512);
        ZS_LAUNCHED(ctx);
        zs_k_fc_small<true><<<dim3(tiles, 256 / kSmallCo), 256, sm2, st>>>(g1, w.f32t + kOffF2t, w.f32 + ZS_OFF_C2, g2, n, 512, 256);
        ZS_LAUNCHED(ctx);
        zs_k_fc_out<<<(n * 32 + 255) / 256, 256, 0, st>>>(g2, w.f32 + ZS_OFF_F3, w.f32 + ZS_OFF_C3, scores, n);
        ZS_LAUNCHED(ctx);
        return ZS_OK;
    }
    dim3 g_fc1((n + kRows - 1) / kRows, 512 / 128), g_fc2((n + kRows - 1) / kRows, 256 / 128);
    zs_k_fc<true><<<g_fc1, kThreadsMlp, 0, st>>>(pooled, w.f32t + kOffF1t, w.f32 + ZS_OFF_C1, g1, n, 1024, 512);
    ZS_LAUNCHED(ctx);
    zs_k_fc<true><<<g_fc2, kThreadsMlp, 0, st>>>(g1, w.f32t + kOffF2t, w.f32 + ZS_OFF_C2, g2, n, 512, 256);
    ZS_LAUNCHED(ctx);
    zs_k_fc_out<<<(n * 32 + 255) / 256, 256, 0, st>>>(g2, w.f32 + ZS_OFF_F3, w.f32 + ZS_OFF_C3, scores, n);
    ZS_LAUNCHED(ctx);
    return ZS_OK;
}

static int pool_impl(zs_ctx* ctx, int slot, const void* feat, int feat_dtype, int m, int n_pts, float* pooled, cudaStream_t st) {
    const zs_weights& w = ctx->w[slot];
    if (feat_dtype == ZS_F32) {
        if (ctx->dyn_n) return zs_fail(ctx, ZS_ERR_UNSUPPORTED, "device-side counts (zs_set_dynamic_count) need bf16 features");
        const size_t smem = (size_t)(8 + 64 + 128) * kRows * sizeof(float) + 1024 * sizeof(float);
        ZS_CUDA(ctx, cudaFuncSetAttribute(zs_k_mlp_f32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        f32_weights fw{w.f32t + kOffW1t, w.f32 + ZS_OFF_B1, w.f32t + kOffW2t, w.f32 + ZS_OFF_B2,
                       w.f32t + kOffW3t, w.f32 + ZS_OFF_B3};
        const int grid = m < ctx->sm_count ? m : ctx->sm_count;
        zs_k_mlp_f32<<<grid, kThreadsMlp, smem, st>>>((const float*)feat, m, n_pts, fw, pooled);
        ZS_LAUNCHED(ctx);
        return ZS_OK;
    }
    if (feat_dtype == ZS_BF16_SPLIT) return zs_score_tc3(ctx, slot, feat, m, n_pts, pooled, nullptr, nullptr, st);
    return zs_score_tc(ctx, slot, (const __nv_bfloat16*)feat, m, n_pts, pooled, st);
}

static int score_args(zs_ctx* ctx, int weight_slot, const void* feat, int feat_dtype, int n, int n_pts) {
    if (!ctx) return ZS_ERR_INVALID;
    if (weight_slot < 0 || weight_slot >= ZS_MAX_WEIGHT_SLOTS || !ctx->w[weight_slot].set)
        return zs_fail(ctx, ZS_ERR_STATE, "weight slot %d not set", weight_slot);
    if (n < 0 || n_pts <= 0) return zs_fail(ctx, ZS_ERR_INVALID, "n %d n_pts %d", n, n_pts);
    if (feat_dtype != ZS_F32 && feat_dtype != ZS_BF16 && feat_dtype != ZS_BF16_SPLIT)
        return zs_fail(ctx, ZS_ERR_INVALID, "feat_dtype %d", feat_dtype);
    if (n > 0 && (!feat || ((uintptr_t)feat & 15))) return zs_fail(ctx, ZS_ERR_INVALID, "feat must be non-null, 16-byte aligned");
    return ZS_OK;
}

extern "C" int zs_pool(zs_ctx* ctx, int weight_slot, const void* feat, int feat_dtype, int n, int n_pts,
                       float* pooled_out, void* stream) {
    int rc = score_args(ctx, weight_slot, feat, feat_dtype, n, n_pts);
    if (rc) return rc;
    if (n == 0) return ZS_OK;
    if (!pooled_out) return zs_fail(ctx, ZS_ERR_INVALID, "pooled_out");
    ZS_CUDA(ctx, cudaSetDevice(ctx->device));
    return pool_impl(ctx, weight_slot, feat, feat_dtype, n, n_pts, pooled_out, (cudaStream_t)stream);
}

extern "C" int zs_head(zs_ctx* ctx, int weight_slot, const float* pooled, int n, int precision, float* scores_out,
                       void* stream) {
    if (!ctx) return ZS_ERR_INVALID;
    if (weight_slot < 0 || weight_slot >= ZS_MAX_WEIGHT_SLOTS || !ctx->w[weight_slot].set)
        return zs_fail(ctx, ZS_ERR_STATE, "weight slot %d not set", weight_slot);
    if (n < 0 || (precision != ZS_F32 && precision != ZS_BF16 && precision != ZS_BF16_SPLIT) ||
        (n > 0 && (!pooled || !scores_out || ((uintptr_t)pooled & 15))))
        return zs_fail(ctx, ZS_ERR_INVALID, "zs_head arguments");
    if (n == 0) return ZS_OK;
    ZS_CUDA(ctx, cudaSetDevice(ctx->device));
    const int chunk = n < kScoreChunk ? n : kScoreChunk;
    int rc = zs_reserve_ws(ctx, (size_t)chunk * ZS_HEAD_WS_FLOATS * sizeof(float));
    if (rc) return rc;
    float* g1 = (float*)ctx->ws + (size_t)chunk * 1024;
    float* g2 = g1 + (size_t)chunk * 512;      // followed by the lo(pooled) | lo(g1) scratch of the 3-term head
    for (int s = 0; s < n; s += chunk) {
        const int m = (n - s) < chunk ? (n - s) : chunk;
        rc = head_impl(ctx, weight_slot, pooled + (size_t)s * 1024, m, scores_out + s, g1, g2, precision, (cudaStream_t)stream);
        if (rc) return rc;
    }
    return ZS_OK;
}

extern "C" int zs_score(zs_ctx* ctx, int weight_slot, const void* feat, int feat_dtype, int n, int n_pts,
                        int precision, float* scores_out, void* stream) {
    int rc = score_args(ctx, weight_slot, feat, feat_dtype, n, n_pts);
    if (rc) return rc;
    if (precision != feat_dtype && !(precision == ZS_F32 && feat_dtype == ZS_BF16_SPLIT))
        return zs_fail(ctx, ZS_ERR_UNSUPPORTED, "precision %d needs matching feature dtype (got %d)", precision, feat_dtype);
    if (ctx->dyn_n) return zs_fail(ctx, ZS_ERR_UNSUPPORTED, "device-side counts: use zs_pool + zs_head");
    if (n == 0) return ZS_OK;
    if (!scores_out) return zs_fail(ctx, ZS_ERR_INVALID, "scores_out");
    ZS_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int chunk = n < kScoreChunk ? n : kScoreChunk;
    rc = zs_reserve_ws(ctx, (size_t)chunk * ZS_HEAD_WS_FLOATS * sizeof(float));
    if (rc) return rc;
    float* pooled = (float*)ctx->ws;
    float* g1 = pooled + (size_t)chunk * 1024;
    float* g2 = g1 + (size_t)chunk * 512;
    const size_t esz = feat_dtype == ZS_BF16 ? 2 : 4;          // ZS_BF16_SPLIT: two bf16 planes = 32 bytes per point as well
    for (int s = 0; s < n; s += chunk) {
        const int m = (n - s) < chunk ? (n - s) : chunk;
        rc = pool_impl(ctx, weight_slot, (const char*)feat + (size_t)s * n_pts * 8 * esz, feat_dtype, m, n_pts, pooled, st);
        if (rc) return rc;
        rc = head_impl(ctx, weight_slot, pooled, m, scores_out + s, g1, g2,
                       feat_dtype == ZS_BF16_SPLIT ? ZS_BF16_SPLIT : precision, st);   // split path: tensor-core head as well
        if (rc) return rc;
    }
    return ZS_OK;
}
