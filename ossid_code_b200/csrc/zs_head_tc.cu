// Head of the scorer on tensor cores: pooled[n][1024] -> fc1(512)+ReLU -> fc2(256)+ReLU -> fc3(1).
// Used by the tensor-core scoring path; the fp32 CUDA-core head (zs_score_f32.cu) stays the 1e-4
// parity path.  Reference call: model({"point_x": ...}), python/ossid/utils/zephyr_utils.py:34.
//
// One generic kernel, out[n][CO] = relu(A[n][K] . W[CO][K]^T + b), kind::tf32 (fp32 operands read
// straight from the fp32 buffers, 10-bit mantissa in the multiplier, fp32 accumulate in TMEM):
//   * CTA tile 256 rows x 256 output channels (two 128-row accumulators = all 512 TMEM columns), K walked in
//     blocks of 32 floats (= one 128-byte swizzle row); A and W tiles arrive by TMA (cp.async.bulk.tensor.2d,
//     SWIZZLE_128B) through a 3-stage mbarrier ring of 64 KB stages; 8 MMAs (M=128, N=256, K=8) per block
//     issued by one thread.  With kTerms = 3 the K loop runs three times (hi.hi + lo.hi + hi.lo): fp32-accurate.
//   * warps: 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2-5 = epilogue (thread = row).
//   * fc2's epilogue folds fc3 in: each thread owns all 256 hidden values of its hypothesis and
//     reduces them against F3 in registers, so only one float per hypothesis is written.
#include <cuda.h>
#include <cudaTypedefs.h>

#include "zs_common.cuh"

namespace {

constexpr int kThreadsFc = 192;
constexpr int kFcStages = 3;
// CTA tile = 256 rows x 256 output channels as TWO 128-row MMAs per K step (two 256-column accumulators, all 512 TMEM
// columns): per K-block a CTA pulls 32 KB of A + 32 KB of W for 8 MMAs instead of 16 + 32 KB for 4.  fc1 is bound by
// L2 -> shared-memory traffic (every CTA streams all of F1), so a third less traffic per MMA is a third less time.
constexpr int kBM = 256, kMmaM = 128, kBN = 256, kBK = 32;    // rows, rows per MMA, output channels, floats per k-block
constexpr uint32_t kStageA = kBM * kBK * 4;                   // 32 KB
constexpr uint32_t kStageW = kBN * kBK * 4;                   // 32 KB
constexpr uint32_t kStageBytes = kStageA + kStageW;           // 64 KB
constexpr uint32_t kFcBar = kFcStages * kStageBytes;          // barriers after the ring
constexpr uint32_t kFcTmemPtr = kFcBar + 128;
constexpr uint32_t kFcSmAlloc = kFcTmemPtr + 16 + 1024;
static_assert(kFcSmAlloc <= 232448, "exceeds 227 KB of shared memory per CTA");
// fc2's fused fc3 epilogue writes out[row] = partial dot + c3 from ONE CTA per row tile: it needs grid.x = 256 / kBN = 1
static_assert(kBN == 256, "the fused fc3 epilogue assumes one CTA covers all 256 channels of fc2");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
// K-major, 128-byte-swizzled operand descriptor (8-row groups 1024 B apart), version 1
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// fp32 accumulate, tf32 x tf32, both K-major
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// kFuseOut: instead of storing relu(acc + b), reduce it against `f3` and store one score per row.
// n_terms = 3: fp32-accurate GEMM out of tf32 MMAs.  The tensor core reads an fp32 operand as tf32 by ignoring its low 13
// mantissa bits, i.e. it sees hi(x) = x & 0xffffe000; with lo(x) = x - hi(x) (exact in fp32) the product is accumulated as
// hi(a).hi(w) + lo(a).hi(w) + hi(a).lo(w): the K loop simply runs three times, over (A, W), (A_lo, W), (A, W_lo).  The
// dropped lo.lo term and the truncation of lo itself are 2^-20 relative.  out_lo (optional) receives lo(out) so that
// the next layer needs no extra pass.
__device__ __forceinline__ float tf32_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

template <bool kFuseOut, int kTerms>
__global__ void __launch_bounds__(kThreadsFc, 1)
zs_k_fc_tc(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
           const __grid_constant__ CUtensorMap map_a_lo, const __grid_constant__ CUtensorMap map_w_lo,
           const float* __restrict__ bias, float* __restrict__ out, float* __restrict__ out_lo, int n, int K, int CO,
           const float* __restrict__ f3, const float* __restrict__ c3) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (sbase - smem_u32(smem_raw));
    auto full = [&](int s) { return sbase + kFcBar + 8u * (uint32_t)s; };
    auto empty = [&](int s) { return sbase + kFcBar + 8u * (uint32_t)(kFcStages + s); };
    const uint32_t acc_full = sbase + kFcBar + 8u * (2 * kFcStages);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int co0 = blockIdx.x * kBN, row0 = blockIdx.y * kBM;
    const int n_kb1 = K / kBK, n_kb = n_kb1 * kTerms;

    if (tid == 0) {
        for (int s = 0; s < kFcStages; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + kFcTmemPtr), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + kFcTmemPtr);

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < n_kb; ++kb) {
                const int s = kb % kFcStages;
                mbar_wait(empty(s), ((kb / kFcStages) & 1) ^ 1);
                mbar_expect_tx(full(s), kStageBytes);
                const int term = kTerms == 1 ? 0 : kb / n_kb1, kk = kb - term * n_kb1;
                tma_load_2d(sbase + s * kStageBytes, (kTerms == 3 && term == 1) ? &map_a_lo : &map_a, kk * kBK, row0, full(s));
                tma_load_2d(sbase + s * kStageBytes + kStageA, (kTerms == 3 && term == 2) ? &map_w_lo : &map_w, kk * kBK, co0, full(s));
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_tf32(kMmaM, kBN);
            for (int kb = 0; kb < n_kb; ++kb) {
                const int s = kb % kFcStages;
                mbar_wait(full(s), (kb / kFcStages) & 1);
                tc_fence_after();
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {              // 4 x (K = 8 floats = 32 bytes)
                    const uint64_t da = make_desc_sw128(sbase + s * kStageBytes + kk * 32);
                    const uint64_t da2 = make_desc_sw128(sbase + s * kStageBytes + kMmaM * kBK * 4 + kk * 32);   // rows 128-255
                    const uint64_t db = make_desc_sw128(sbase + s * kStageBytes + kStageA + kk * 32);
                    tc_mma_tf32(tmem, da, db, idesc, (kb | kk) != 0);
                    tc_mma_tf32(tmem + kBN, da2, db, idesc, (kb | kk) != 0);
                }
                tc_commit(empty(s));
            }
            tc_commit(acc_full);
        }
    } else {
        const int q = warp & 3;                               // TMEM lane quadrant this warp may read
        mbar_wait(acc_full, 0);
        tc_fence_after();
#pragma unroll 1
      for (int half = 0; half < kBM / kMmaM; ++half) {        // accumulator of rows [0,128) then of rows [128,256)
        const int row = row0 + half * kMmaM + q * 32 + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(q * 32) << 16) + half * kBN;
        float dot = 0.f;
#pragma unroll 1
        for (int g = 0; g < kBN / 64; ++g) {
            uint32_t v0[32], v1[32];
            tc_ld32(lane_addr + g * 64, v0);
            tc_ld32(lane_addr + g * 64 + 32, v1);
            tc_wait_ld();
#pragma unroll
            for (int hlf = 0; hlf < 2; ++hlf) {
                const uint32_t(&v)[32] = hlf ? v1 : v0;
                const int c0 = co0 + g * 64 + hlf * 32;
#pragma unroll
                for (int c = 0; c < 32; c += 4) {
                    const float4 b = __ldg(reinterpret_cast<const float4*>(bias + c0 + c));
                    float4 o;
                    o.x = fmaxf(__uint_as_float(v[c + 0]) + b.x, 0.f); o.y = fmaxf(__uint_as_float(v[c + 1]) + b.y, 0.f);
                    o.z = fmaxf(__uint_as_float(v[c + 2]) + b.z, 0.f); o.w = fmaxf(__uint_as_float(v[c + 3]) + b.w, 0.f);
                    if (kFuseOut) {
                        const float4 f = __ldg(reinterpret_cast<const float4*>(f3 + c0 + c));
                        dot = fmaf(o.x, f.x, dot); dot = fmaf(o.y, f.y, dot);
                        dot = fmaf(o.z, f.z, dot); dot = fmaf(o.w, f.w, dot);
                    } else if (row < n) {
                        *reinterpret_cast<float4*>(out + (size_t)row * CO + c0 + c) = o;
                        if (kTerms == 3 && out_lo)
                            *reinterpret_cast<float4*>(out_lo + (size_t)row * CO + c0 + c) =
                                make_float4(tf32_lo(o.x), tf32_lo(o.y), tf32_lo(o.z), tf32_lo(o.w));
                    }
                }
            }
        }
        if (kFuseOut && row < n) out[row] = dot + __ldg(c3);
      }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

int get_encoder(zs_ctx* ctx) {
    if (g_encode) return ZS_OK;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
        qres != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return zs_fail(ctx, ZS_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    }
    g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    return ZS_OK;
}

// fp32 row-major [rows][K] matrix, box = [box_rows][32 floats], 128-byte swizzle, OOB rows read as zero
int make_map(zs_ctx* ctx, CUtensorMap* map, const float* base, int rows, int K, int box_rows) {
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return zs_fail(ctx, ZS_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows %d K %d", (int)r, rows, K);
    return ZS_OK;
}

__global__ void zs_k_tf32_lo(const float4* __restrict__ in, size_t n4, float4* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(in + i);
        out[i] = make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
    }
}

}  // namespace

// lo(F1), lo(F2) for the 3-term head, built with the weights
int zs_head_prepare_weights(zs_ctx* ctx, int slot, cudaStream_t st) {
    zs_weights& w = ctx->w[slot];
    const size_t n = (size_t)512 * 1024 + (size_t)256 * 512;
    if (!w.head_lo && cudaMalloc(&w.head_lo, n * sizeof(float)) != cudaSuccess) {
        cudaGetLastError();
        return zs_fail(ctx, ZS_ERR_NOMEM, "head weight remainders");
    }
    zs_k_tf32_lo<<<256, 256, 0, st>>>(reinterpret_cast<const float4*>(w.f32 + ZS_OFF_F1), (size_t)512 * 1024 / 4,
                                      reinterpret_cast<float4*>(w.head_lo));
    ZS_LAUNCHED(ctx);
    zs_k_tf32_lo<<<128, 256, 0, st>>>(reinterpret_cast<const float4*>(w.f32 + ZS_OFF_F2), (size_t)256 * 512 / 4,
                                      reinterpret_cast<float4*>(w.head_lo + (size_t)512 * 1024));
    ZS_LAUNCHED(ctx);
    return ZS_OK;
}

// pooled [n][1024] -> scores [n]; g1 [n][512] scratch.  accurate: 3-term tf32 products (fp32-level result); then
// lo_ws [n][1024 + 512] scratch for lo(pooled) and lo(g1).
int zs_head_tc(zs_ctx* ctx, int slot, const float* pooled, int n, float* scores, float* g1, bool accurate, float* lo_ws,
               cudaStream_t st) {
    int rc = get_encoder(ctx);
    if (rc) return rc;
    const zs_weights& w = ctx->w[slot];
    CUtensorMap a1, w1, a2, w2, a1l, w1l, a2l, w2l;
    if ((rc = make_map(ctx, &a1, pooled, n, 1024, kBM)) || (rc = make_map(ctx, &w1, w.f32 + ZS_OFF_F1, 512, 1024, kBN)) ||
        (rc = make_map(ctx, &a2, g1, n, 512, kBM)) || (rc = make_map(ctx, &w2, w.f32 + ZS_OFF_F2, 256, 512, kBN)))
        return rc;
    a1l = a1; w1l = w1; a2l = a2; w2l = w2;
    float* g1_lo = nullptr;
    if (accurate) {
        float* p_lo = lo_ws;
        g1_lo = lo_ws + (size_t)n * 1024;
        if ((rc = make_map(ctx, &a1l, p_lo, n, 1024, kBM)) || (rc = make_map(ctx, &w1l, w.head_lo, 512, 1024, kBN)) ||
            (rc = make_map(ctx, &a2l, g1_lo, n, 512, kBM)) ||
            (rc = make_map(ctx, &w2l, w.head_lo + (size_t)512 * 1024, 256, 512, kBN)))
            return rc;
        const size_t n4 = (size_t)n * 1024 / 4;
        zs_k_tf32_lo<<<(unsigned)((n4 + 255) / 256 < 4096 ? (n4 + 255) / 256 : 4096), 256, 0, st>>>(
            reinterpret_cast<const float4*>(pooled), n4, reinterpret_cast<float4*>(p_lo));
        ZS_LAUNCHED(ctx);
    }
    const int row_tiles = (n + kBM - 1) / kBM;
#define ZS_LAUNCH_HEAD(TERMS)                                                                                                  \
    do {                                                                                                                       \
        ZS_CUDA(ctx, cudaFuncSetAttribute(zs_k_fc_tc<false, TERMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFcSmAlloc)); \
        ZS_CUDA(ctx, cudaFuncSetAttribute(zs_k_fc_tc<true, TERMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFcSmAlloc));  \
        zs_k_fc_tc<false, TERMS><<<dim3(512 / kBN, row_tiles), kThreadsFc, kFcSmAlloc, st>>>(                                   \
            a1, w1, a1l, w1l, w.f32 + ZS_OFF_C1, g1, g1_lo, n, 1024, 512, nullptr, nullptr);                                    \
        ZS_LAUNCHED(ctx);                                                                                                      \
        zs_k_fc_tc<true, TERMS><<<dim3(256 / kBN, row_tiles), kThreadsFc, kFcSmAlloc, st>>>(                                    \
            a2, w2, a2l, w2l, w.f32 + ZS_OFF_C2, scores, nullptr, n, 512, 256, w.f32 + ZS_OFF_F3, w.f32 + ZS_OFF_C3);           \
        ZS_LAUNCHED(ctx);                                                                                                      \
    } while (0)
    if (accurate) ZS_LAUNCH_HEAD(3); else ZS_LAUNCH_HEAD(1);
#undef ZS_LAUNCH_HEAD
    return ZS_OK;
}
