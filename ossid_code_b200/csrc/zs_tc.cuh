// tcgen05 / TMEM / mbarrier / bulk-copy PTX wrappers shared by the tensor-core scorer kernels (zs_score_tc.cu: bf16,
// zs_score_tc3.cu: fp32-accurate 3-term bf16 split).  sm_100a only.  Descriptor bit layouts follow
// cute::UMMA::SmemDescriptor / InstrDescriptor; the code is hand-written.
#pragma once

#include <cuda.h>

#include "zs_common.cuh"

namespace zs_tc {

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
// The leader's barriers also receive arrivals from the peer CTA.  Waits and arrives keep the default (.acquire /
// .release at .cta scope) semantics, as CUTLASS' ClusterBarrier does for its 2-SM pipelines: an explicit
// .release.cluster arrive cost several hundred cycles per arrival here.
__device__ __forceinline__ uint32_t leader_addr(uint32_t local_addr) {     // same offset in the shared memory of cluster rank 0
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(r) : "r"(local_addr));
    return r;
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// One elected lane of a converged warp (elect.sync): unlike `lane == 0`, the compiler then knows the region is
// single-threaded and issues UTCHMMA / UTCBAR without a per-instruction ELECT + BRA.U.ANY guard loop.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void tc_commit(uint32_t bar) {
    // arrives on the barrier at this offset in BOTH CTAs of the pair once all prior MMAs have completed
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns -> 32 registers per thread
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

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout, version 1).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}
constexpr uint32_t kLayoutNone = 0, kLayoutSw128 = 2;
// The same descriptor split in words: low = start address | leading byte offset, high = stride byte offset | version |
// layout.  desc_at() adds a byte offset to the start-address field (no carry: shared addresses stay below 256 KB).
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) { return ((saddr >> 4) & 0x3fff) | (((lbo_bytes >> 4) & 0x3fff) << 16); }
__host__ __device__ constexpr uint32_t desc_hi(uint32_t sbo_bytes, uint32_t layout) { return ((sbo_bytes >> 4) & 0x3fff) | (1u << 14) | (layout << 29); }
__device__ __forceinline__ uint64_t desc_at(uint32_t lo, uint32_t hi, uint32_t off_bytes) {
    uint64_t d;
    asm("mov.b64 %0, {%1,%2};" : "=l"(d) : "r"(lo + (off_bytes >> 4)), "r"(hi));
    return d;
}

// Instruction descriptor: fp32 accumulate, bf16 x bf16, both operands K-major.
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// byte offset of 16-byte chunk `c` (8 bf16) of row `r` in a [rows x 64 bf16] 128B-swizzled K-major tile
__host__ __device__ __forceinline__ uint32_t sw128_off(uint32_t r, uint32_t c) { return r * 128u + (((c ^ (r & 7u)) & 7u) << 4); }

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float max3(float a, float b, float c) {      // FMNMX3
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
// bf16x2( relu(a0 + b0), relu(a1 + b1) ): one FADD2 + one F2FP.RELU.BF16.PACK_AB for two channels
__device__ __forceinline__ uint32_t bias_relu_pack(uint32_t a0, uint32_t a1, float2 b) {
    uint64_t acc, bias, sum;
    uint32_t r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(acc) : "r"(a0), "r"(a1));
    asm("mov.b64 %0, {%1,%2};" : "=l"(bias) : "f"(b.x), "f"(b.y));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(sum) : "l"(acc), "l"(bias));
    float lo, hi;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(sum));
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// 32 accumulator columns of one point -> 4 swizzled 16-byte chunks of the next layer's operand row
__device__ __forceinline__ void epi_store32(const uint32_t (&v)[32], const float* __restrict__ bias, uint8_t* tile,
                                            uint32_t r, uint32_t chunk0) {
#pragma unroll
    for (int c8 = 0; c8 < 4; ++c8) {
        float4 bA = make_float4(0.f, 0.f, 0.f, 0.f), bB = bA;
        if (!EXP(2)) {
            bA = *reinterpret_cast<const float4*>(bias + c8 * 8);
            bB = *reinterpret_cast<const float4*>(bias + c8 * 8 + 4);
        }
        uint4 o;
        o.x = bias_relu_pack(v[c8 * 8 + 0], v[c8 * 8 + 1], make_float2(bA.x, bA.y));
        o.y = bias_relu_pack(v[c8 * 8 + 2], v[c8 * 8 + 3], make_float2(bA.z, bA.w));
        o.z = bias_relu_pack(v[c8 * 8 + 4], v[c8 * 8 + 5], make_float2(bB.x, bB.y));
        o.w = bias_relu_pack(v[c8 * 8 + 6], v[c8 * 8 + 7], make_float2(bB.z, bB.w));
        if (!EXP(1)) *reinterpret_cast<uint4*>(tile + sw128_off(r, chunk0 + c8)) = o;
    }
}
__device__ __forceinline__ void dbg_dump32(float* __restrict__ dst, const uint32_t (&v)[32], const float* __restrict__ bias) {
    for (int c = 0; c < 32; ++c)
        dst[c] = __bfloat162float(__float2bfloat16_rn(fmaxf(__uint_as_float(v[c]) + bias[c], 0.f)));
}
// running max over 32 columns with four independent FMNMX3 chains
__device__ __forceinline__ void max32(const uint32_t (&v)[32], float (&m)[4]) {
#pragma unroll
    for (int c = 0; c < 32; c += 8) {
        m[0] = max3(m[0], __uint_as_float(v[c + 0]), __uint_as_float(v[c + 1]));
        m[1] = max3(m[1], __uint_as_float(v[c + 2]), __uint_as_float(v[c + 3]));
        m[2] = max3(m[2], __uint_as_float(v[c + 4]), __uint_as_float(v[c + 5]));
        m[3] = max3(m[3], __uint_as_float(v[c + 6]), __uint_as_float(v[c + 7]));
    }
}

}  // namespace zs_tc
