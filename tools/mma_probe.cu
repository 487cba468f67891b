// Micro-benchmark: cycles per tcgen05.mma (kind::f16, bf16 operands from shared memory, K=16) as a function of
// N, of the number of accumulators the instruction stream alternates between, and of cta_group (1 or 2).
// Mirrors layer 3 of zs_k_mlp_tc: "chunks" of 8 K-steps into one accumulator, operands 128B-swizzled K-major.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/mma_probe tools/mma_probe.cu && gpurun_out/mma_probe
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int CG>
__device__ __forceinline__ void tc_mma(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    if (CG == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
template <int CG>
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    if (CG == 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    else
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
}

constexpr uint32_t kSmA = 0, kSmB = 131072, kSmBar = kSmB + 65536, kSmTmem = kSmBar + 64, kSmTotal = kSmTmem + 16 + 1024;

// mode bit0: two accumulators alternate per MMA (instead of one chunk after the other)
template <int CG, int NACC>
__global__ void __launch_bounds__(128, 1) probe(int N, int chunks, int col0, int col_stride, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (sbase - smem_u32(smem_raw));
    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t rank = 0;
    if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    // small pseudo-random bf16 operands
    for (uint32_t i = tid; i < (kSmBar) / 4; i += 128) {
        uint32_t h = (i * 2654435761u) ^ (blockIdx.x * 40503u);
        uint32_t lo = 0x3c00u | (h & 0x80ffu), hi = 0x3c00u | ((h >> 16) & 0x80ffu);   // |x| ~ 0.0078..0.0156
        reinterpret_cast<uint32_t*>(sm)[i] = lo | (hi << 16);
    }
    if (tid == 0) { mbar_init(sbase + kSmBar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        if (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + kSmTmem), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + kSmTmem), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(sm + kSmTmem);

    long long t0 = 0, t1 = 0;
    if (tid == 0 && rank == 0) {
        const uint32_t idesc = make_idesc(128 * CG, N);
        // descriptors differ only in the 14-bit start-address field: add the slab offset to the low word
        const uint64_t da0 = make_desc(sbase + kSmA, 16, 1024, 2), db0 = make_desc(sbase + kSmB, 16, 1024, 2);
        const uint32_t d0 = tmem + col0, d1 = tmem + col0 + col_stride;
        t0 = clock64();
        for (int c = 0; c < chunks; c += 4) {
#pragma unroll
            for (int cc = 0; cc < 4; cc += NACC) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
#pragma unroll
                    for (int a = 0; a < NACC; ++a) {
                        const uint32_t kb = k >> 2, kk = k & 3, cb = cc + a;
                        tc_mma<CG>(((cc + a) & 1) ? d1 : d0, da0 + (((cb * 2 + kb) * 16384 + kk * 32) >> 4),
                                   db0 + ((kb * 32768 + kk * 32) >> 4), idesc, k > 0);
                    }
                }
            }
        }
        tc_commit<CG>(sbase + kSmBar);
    }
    if (tid == 0) {
        mbar_wait(sbase + kSmBar, 0);
        t1 = clock64();
        if (rank == 0) out[blockIdx.x / CG] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

template <int CG, int NACC>
void run(const char* name, int N, int col0, int col_stride, int chunks, long long* d_out) {
    const int n_acc = NACC;
    auto k = probe<CG, NACC>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmTotal));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148, 1, 1);
    cfg.blockDim = dim3(128, 1, 1);
    cfg.dynamicSmemBytes = kSmTotal;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaLaunchKernelEx(&cfg, k, N, chunks / 8, col0, col_stride, d_out));   // warm-up
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    CK(cudaLaunchKernelEx(&cfg, k, N, chunks, col0, col_stride, d_out));
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    long long h[148];
    CK(cudaMemcpy(h, d_out, sizeof(long long) * (148 / CG), cudaMemcpyDeviceToHost));
    double avg = 0;
    for (int i = 0; i < 148 / CG; ++i) avg += (double)h[i];
    avg /= (148 / CG);
    const double n_mma = (double)chunks * 8;
    const double macs_per_sm = n_mma * 128.0 * N * 16.0;     // per SM (each CTA of a pair does 128 x N x 16)
    const double tflops = 2.0 * macs_per_sm * 148 / (ms * 1e-3) / 1e12;
    printf("%-44s N=%3d acc=%d  %7.1f cyc/mma  %6.1f MACs/cyc/SM  %8.3f ms  eff.clock %5.0f MHz  %7.1f TFLOP/s\n", name, N, n_acc,
           avg / n_mma, macs_per_sm / avg, ms, avg / (ms * 1e-3) / 1e6, tflops);
}

int main() {
    long long* d_out;
    CK(cudaMalloc(&d_out, sizeof(long long) * 148));
    const int chunks = 1 << 15;   // 262,144 MMAs per SM
    run<1, 1>("cg1 one accumulator per chunk", 128, 128, 128, chunks, d_out);
    run<1, 2>("cg1 two chunks interleaved", 128, 128, 128, chunks, d_out);
    run<1, 1>("cg1 one accumulator per chunk", 64, 128, 128, chunks, d_out);
    run<1, 2>("cg1 two chunks interleaved", 64, 128, 128, chunks, d_out);
    run<1, 1>("cg1 one accumulator per chunk", 32, 128, 128, chunks, d_out);
    run<1, 1>("cg1 one accumulator per chunk", 256, 0, 256, chunks / 2, d_out);
    run<1, 2>("cg1 two chunks interleaved", 256, 0, 256, chunks / 2, d_out);
    run<2, 1>("cg2 M=256 one accumulator per chunk", 128, 128, 128, chunks, d_out);
    run<2, 2>("cg2 M=256 two chunks interleaved", 128, 128, 128, chunks, d_out);
    run<2, 1>("cg2 M=256 one accumulator per chunk", 256, 0, 256, chunks / 2, d_out);
    run<2, 2>("cg2 M=256 two chunks interleaved", 256, 0, 256, chunks / 2, d_out);
    run<2, 1>("cg2 M=256 one acc, cols 64/272", 208, 64, 208, chunks / 2, d_out);
    run<2, 1>("cg2 M=256 one acc, cols 64/288", 224, 64, 224, chunks / 2, d_out);
    run<2, 1>("cg2 M=256 one acc, cols 128/320", 192, 128, 192, chunks / 2, d_out);
    run<2, 1>("cg2 M=256 one acc", 160, 128, 160, chunks / 2, d_out);
    printf("done\n");
    return 0;
}
