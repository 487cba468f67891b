// Batched pose errors for all hypotheses of one object against a ground-truth pose.
// The reference evaluates them one hypothesis at a time in a Python list comprehension,
//   pp_err = [err_func(mat[:3,:3], mat[:3,3], mat_gt[:3,:3], mat_gt[:3,3], model_points) for mat in poses_all]
// (python/ossid/scripts/online_learning.py:452, err_func = add | adi from the un-vendored zephyr.utils.metrics,
// :32,337-339), i.e. the BOP definitions:
//   ADD = mean_p || (R p + t) - (Rg p + tg) ||
//   ADI = mean_p min_q || (R p + t) - (Rg q + tg) ||        (symmetric objects; nearest neighbour)
// One CTA per hypothesis; the ground-truth-transformed cloud sits in shared memory and is read as broadcasts.
#include "zs_common.cuh"

namespace {

constexpr int kThreadsMet = 256;

template <bool kAdi>
__global__ void __launch_bounds__(kThreadsMet)
zs_k_pose_errors(const float* __restrict__ poses, int n, const float* __restrict__ gt, const float* __restrict__ pts,
                 int n_pts, float* __restrict__ err_out) {
    extern __shared__ __align__(16) float sm[];
    float* gx = sm;                 // ground-truth transformed points, SoA
    float* gy = gx + n_pts;
    float* gz = gy + n_pts;
    float* px = gz + n_pts;         // raw model points
    float* py = px + n_pts;
    float* pz = py + n_pts;
    __shared__ float s_part[kThreadsMet / 32];
    for (int i = threadIdx.x; i < n_pts; i += blockDim.x) {
        const float x = __ldg(pts + 3 * i), y = __ldg(pts + 3 * i + 1), z = __ldg(pts + 3 * i + 2);
        px[i] = x; py[i] = y; pz[i] = z;
        gx[i] = fmaf(gt[0], x, fmaf(gt[1], y, fmaf(gt[2], z, gt[3])));
        gy[i] = fmaf(gt[4], x, fmaf(gt[5], y, fmaf(gt[6], z, gt[7])));
        gz[i] = fmaf(gt[8], x, fmaf(gt[9], y, fmaf(gt[10], z, gt[11])));
    }
    __syncthreads();
    for (int h = blockIdx.x; h < n; h += gridDim.x) {
        const zs_pose T = zs_load_pose(poses, h);
        float acc = 0.f;
        for (int i = threadIdx.x; i < n_pts; i += blockDim.x) {
            const float x = fmaf(T.r[0], px[i], fmaf(T.r[1], py[i], fmaf(T.r[2], pz[i], T.r[3])));
            const float y = fmaf(T.r[4], px[i], fmaf(T.r[5], py[i], fmaf(T.r[6], pz[i], T.r[7])));
            const float z = fmaf(T.r[8], px[i], fmaf(T.r[9], py[i], fmaf(T.r[10], pz[i], T.r[11])));
            float d2;
            if (kAdi) {
                d2 = INFINITY;
#pragma unroll 4
                for (int q = 0; q < n_pts; ++q) {
                    const float dx = x - gx[q], dy = y - gy[q], dz = z - gz[q];
                    d2 = fminf(d2, fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
                }
            } else {
                const float dx = x - gx[i], dy = y - gy[i], dz = z - gz[i];
                d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
            }
            acc += sqrtf(d2);
        }
        for (int d = 16; d; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
        if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            float s = 0.f;
            for (int w = 0; w < kThreadsMet / 32; ++w) s += s_part[w];
            err_out[h] = s / (float)n_pts;
        }
        __syncthreads();
    }
}

}  // namespace

extern "C" int zs_pose_errors(zs_ctx* ctx, const float* poses, int n, const float* gt_pose, const float* pts, int n_pts,
                              int symmetric, float* err_out, void* stream) {
    if (!ctx) return ZS_ERR_INVALID;
    if (n == 0) return ZS_OK;
    if (n < 0 || n_pts <= 0 || !poses || ((uintptr_t)poses & 15) || !gt_pose || !pts || !err_out)
        return zs_fail(ctx, ZS_ERR_INVALID, "zs_pose_errors arguments");
    const size_t smem = (size_t)n_pts * 6 * sizeof(float);
    if (smem > 200 * 1024) return zs_fail(ctx, ZS_ERR_UNSUPPORTED, "%d model points", n_pts);
    ZS_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = n < ctx->sm_count * 4 ? n : ctx->sm_count * 4;
    if (symmetric) {
        if (smem > 48 * 1024)
            ZS_CUDA(ctx, cudaFuncSetAttribute(zs_k_pose_errors<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        zs_k_pose_errors<true><<<grid, kThreadsMet, smem, st>>>(poses, n, gt_pose, pts, n_pts, err_out);
    } else {
        if (smem > 48 * 1024)
            ZS_CUDA(ctx, cudaFuncSetAttribute(zs_k_pose_errors<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        zs_k_pose_errors<false><<<grid, kThreadsMet, smem, st>>>(poses, n, gt_pose, pts, n_pts, err_out);
    }
    ZS_LAUNCHED(ctx);
    return ZS_OK;
}
