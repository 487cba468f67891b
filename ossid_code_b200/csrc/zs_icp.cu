// Post-scoring refinement: batched point-to-point ICP of the winning pose(s) against the depth
// image, and the visibility-mask rule.  Reference call sites (python/ossid/scripts/online_learning.py):
//   :476-479  pred_pose, _ = icpRefinement(depth, uv_original[pred_idx], pred_pose, cam_K, model_points,
//                                          inpaint_depth=False, icp_max_dist=0.01)        (zephyr + Open3D, CPU)
//   :497      pred_mask_visib = estimate_visib_mask_gt(depth, pred_depth, 15/1000.)       (bop_toolkit, CPU)
// Both callees are un-vendored; oracle/icp_oracle.py restates their published algorithms and is what the tests compare
// this file with (parity unpinned, see its header).
//
// ICP: one CTA per pose, so the k candidates of every object of a frame refine in one launch.  Target cloud = depth
// back-projected at the pixels `uv` (pixels without depth become points at 1e30 that never match); source = model
// points under the current estimate.  Per iteration every thread carries source points through a brute-force
// nearest-neighbour scan of the target cloud in shared memory (broadcast reads; 10^6 distance evaluations per iteration
// at 1,000 points), the correspondences' sums are reduced in fp64 and one thread solves the closed-form rigid update
// (Horn's unit-quaternion method: largest eigenvector of a 4x4 symmetric matrix by cyclic Jacobi - always a proper
// rotation, i.e. Umeyama's result with its reflection guard).  Loop and stopping rule are Open3D's.
#include "zs_common.cuh"

namespace {

constexpr int kThreadsIcp = 512;
constexpr int kSums = 17;     // n, sum p (3), sum q (3), sum p q^T (9), sum d^2

// eigen-decomposition of a symmetric 4x4 (cyclic Jacobi); returns the unit eigenvector of the largest eigenvalue
__device__ void largest_eigvec4(double A[4][4], double q[4]) {
    double V[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    for (int sweep = 0; sweep < 24; ++sweep) {
        double off = 0.0;
        for (int i = 0; i < 4; ++i)
            for (int j = i + 1; j < 4; ++j) off += A[i][j] * A[i][j];
        if (off < 1e-30) break;
        for (int p = 0; p < 3; ++p)
            for (int r = p + 1; r < 4; ++r) {
                if (fabs(A[p][r]) < 1e-300) continue;
                const double theta = (A[r][r] - A[p][p]) / (2.0 * A[p][r]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 4; ++k) {       // A <- A J
                    const double akp = A[k][p], akr = A[k][r];
                    A[k][p] = c * akp - s * akr;
                    A[k][r] = s * akp + c * akr;
                }
                for (int k = 0; k < 4; ++k) {       // A <- J^T A
                    const double apk = A[p][k], ark = A[r][k];
                    A[p][k] = c * apk - s * ark;
                    A[r][k] = s * apk + c * ark;
                }
                for (int k = 0; k < 4; ++k) {
                    const double vkp = V[k][p], vkr = V[k][r];
                    V[k][p] = c * vkp - s * vkr;
                    V[k][r] = s * vkp + c * vkr;
                }
            }
    }
    int b = 0;
    for (int i = 1; i < 4; ++i)
        if (A[i][i] > A[b][b]) b = i;
    double nrm = 0.0;
    for (int k = 0; k < 4; ++k) nrm += V[k][b] * V[k][b];
    nrm = 1.0 / sqrt(nrm);
    for (int k = 0; k < 4; ++k) q[k] = V[k][b] * nrm;
}

// rigid update (R | t) minimising sum ||R p + t - q||^2 from the correspondence sums; composed into T (row-major 3x4)
__device__ void rigid_update(const double* S, double* T) {
    const double n = S[0], inv = 1.0 / n;
    const double pm[3] = {S[1] * inv, S[2] * inv, S[3] * inv}, qm[3] = {S[4] * inv, S[5] * inv, S[6] * inv};
    double M[3][3];                                  // M[a][b] = sum (p_a - pm_a)(q_b - qm_b)
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) M[a][b] = S[7 + 3 * a + b] - n * pm[a] * qm[b];
    double N[4][4] = {
        {M[0][0] + M[1][1] + M[2][2], M[1][2] - M[2][1], M[2][0] - M[0][2], M[0][1] - M[1][0]},
        {M[1][2] - M[2][1], M[0][0] - M[1][1] - M[2][2], M[0][1] + M[1][0], M[2][0] + M[0][2]},
        {M[2][0] - M[0][2], M[0][1] + M[1][0], -M[0][0] + M[1][1] - M[2][2], M[1][2] + M[2][1]},
        {M[0][1] - M[1][0], M[2][0] + M[0][2], M[1][2] + M[2][1], -M[0][0] - M[1][1] + M[2][2]}};
    double q[4];
    largest_eigvec4(N, q);
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    const double R[3][3] = {{1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)},
                            {2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)},
                            {2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)}};
    double t[3];
    for (int a = 0; a < 3; ++a) t[a] = qm[a] - (R[a][0] * pm[0] + R[a][1] * pm[1] + R[a][2] * pm[2]);
    double Tn[12];
    for (int a = 0; a < 3; ++a) {
        for (int b = 0; b < 4; ++b)
            Tn[4 * a + b] = R[a][0] * T[b] + R[a][1] * T[4 + b] + R[a][2] * T[8 + b];
        Tn[4 * a + 3] += t[a];
    }
    for (int k = 0; k < 12; ++k) T[k] = Tn[k];
}

__global__ void __launch_bounds__(kThreadsIcp)
zs_k_icp(const float* __restrict__ poses, int n, const float* __restrict__ src, int n_src, const int32_t* __restrict__ uv,
         long long uv_stride, const float* __restrict__ depth, int depth_stride, int H, int W, float fx, float fy, float cx,
         float cy, float max_dist, int max_iter, float rel_fit, float rel_rmse, float* __restrict__ poses_out,
         float* __restrict__ stats_out) {
    extern __shared__ __align__(16) float sm_f[];
    float *sx = sm_f, *sy = sx + n_src, *sz = sy + n_src, *tx = sz + n_src, *ty = tx + n_src, *tz = ty + n_src;
    __shared__ double s_part[kThreadsIcp / 32][kSums];
    __shared__ double s_T[12];
    __shared__ double s_stat[4];      // fitness, rmse, n_corr, stop flag
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const float max_d2 = max_dist * max_dist;

    for (int h = blockIdx.x; h < n; h += gridDim.x) {
        const int32_t* uvh = uv + (size_t)h * uv_stride;
        for (int i = tid; i < n_src; i += kThreadsIcp) {
            sx[i] = __ldg(src + 3 * i); sy[i] = __ldg(src + 3 * i + 1); sz[i] = __ldg(src + 3 * i + 2);
            const int u = uvh[2 * i], v = uvh[2 * i + 1];
            float X = 1e30f, Y = 1e30f, Z = 1e30f;
            if (u >= 0 && u < W && v >= 0 && v < H) {
                const float d = __ldg(depth + ((size_t)v * W + u) * depth_stride);
                if (d > 0.f && d <= 3.402823466e38f) {
                    X = __fdiv_rn(__fmul_rn(__fsub_rn((float)u, cx), d), fx);
                    Y = __fdiv_rn(__fmul_rn(__fsub_rn((float)v, cy), d), fy);
                    Z = d;
                }
            }
            tx[i] = X; ty[i] = Y; tz[i] = Z;
        }
        if (tid < 12) s_T[tid] = (double)poses[(size_t)h * 12 + tid];
        __syncthreads();

        double prev_fit = 0.0, prev_rmse = 0.0;
        int iters = 0;
        for (int ev = 0; ev <= max_iter; ++ev) {          // evaluation 0 = the initial pose; evaluation e > 0 follows update e
            float Tf[12];
            for (int k = 0; k < 12; ++k) Tf[k] = (float)s_T[k];
            double acc[kSums];
            for (int k = 0; k < kSums; ++k) acc[k] = 0.0;
            for (int i = tid; i < n_src; i += kThreadsIcp) {
                const float px = fmaf(Tf[0], sx[i], fmaf(Tf[1], sy[i], fmaf(Tf[2], sz[i], Tf[3])));
                const float py = fmaf(Tf[4], sx[i], fmaf(Tf[5], sy[i], fmaf(Tf[6], sz[i], Tf[7])));
                const float pz = fmaf(Tf[8], sx[i], fmaf(Tf[9], sy[i], fmaf(Tf[10], sz[i], Tf[11])));
                float best = INFINITY;
                int bj = 0;
#pragma unroll 4
                for (int j = 0; j < n_src; ++j) {
                    const float dx = px - tx[j], dy = py - ty[j], dz = pz - tz[j];
                    const float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
                    if (d2 < best) { best = d2; bj = j; }
                }
                if (best <= max_d2) {
                    const double p[3] = {px, py, pz}, q[3] = {tx[bj], ty[bj], tz[bj]};
                    acc[0] += 1.0;
                    for (int a = 0; a < 3; ++a) {
                        acc[1 + a] += p[a];
                        acc[4 + a] += q[a];
                        for (int b = 0; b < 3; ++b) acc[7 + 3 * a + b] += p[a] * q[b];
                    }
                    acc[16] += (double)best;
                }
            }
            for (int k = 0; k < kSums; ++k) {
                double v = acc[k];
                for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
                if (lane == 0) s_part[wid][k] = v;
            }
            __syncthreads();
            if (tid == 0) {
                double S[kSums];
                for (int k = 0; k < kSums; ++k) {
                    S[k] = 0.0;
                    for (int w = 0; w < kThreadsIcp / 32; ++w) S[k] += s_part[w][k];
                }
                const double fit = S[0] / (double)n_src, rmse = S[0] > 0 ? sqrt(S[16] / S[0]) : 0.0;
                bool stop = ev > 0 && fabs(prev_fit - fit) < (double)rel_fit && fabs(prev_rmse - rmse) < (double)rel_rmse;
                if (ev == max_iter || S[0] < 3.0) stop = true;
                s_stat[0] = fit; s_stat[1] = rmse; s_stat[2] = S[0]; s_stat[3] = stop ? 1.0 : 0.0;
                if (!stop) rigid_update(S, s_T);
            }
            __syncthreads();
            prev_fit = s_stat[0];
            prev_rmse = s_stat[1];
            iters = ev;
            if (s_stat[3] != 0.0) break;
        }
        if (tid < 12) poses_out[(size_t)h * 12 + tid] = (float)s_T[tid];
        if (tid == 0 && stats_out) {
            stats_out[4 * h + 0] = (float)s_stat[0];
            stats_out[4 * h + 1] = (float)s_stat[1];
            stats_out[4 * h + 2] = (float)iters;
            stats_out[4 * h + 3] = (float)s_stat[2];
        }
        __syncthreads();
    }
}

__global__ void zs_k_visib_mask(const float* __restrict__ d_test, const float* __restrict__ d_model, size_t n, float delta,
                                int bop18, uint8_t* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float dt = d_test[i], dm = d_model[i];
        const bool near = __fsub_rn(dm, dt) <= delta;
        out[i] = bop18 ? (near && dt > 0.f && dm > 0.f) : ((near || dt == 0.f) && dm > 0.f);
    }
}

}  // namespace

extern "C" int zs_icp_refine(zs_ctx* ctx, const float* poses, int n, const float* src_pts, int n_src, const int32_t* uv,
                             int uv_per_pose, const float* depth, int H, int W, float fx, float fy, float cx, float cy,
                             float max_dist, int max_iter, float* poses_out, float* stats_out, void* stream) {
    if (!ctx) return ZS_ERR_INVALID;
    if (n == 0) return ZS_OK;
    if (n < 0 || n_src <= 0 || !poses || !src_pts || !uv || !poses_out || max_iter < 0 || !(max_dist > 0.f))
        return zs_fail(ctx, ZS_ERR_INVALID, "zs_icp_refine arguments");
    int stride = 1;
    if (!depth) {                                     // the context's resident frame
        if (!ctx->frame.set) return zs_fail(ctx, ZS_ERR_STATE, "frame not set");
        const zs_frame& f = ctx->frame;
        depth = reinterpret_cast<const float*>(f.packed);
        stride = 4;
        H = f.H; W = f.W; fx = f.fx; fy = f.fy; cx = f.cx; cy = f.cy;
    } else if (H <= 0 || W <= 0) {
        return zs_fail(ctx, ZS_ERR_INVALID, "depth image %d x %d", H, W);
    }
    const size_t smem = (size_t)n_src * 6 * sizeof(float);
    if (smem > 160 * 1024) return zs_fail(ctx, ZS_ERR_UNSUPPORTED, "%d model points", n_src);
    ZS_CUDA(ctx, cudaSetDevice(ctx->device));
    if (smem > 32 * 1024)
        ZS_CUDA(ctx, cudaFuncSetAttribute(zs_k_icp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = n < ctx->sm_count * 2 ? n : ctx->sm_count * 2;
    zs_k_icp<<<grid, kThreadsIcp, smem, (cudaStream_t)stream>>>(poses, n, src_pts, n_src, uv,
                                                               uv_per_pose ? (long long)n_src * 2 : 0LL, depth, stride, H, W,
                                                               fx, fy, cx, cy, max_dist, max_iter, 1e-6f, 1e-6f, poses_out,
                                                               stats_out);
    ZS_LAUNCHED(ctx);
    return ZS_OK;
}

extern "C" int zs_visib_mask(zs_ctx* ctx, const float* d_test, const float* d_model, size_t n, float delta, int bop18,
                             uint8_t* mask_out, void* stream) {
    if (!ctx) return ZS_ERR_INVALID;
    if (n == 0) return ZS_OK;
    if (!d_test || !d_model || !mask_out) return zs_fail(ctx, ZS_ERR_INVALID, "zs_visib_mask arguments");
    ZS_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t blocks = (n + 255) / 256;
    const int grid = (int)(blocks < (size_t)ctx->sm_count * 8 ? blocks : (size_t)ctx->sm_count * 8);
    zs_k_visib_mask<<<grid, 256, 0, (cudaStream_t)stream>>>(d_test, d_model, n, delta, bop18, mask_out);
    ZS_LAUNCHED(ctx);
    return ZS_OK;
}
