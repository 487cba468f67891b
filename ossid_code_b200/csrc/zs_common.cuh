// Shared internals of libzs.so: context, error plumbing, exact-arithmetic device helpers.
// Public ABI: include/zs.h.
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/zs.h"

#define ZS_SCORE_CHUNK 131072   // hypotheses per scoring chunk (bounds the head's workspace: 13 KB per hypothesis)
// scratch floats per hypothesis of a chunk: pooled 1024 | g1 512 | g2 256 | lo(pooled) 1024 | lo(g1) 512
#define ZS_HEAD_WS_FLOATS (1024 + 512 + 256 + 1024 + 512)
#define ZS_VIOL_MASKED 0x7fffffff   // zs_violations: the hypothesis failed the mask-overlap test (zs_filter never keeps it)
#define ZS_DEPTH_MARGIN 0.02f   // metres; reference: python/ossid/datasets/ycbv_sift_dataset.py:325

struct zs_frame {
    float4* packed = nullptr;   // H*W x {depth/camera_scale, H, S, V}
    size_t cap_px = 0;
    int H = 0, W = 0;
    float fx = 0, fy = 0, cx = 0, cy = 0, inv_fx = 0, inv_fy = 0;
    bool set = false;
};

struct zs_object {
    // {px,py,pz,Hm}, {nx,ny,nz,Sm}, Vm  -- 36 B per model point
    float4* pA = nullptr;
    float4* pB = nullptr;
    float* pV = nullptr;
    int n_pts = 0, cap = 0;
};

struct zs_weights {
    float* f32 = nullptr;            // ZS_WEIGHT_FLOATS, layout of zs_set_weights
    float* f32t = nullptr;           // transposed ([K][CO]) copies of W1 W2 W3 F1 F2 (zs_score_f32.cu)
    __nv_bfloat16* bf16 = nullptr;   // tensor-core operand images of W1..W3 (see zs_score_tc.cu)
    __nv_bfloat16* bf16x2 = nullptr; // the same as bf16 hi + lo pairs for the fp32-accurate kernel (zs_score_tc3.cu)
    float* head_lo = nullptr;        // tf32 remainders of F1, F2 for the 3-term tensor-core head (zs_head_tc.cu)
    bool set = false;
};

struct zs_ctx {
    int device = 0;
    int sm_count = 148;
    int64_t launches = 0;
    int64_t alloc_gen = 0;           // bumped whenever a context-owned device buffer moves (frame, cloud, scratch)
    char err[512] = {0};
    zs_frame frame;
    zs_object obj[ZS_MAX_OBJECTS];
    zs_weights w[ZS_MAX_WEIGHT_SLOTS];
    float* lut255 = nullptr;         // i/255 as fp64 division rounded to fp32, i = 0..255 (zephyr_utils.py:14)
    void* ws = nullptr;              // scratch (pooled vectors and the head's hidden layers)
    size_t ws_bytes = 0;
    void* tc_state = nullptr;        // owned by zs_score_tc.cu (tensor maps etc.)
    const int32_t* dyn_n = nullptr;  // zs_set_dynamic_count: device-side hypothesis count for zs_features / zs_pool
    int dyn_off = 0;
};

// offsets (in floats) into the weight blob
enum : int {
    ZS_OFF_W1 = 0, ZS_OFF_B1 = ZS_OFF_W1 + 64 * 8,
    ZS_OFF_W2 = ZS_OFF_B1 + 64, ZS_OFF_B2 = ZS_OFF_W2 + 128 * 64,
    ZS_OFF_W3 = ZS_OFF_B2 + 128, ZS_OFF_B3 = ZS_OFF_W3 + 1024 * 128,
    ZS_OFF_F1 = ZS_OFF_B3 + 1024, ZS_OFF_C1 = ZS_OFF_F1 + 512 * 1024,
    ZS_OFF_F2 = ZS_OFF_C1 + 512, ZS_OFF_C2 = ZS_OFF_F2 + 256 * 512,
    ZS_OFF_F3 = ZS_OFF_C2 + 256, ZS_OFF_C3 = ZS_OFF_F3 + 256,
    ZS_OFF_END = ZS_OFF_C3 + 1
};
static_assert(ZS_OFF_END == ZS_WEIGHT_FLOATS, "weight blob layout");

int zs_fail(zs_ctx* ctx, int code, const char* fmt, ...);
int zs_reserve_ws(zs_ctx* ctx, size_t bytes);
int zs_tc_prepare_weights(zs_ctx* ctx, int slot, cudaStream_t st);   // zs_score_tc.cu
void zs_tc_destroy(zs_ctx* ctx);
int zs_score_tc(zs_ctx* ctx, int slot, const __nv_bfloat16* feat, int n, int n_pts, float* pooled, cudaStream_t st);
int zs_tc3_prepare_weights(zs_ctx* ctx, int slot, cudaStream_t st);  // zs_score_tc3.cu
int zs_score_tc3(zs_ctx* ctx, int slot, const void* feat_split, int n, int n_pts, float* pooled, float* dbg_h1, float* dbg_h2,
                 cudaStream_t st);
int zs_head_tc(zs_ctx* ctx, int slot, const float* pooled, int n, float* scores, float* g1, bool accurate, float* lo_ws,
               cudaStream_t st);                                             // zs_head_tc.cu
int zs_head_prepare_weights(zs_ctx* ctx, int slot, cudaStream_t st);
int zs_f32_prepare_weights(zs_ctx* ctx, int slot, cudaStream_t st);  // zs_score_f32.cu

#define ZS_CUDA(ctx, call)                                                                      \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess)                                                                  \
            return zs_fail((ctx), ZS_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call,      \
                           cudaGetErrorString(e_));                                             \
    } while (0)

#define ZS_LAUNCHED(ctx)                                                                        \
    do {                                                                                        \
        (ctx)->launches++;                                                                      \
        cudaError_t e_ = cudaGetLastError();                                                    \
        if (e_ != cudaSuccess)                                                                  \
            return zs_fail((ctx), ZS_ERR_CUDA, "%s:%d launch: %s", __FILE__, __LINE__,         \
                           cudaGetErrorString(e_));                                             \
    } while (0)

// Effective hypothesis count of a launch whose count lives on the device (zs_set_dynamic_count): entries
// [n_off, n_off + n_cap) of a list of *n_dev.
__device__ __forceinline__ int zs_dyn_count(const int32_t* __restrict__ n_dev, int n_off, int n_cap) {
    if (!n_dev) return n_cap;
    const int m = __ldg(n_dev) - n_off;
    return m < 0 ? 0 : (m < n_cap ? m : n_cap);
}

// ---------------------------------------------------------------------------------------
// Exact arithmetic.  The oracle performs one IEEE fp32 operation per step in a fixed order;
// the intrinsics below are never contracted into FMAs by nvcc, whatever -fmad says.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float xmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float xadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float xsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float xdiv(float a, float b) { return __fdiv_rn(a, b); }

// ((r0*px + r1*py) + r2*pz) [+ t]
__device__ __forceinline__ float xdot3(float r0, float r1, float r2, float px, float py, float pz) {
    return xadd(xadd(xmul(r0, px), xmul(r1, py)), xmul(r2, pz));
}

// Hexcone RGB -> HSV, H in [0,1).  Same operation order as oracle/zephyr_oracle.py:rgb_to_hsv.
__device__ __forceinline__ void zs_rgb_to_hsv(float r, float g, float b, float& h, float& s, float& v) {
    float mx = fmaxf(fmaxf(r, g), b);
    float mn = fminf(fminf(r, g), b);
    float df = xsub(mx, mn);
    float dfs = df > 0.f ? df : 1.f;
    float hh;
    if (mx == r) {
        hh = xdiv(xsub(g, b), dfs);
        if (hh < 0.f) hh = xadd(hh, 6.f);
    } else if (mx == g) {
        hh = xadd(xdiv(xsub(b, r), dfs), 2.f);
    } else {
        hh = xadd(xdiv(xsub(r, g), dfs), 4.f);
    }
    h = df > 0.f ? xdiv(hh, 6.f) : 0.f;
    s = mx > 0.f ? xdiv(df, mx) : 0.f;
    v = mx;
}

struct zs_cam {
    float fx, fy, cx, cy, inv_fx, inv_fy;
    int H, W;
};

struct zs_pose {
    float r[12];
};

__device__ __forceinline__ zs_pose zs_load_pose(const float* __restrict__ poses, int h) {
    zs_pose p;
    const float4* q = reinterpret_cast<const float4*>(poses + (size_t)h * 12);
    float4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    p.r[0] = a.x; p.r[1] = a.y; p.r[2] = a.z; p.r[3] = a.w;
    p.r[4] = b.x; p.r[5] = b.y; p.r[6] = b.z; p.r[7] = b.w;
    p.r[8] = c.x; p.r[9] = c.y; p.r[10] = c.z; p.r[11] = c.w;
    return p;
}

// R.p + t with the oracle's association.
__device__ __forceinline__ void zs_transform(const zs_pose& T, float px, float py, float pz,
                                             float& x, float& y, float& z) {
    x = xadd(xdot3(T.r[0], T.r[1], T.r[2], px, py, pz), T.r[3]);
    y = xadd(xdot3(T.r[4], T.r[5], T.r[6], px, py, pz), T.r[7]);
    z = xadd(xdot3(T.r[8], T.r[9], T.r[10], px, py, pz), T.r[11]);
}

// Rounded (float-valued) pixel coordinates, (x/z)*f + c, round-half-even.
__device__ __forceinline__ void zs_project(const zs_cam& cam, float x, float y, float z, float& ur, float& vr) {
    ur = rintf(xadd(xmul(xdiv(x, z), cam.fx), cam.cx));
    vr = rintf(xadd(xmul(xdiv(y, z), cam.fy), cam.cy));
}

// ---------------------------------------------------------------------------------------
// Per-point featurisation of the hot path, shared by the feature kernels (zs_features.cu) and by the producer warps of
// the fused tensor-core scorer (zs_score_tc.cu) so that both write bit-identical rows.
// Phase 1 (exact arithmetic): projection, validity, the pixel to gather.  Phase 2: residual features
// [u_n, v_n, dH, dS, dV, dD, ncos]; the cosine may use FMA / MUFU (a 1e-4 feature).
// ---------------------------------------------------------------------------------------
#define ZS_FLT_MAX 3.402823466e38f    // 0 < z <= FLT_MAX, i.e. finite (oracle: z < inf)

struct zs_obj_view {
    const float4* pA;   // {px, py, pz, Hm}
    const float4* pB;   // {nx, ny, nz, Sm}
    const float* pV;    // Vm
    int n_pts;
};

__device__ __forceinline__ float zs_rsqrt_fast(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// -> camera-frame point, rounded pixel (0,0 when invalid), validity, linear pixel index of the gather
__device__ __forceinline__ void zs_feat_phase1(const zs_pose& T, const zs_cam& cam, const float4& a, float& x, float& y,
                                               float& z, float& uf, float& vf, bool& valid, int& pix) {
    float ur, vr;
    zs_transform(T, a.x, a.y, a.z, x, y, z);
    zs_project(cam, x, y, z, ur, vr);
    valid = (z > 0.f) && (z <= ZS_FLT_MAX) && (ur >= 0.f) && (ur < (float)cam.W) && (vr >= 0.f) && (vr < (float)cam.H);
    uf = valid ? ur : 0.f;
    vf = valid ? vr : 0.f;
    pix = (int)vf * cam.W + (int)uf;
}

// px = the gathered frame pixel {d_obs, H, S, V}; a, b, vm = the model point; f[0..6] = the seven feature channels
__device__ __forceinline__ void zs_feat_phase2(const zs_pose& T, const zs_cam& cam, const float4& a, const float4& b, float vm,
                                               const float4& px, float x, float y, float z, float uf, float vf,
                                               float (&f)[7]) {
    const float nx = fmaf(T.r[0], b.x, fmaf(T.r[1], b.y, T.r[2] * b.z));
    const float ny = fmaf(T.r[4], b.x, fmaf(T.r[5], b.y, T.r[6] * b.z));
    const float nz = fmaf(T.r[8], b.x, fmaf(T.r[9], b.y, T.r[10] * b.z));
    const float dot = -fmaf(x, nx, fmaf(y, ny, z * nz));
    const bool vd = (px.x > 0.f) && (px.x <= ZS_FLT_MAX);
    float dH = px.y - a.w;
    dH = dH > 0.5f ? dH - 1.0f : dH;
    dH = dH < -0.5f ? dH + 1.0f : dH;
    f[0] = (uf - cam.cx) * cam.inv_fx;
    f[1] = (vf - cam.cy) * cam.inv_fy;
    f[2] = dH;
    f[3] = px.z - b.w;
    f[4] = px.w - vm;
    f[5] = vd ? xsub(px.x, z) : 0.f;
    // cos of the angle between the viewing ray and the rotated normal (python/ossid/datasets/ycbv_object.py:74)
    const float c = dot * zs_rsqrt_fast(fmaf(x, x, fmaf(y, y, z * z))) * zs_rsqrt_fast(fmaf(nx, nx, fmaf(ny, ny, nz * nz)));
    f[6] = (fabsf(c) <= ZS_FLT_MAX) ? c : 0.f;      // NaN / inf (degenerate pose) -> 0
}

__device__ __forceinline__ uint32_t zs_pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
}

// the 16-byte bf16 row of a point (zeros when its projection is invalid)
__device__ __forceinline__ uint4 zs_feat_row_bf16(const float (&f)[7], bool valid) {
    uint4 v;
    v.x = zs_pack_bf16x2(f[0], f[1]); v.y = zs_pack_bf16x2(f[2], f[3]);
    v.z = zs_pack_bf16x2(f[4], f[5]); v.w = zs_pack_bf16x2(f[6], 0.f);
    if (!valid) v = make_uint4(0u, 0u, 0u, 0u);
    return v;
}
