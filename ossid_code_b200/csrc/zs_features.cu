// Kernel 1 family: projection, gather, per-point residual features, masks, violation counts, hypothesis pre-filter
// (free-space test + detection-mask overlap), DTOID box -> mask rasterisation, raw uv projection, mask-overlap count.
// ABI: include/zs.h.
//
// Three feature kernels share the per-point code of zs_common.cuh (zs_feat_phase1 / zs_feat_phase2):
//   zs_k_features_hot / zs_k_features_multi   features only.  Work unit = a 256-point chunk of a hypothesis, owned by
//       one warp; every lane carries two points per iteration (two frame gathers in flight); branch-free body; rows
//       leave as one 16-byte (bf16), 32-byte (fp32, st.global.v8) or 2 x 16-byte (split bf16) store per lane, the warp
//       writing contiguous memory.  _multi walks several objects (segments) in one launch.
//   zs_k_features<.., kAux = true>            features + uv / mask / violation-count side outputs.  A warp owns a whole
//       hypothesis (its violation count is a warp-local sum); fp32 rows are transposed through a per-warp shared-memory
//       staging buffer so that the stores are contiguous.
//   the producer warps of zs_k_mlp_tc<true>   (zs_score_tc.cu) featurise the tile the tensor cores are about to read:
//       the headline path, no feature rows in HBM at all.
// In all of them the pose sits in registers, the model cloud (36 B/point) in shared memory (global / L2 in the fused
// kernel, whose shared memory holds the weights), and the packed frame (16 B/pixel, a few MB, L2-resident) is gathered
// through the read-only path.  One 32-warp CTA per SM; warps walk their units with a grid stride.
//
// Everything that decides an integer or a mask bit uses the non-contractable intrinsics of zs_common.cuh in the
// oracle's order; the remaining float features may use FMA / MUFU.
#include <stdlib.h>

#include "zs_common.cuh"

namespace {

constexpr int kWarpsPerCta = 8;               // default CTA: 8 warps; big model clouds use up to 32 (see cta_shape)
constexpr int kThreads = kWarpsPerCta * 32;
constexpr int kMaxThreads = 1024;
constexpr float kFltMax = 3.402823466e38f;   // 0 < z <= FLT_MAX, i.e. finite (oracle: z < inf)

using obj_view = zs_obj_view;

// bytes of shared memory taken by a staged model cloud, rounded so that what follows is 16-byte aligned
__host__ __device__ __forceinline__ size_t cloud_smem(int n_pts) { return ((size_t)n_pts * 36 + 15) & ~(size_t)15; }

__device__ __forceinline__ void st_cs_f4(float4* p, float4 v) {
    asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_cs_u4(uint4* p, uint4 v) {
    asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
}

// (a, b) -> bf16x2 of the high parts and bf16x2 of the remainders: x = hi + lo to 16 mantissa bits
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
    hi = pack_bf16x2(a, b);
    lo = pack_bf16x2(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
}

// Side-output kernel (zs_k_features<.., true>) only: a lane holds the 32 contiguous bytes of its point's fp32 row; the
// warp's 32 x 32 B go through shared memory and leave as two fully contiguous 512-byte warp stores.  (The hot kernels
// write the row with one 256-bit store per lane instead, see st_f32_row.)  `n_act` = points of this warp-iteration.
__device__ __forceinline__ void store_f32_rows(float4* __restrict__ wbuf, int lane, float4* __restrict__ dst, int n_act,
                                               float4 lo, float4 hi) {
    wbuf[lane * 2] = lo;
    wbuf[lane * 2 + 1] = hi;
    __syncwarp();
    const float4 c0 = wbuf[lane], c1 = wbuf[32 + lane];
    if ((lane >> 1) < n_act) st_cs_f4(dst + lane, c0);
    if (16 + (lane >> 1) < n_act) st_cs_f4(dst + 32 + lane, c1);
    __syncwarp();
}

// Stage the model cloud into shared memory (or leave it in global when it does not fit).
template <bool kSmem>
__device__ __forceinline__ void stage_cloud(const obj_view& o, float4*& sA, float4*& sB, float*& sV, char* smem) {
    if (kSmem) {
        sA = reinterpret_cast<float4*>(smem);
        sB = sA + o.n_pts;
        sV = reinterpret_cast<float*>(sB + o.n_pts);   // the per-warp fp32 store staging follows the cloud (see cloud_smem)
        for (int i = threadIdx.x; i < o.n_pts; i += blockDim.x) {
            sA[i] = __ldg(o.pA + i);
            sB[i] = __ldg(o.pB + i);
            sV[i] = __ldg(o.pV + i);
        }
        __syncthreads();
    } else {
        sA = const_cast<float4*>(o.pA);
        sB = const_cast<float4*>(o.pB);
        sV = const_cast<float*>(o.pV);
    }
}

// ---------------------------------------------------------------------------------------
// Features.  kBf16: feature dtype.  kSmem: model cloud staged in shared memory.
// kAux: some side output (mask / uv / violation count) is requested.  Without side outputs the
// mask bits are never materialised, so the rotated normal and its dot product (which then feed
// only the 1e-4 cosine feature, not the bit-exact front-facing flag) may use FMAs.  Everything
// that selects the gathered pixel or zeroes a feature stays on the exact path in both variants.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float rsqrt_fast(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

template <bool kBf16, bool kSmem, bool kAux>
__global__ void __launch_bounds__(kMaxThreads, 1)
zs_k_features(obj_view o, zs_cam cam, const float4* __restrict__ frame, const float* __restrict__ poses,
              const int32_t* __restrict__ keep_idx, int n_keep, void* __restrict__ feat_out,
              int32_t* __restrict__ uv_out, uint8_t* __restrict__ mask_out, int32_t* __restrict__ viol_out,
              const int32_t* __restrict__ n_dev, int n_off) {
    n_keep = zs_dyn_count(n_dev, n_off, n_keep);
    extern __shared__ __align__(16) char smem[];
    float4 *sA, *sB;
    float* sV;
    stage_cloud<kSmem>(o, sA, sB, sV, smem);

    const int lane = threadIdx.x & 31;
    const int warps_per_cta = blockDim.x >> 5;
    const int warp = blockIdx.x * warps_per_cta + (threadIdx.x >> 5);
    const int n_warps = gridDim.x * warps_per_cta;
    const int N = o.n_pts;
    const float fW = (float)cam.W, fH = (float)cam.H;
    float4* wbuf = reinterpret_cast<float4*>(smem + (kSmem ? cloud_smem(N) : 0)) + (threadIdx.x >> 5) * 64;

    // Work unit: with side outputs a warp owns a whole hypothesis (its violation count is a warp-local
    // sum); without them a unit is a 256-point chunk of a hypothesis, which keeps the last wave of a
    // 10k-hypothesis launch short (units / resident warps ~ 7 instead of ~ 2).
    constexpr int kChunk = 256;
    const int n_chunks = kAux ? 1 : (N + kChunk - 1) / kChunk;
    const long long n_units = (long long)n_keep * n_chunks;
    for (long long u = warp; u < n_units; u += n_warps) {
        const int hk = kAux ? (int)u : (int)(u / n_chunks);
        const int p_begin = kAux ? 0 : (int)(u - (long long)hk * n_chunks) * kChunk;
        const int p_end = kAux ? N : min(N, p_begin + kChunk);
        const int h = keep_idx ? __ldg(keep_idx + hk) : hk;
        const zs_pose T = zs_load_pose(poses, h);
        const size_t row = (size_t)hk * N;
        int viol = 0;
        for (int p0 = p_begin; p0 < p_end; p0 += 32) {
            const int p = p0 + lane;
            const bool act = p < p_end;
            uint32_t mk = 0;
            float f0 = 0.f, f1 = 0.f, f2 = 0.f, f3 = 0.f, f4 = 0.f, f5 = 0.f, f6 = 0.f;
            int ui = 0, vi = 0;
            if (act) {
                const float4 a = kSmem ? sA[p] : __ldg(sA + p);
                const float4 b = kSmem ? sB[p] : __ldg(sB + p);
                float x, y, z, ur, vr;
                zs_transform(T, a.x, a.y, a.z, x, y, z);
                zs_project(cam, x, y, z, ur, vr);
                const bool valid = (z > 0.f) && (z <= kFltMax) && (ur >= 0.f) && (ur < fW) && (vr >= 0.f) && (vr < fH);
                float nx, ny, nz, dot;
                if (kAux) {     // R.n and the dot product with the oracle's association: the front-facing bit is exact
                    nx = xdot3(T.r[0], T.r[1], T.r[2], b.x, b.y, b.z);
                    ny = xdot3(T.r[4], T.r[5], T.r[6], b.x, b.y, b.z);
                    nz = xdot3(T.r[8], T.r[9], T.r[10], b.x, b.y, b.z);
                    dot = -xadd(xadd(xmul(x, nx), xmul(y, ny)), xmul(z, nz));
                    mk = dot > 0.f ? ZS_BIT_FRONT : 0;
                }
                if (valid) {
                    ui = (int)ur;
                    vi = (int)vr;
                    const float4 px = __ldg(frame + (size_t)vi * cam.W + ui);   // {d_obs, H, S, V}
                    const float vm = kSmem ? sV[p] : __ldg(sV + p);
                    if (!kAux) {
                        nx = fmaf(T.r[0], b.x, fmaf(T.r[1], b.y, T.r[2] * b.z));
                        ny = fmaf(T.r[4], b.x, fmaf(T.r[5], b.y, T.r[6] * b.z));
                        nz = fmaf(T.r[8], b.x, fmaf(T.r[9], b.y, T.r[10] * b.z));
                        dot = -fmaf(x, nx, fmaf(y, ny, z * nz));
                    }
                    const bool vd = (px.x > 0.f) && (px.x <= kFltMax);
                    const float dD = vd ? xsub(px.x, z) : 0.f;
                    if (kAux) {
                        mk |= ZS_BIT_VALID_PROJ | (vd ? ZS_BIT_VALID_DEPTH : 0);
                        if (vd && dD > ZS_DEPTH_MARGIN) mk |= ZS_BIT_FREE_SPACE;
                        if (vd && dD < -ZS_DEPTH_MARGIN) mk |= ZS_BIT_OCCLUDED;
                    }
                    float dH = px.y - a.w;
                    dH = dH > 0.5f ? dH - 1.0f : dH;
                    dH = dH < -0.5f ? dH + 1.0f : dH;
                    f0 = ((float)ui - cam.cx) * cam.inv_fx;
                    f1 = ((float)vi - cam.cy) * cam.inv_fy;
                    f2 = dH;
                    f3 = px.z - b.w;
                    f4 = px.w - vm;
                    f5 = dD;
                    // cos of the angle between the viewing ray and the rotated normal
                    // (python/ossid/datasets/ycbv_object.py:74); a 1e-4 feature, so MUFU.RSQ is fine
                    const float c = dot * rsqrt_fast(fmaf(x, x, fmaf(y, y, z * z))) * rsqrt_fast(fmaf(nx, nx, fmaf(ny, ny, nz * nz)));
                    f6 = (fabsf(c) <= kFltMax) ? c : 0.f;      // NaN / inf (degenerate pose) -> 0
                }
            }
            if (kAux) viol += __popc(__ballot_sync(0xffffffffu, (mk & ZS_BIT_FREE_SPACE) != 0));
            if (kBf16) {
                if (act) {
                    uint4 v;
                    v.x = pack_bf16x2(f0, f1); v.y = pack_bf16x2(f2, f3);
                    v.z = pack_bf16x2(f4, f5); v.w = pack_bf16x2(f6, 0.f);
                    st_cs_u4(reinterpret_cast<uint4*>(feat_out) + row + p, v);
                }
            } else {
                store_f32_rows(wbuf, lane, reinterpret_cast<float4*>(feat_out) + (row + p0) * 2, p_end - p0,
                               make_float4(f0, f1, f2, f3), make_float4(f4, f5, f6, 0.f));
            }
            if (act) {
                if (kAux) {
                    if (mask_out) mask_out[row + p] = (uint8_t)mk;
                    if (uv_out) reinterpret_cast<int2*>(uv_out)[row + p] = make_int2(ui, vi);
                }
            }
        }
        if (kAux && viol_out && lane == 0) viol_out[hk] = viol;
    }
}

// ---------------------------------------------------------------------------------------
// Hot variant: features only (no mask / uv / violation outputs).  Work unit = 256-point chunk of a
// hypothesis (short last wave); each lane carries kIlp points per iteration so that kIlp frame
// gathers are in flight per warp.  The body is branch-free: out-of-range lanes of the last
// iteration recompute the chunk's last point and only the store is predicated; a point that does
// not project into the frame gathers pixel 0 and has its features zeroed by a select at the end
// (same values as zs_k_features<.,.,false>, which skips the work instead).  fp32 rows (32 B) leave
// as one 256-bit store per lane: every lane writes one full sector, the warp 1 KB contiguous.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void st_f32_row(float* p, float4 lo, float4 hi, bool aligned32) {
    if (aligned32) {
        asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(lo.x), "f"(lo.y), "f"(lo.z), "f"(lo.w),
                     "f"(hi.x), "f"(hi.y), "f"(hi.z), "f"(hi.w) : "memory");
    } else {
        st_cs_f4(reinterpret_cast<float4*>(p), lo);
        st_cs_f4(reinterpret_cast<float4*>(p) + 1, hi);
    }
}

// kFmt: ZS_F32 (32-byte rows), ZS_BF16 (16-byte rows) or ZS_BF16_SPLIT (per hypothesis a plane of bf16(x) rows followed by
// a plane of bf16(x - bf16(x)) rows: the two K halves of the fp32-accurate scorer's layer-1 operand, zs_score_tc3.cu).
// ---- experiment variant, build with -DZS_CROP_STAGE (tools/README.md): the crop of the packed frame named by the
// environment variable ZS_CROP_RECT="x0,y0,w,h" is staged into shared memory next to the model cloud by the TMA unit
// (one cp.async.bulk per crop row, completion on an mbarrier), and gathers that fall inside it read shared memory;
// the others keep the read-only global path.  This is the "RGB-D crop resident in shared memory" layout that
// BASELINE.json's north_star describes; measured slower than the L2 path, hence not the default.
struct crop_view { const float4* s; int x0, y0, w, h; };
#ifdef ZS_CROP_STAGE
__device__ __forceinline__ float4 gather_px(const float4* __restrict__ frame, const zs_cam& cam, const crop_view& cv, int pix) {
    const int v = pix / cam.W, u = pix - v * cam.W;
    const unsigned du = (unsigned)(u - cv.x0), dv = (unsigned)(v - cv.y0);
    if (du < (unsigned)cv.w && dv < (unsigned)cv.h) return cv.s[dv * cv.w + du];
    return __ldg(frame + pix);
}
__device__ __forceinline__ void stage_crop(const float4* __restrict__ frame, const zs_cam& cam, crop_view& cv, char* smem_crop) {
    __shared__ __align__(8) unsigned long long crop_bar;
    cv.s = reinterpret_cast<const float4*>(smem_crop);
    if (cv.w <= 0 || cv.h <= 0) return;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&crop_bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(cv.w * cv.h * 16)) : "memory");
    __syncthreads();
    for (int r = threadIdx.x; r < cv.h; r += blockDim.x) {          // one bulk copy (TMA unit) per crop row
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_crop + (size_t)r * cv.w * 16);
        const float4* src = frame + (size_t)(cv.y0 + r) * cam.W + cv.x0;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(src), "r"((uint32_t)(cv.w * 16)), "r"(bar) : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tWAITC_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra DONEC_%=;\n\tbra WAITC_%=;\n\tDONEC_%=:\n\t}"
                 ::"r"(bar) : "memory");
}
#endif

// The units of ONE object that this warp owns: first_unit, first_unit + n_warps, ...
template <int kFmt, bool kSmem>
__device__ __forceinline__ void hot_units(const obj_view& o, const zs_cam& cam, const float4* __restrict__ frame,
                                          const float* __restrict__ poses, const int32_t* __restrict__ keep_idx, int n_keep,
                                          void* __restrict__ feat_out, int aligned32, const float4* sA, const float4* sB,
                                          const float* sV, long long first_unit, int n_warps, const crop_view cv = crop_view{}) {
    constexpr int kIlp = 2, kChunk = 256;
    const int lane = threadIdx.x & 31;
    const int N = o.n_pts;
    const int n_chunks = (N + kChunk - 1) / kChunk;
    const long long n_units = (long long)n_keep * n_chunks;
    for (long long u = first_unit; u < n_units; u += n_warps) {
        const int hk = (int)(u / n_chunks);
        const int p_begin = (int)(u - (long long)hk * n_chunks) * kChunk;
        const int p_end = min(N, p_begin + kChunk);
        const int h = keep_idx ? __ldg(keep_idx + hk) : hk;
        const zs_pose T = zs_load_pose(poses, h);
        const size_t row = (size_t)hk * N;
        for (int p0 = p_begin; p0 < p_end; p0 += 32 * kIlp) {
            float4 a[kIlp], px[kIlp];
            float x[kIlp], y[kIlp], z[kIlp], uf[kIlp], vf[kIlp];
            int q[kIlp];
            bool valid[kIlp];
#pragma unroll
            for (int j = 0; j < kIlp; ++j) {                 // phase 1: exact projection, issue the gather
                int pix;
                q[j] = min(p0 + j * 32 + lane, p_end - 1);
                a[j] = kSmem ? sA[q[j]] : __ldg(sA + q[j]);
                zs_feat_phase1(T, cam, a[j], x[j], y[j], z[j], uf[j], vf[j], valid[j], pix);
#ifdef ZS_CROP_STAGE
                px[j] = gather_px(frame, cam, cv, pix);
#else
                px[j] = __ldg(frame + pix);                  // {d_obs, H, S, V}; pixel 0 when invalid
#endif
            }
#pragma unroll
            for (int j = 0; j < kIlp; ++j) {                 // phase 2: residual features, store
                const int p = p0 + j * 32 + lane;
                const float4 b = kSmem ? sB[q[j]] : __ldg(sB + q[j]);
                const float vm = kSmem ? sV[q[j]] : __ldg(sV + q[j]);
                float f[7];
                zs_feat_phase2(T, cam, a[j], b, vm, px[j], x[j], y[j], z[j], uf[j], vf[j], f);
                const float f0 = f[0], f1 = f[1], dH = f[2], f3 = f[3], f4 = f[4], f5 = f[5], f6 = f[6];
                if (kFmt == ZS_BF16) {
                    const uint4 v = zs_feat_row_bf16(f, valid[j]);
                    if (p < p_end) st_cs_u4(reinterpret_cast<uint4*>(feat_out) + row + p, v);
                } else if (kFmt == ZS_BF16_SPLIT) {
                    uint4 h, l;
                    split_bf16x2(f0, f1, h.x, l.x); split_bf16x2(dH, f3, h.y, l.y);
                    split_bf16x2(f4, f5, h.z, l.z); split_bf16x2(f6, 0.f, h.w, l.w);
                    if (!valid[j]) h = l = make_uint4(0u, 0u, 0u, 0u);
                    if (p < p_end) {
                        st_cs_u4(reinterpret_cast<uint4*>(feat_out) + 2 * row + p, h);
                        st_cs_u4(reinterpret_cast<uint4*>(feat_out) + 2 * row + N + p, l);
                    }
                } else {
                    float4 lo = make_float4(f0, f1, dH, f3), hi = make_float4(f4, f5, f6, 0.f);
                    if (!valid[j]) lo = hi = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (p < p_end) st_f32_row(reinterpret_cast<float*>(feat_out) + (row + p) * 8, lo, hi, aligned32 != 0);
                }
            }
        }
    }
}

template <int kFmt, bool kSmem>
__global__ void __launch_bounds__(kMaxThreads, 1)
zs_k_features_hot(obj_view o, zs_cam cam, const float4* __restrict__ frame, const float* __restrict__ poses,
                  const int32_t* __restrict__ keep_idx, int n_keep, void* __restrict__ feat_out, int aligned32,
                  const int32_t* __restrict__ n_dev, int n_off, crop_view cv) {
    n_keep = zs_dyn_count(n_dev, n_off, n_keep);
    extern __shared__ __align__(16) char smem[];
    float4 *sA, *sB;
    float* sV;
    stage_cloud<kSmem>(o, sA, sB, sV, smem);
#ifdef ZS_CROP_STAGE
    stage_crop(frame, cam, cv, smem + (kSmem ? cloud_smem(o.n_pts) : 0));
#endif
    const int warps_per_cta = blockDim.x >> 5;
    hot_units<kFmt, kSmem>(o, cam, frame, poses, keep_idx, n_keep, feat_out, aligned32, sA, sB, sV,
                           blockIdx.x * warps_per_cta + (threadIdx.x >> 5), gridDim.x * warps_per_cta, cv);
}

// Several objects in one launch (a frame's objects share the frame but not the model cloud): every CTA walks the
// segment list, re-stages the cloud of each segment it has units of, and takes its grid-stride share of that
// segment's units; the first unit of segment s goes to CTA s * (grid / segments), so that many short segments (the
// re-rank: k candidates per object) run side by side instead of all starting on CTA 0.
struct feat_seg {
    obj_view o;
    const float* poses;
    const int32_t* keep_idx;
    const int32_t* n_dev;
    void* out;
    int n_keep, n_off, aligned32, pad_;
};
constexpr int kMaxSegsPerLaunch = 32;
struct feat_segs { feat_seg s[kMaxSegsPerLaunch]; };

template <int kFmt, bool kSmem>
__global__ void __launch_bounds__(kMaxThreads, 1)
zs_k_features_multi(const __grid_constant__ feat_segs segs, int n_seg, zs_cam cam, const float4* __restrict__ frame) {
    extern __shared__ __align__(16) char smem[];
    const int warps_per_cta = blockDim.x >> 5;
    const int n_warps = gridDim.x * warps_per_cta;
    const int stride_ctas = max(1, (int)gridDim.x / n_seg);
    for (int sgi = 0; sgi < n_seg; ++sgi) {
        const feat_seg& sg = segs.s[sgi];
        const int n_keep = zs_dyn_count(sg.n_dev, sg.n_off, sg.n_keep);
        const long long n_units = (long long)n_keep * ((sg.o.n_pts + 255) / 256);
        const int cta_first = (int)(((long long)blockIdx.x + gridDim.x - (long long)(sgi * stride_ctas) % gridDim.x) % gridDim.x);
        if ((long long)cta_first * warps_per_cta >= n_units) continue;            // CTA-uniform: nothing of this segment here
        float4 *sA, *sB;
        float* sV;
        __syncthreads();                                                           // previous segment's cloud no longer in use
        stage_cloud<kSmem>(sg.o, sA, sB, sV, smem);
        hot_units<kFmt, kSmem>(sg.o, cam, frame, sg.poses, sg.keep_idx, n_keep, sg.out, sg.aligned32, sA, sB, sV,
                               (long long)cta_first * warps_per_cta + (threadIdx.x >> 5), n_warps);
    }
}

// ---------------------------------------------------------------------------------------
// Violation count only (pre-filter pass): exact part of the feature kernel, depth gather only.
// ---------------------------------------------------------------------------------------
// kMask: the hypothesis must also project at least `mask_min` of its points onto non-zero pixels of `mask`
// (filterHypoByMask, python/ossid/utils/zephyr_utils.py:49-71; bounds predicate :61-62).  The warp stops working on a
// hypothesis as soon as the points still to come cannot lift it over that bar (early-out: a hypothesis far from the
// detector's box is dropped after about half of its projections and without the rest of its depth gathers) and
// reports ZS_VIOL_MASKED instead of a count.
// One hypothesis, one warp: lanes stride over the model points, two points per lane per iteration (two gathers in flight).
template <bool kSmem, bool kMask>
__device__ __forceinline__ int viol_hypothesis(int N, const zs_cam& cam, const float* __restrict__ frame_d, const zs_pose& T,
                                               const float4* sA, const uint8_t* __restrict__ mask, int mask_min, int lane) {
    const float fW = (float)cam.W, fH = (float)cam.H;
    int viol = 0, in_mask = 0;
    for (int p0 = 0; p0 < N; p0 += 64) {
        float z[2], d[2];
        bool need_d[2], im[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int p = p0 + j * 32 + lane;
            need_d[j] = im[j] = false;
            z[j] = d[j] = 0.f;
            if (p < N) {
                const float4 a = kSmem ? sA[p] : __ldg(sA + p);
                float x, y, ur, vr;
                zs_transform(T, a.x, a.y, a.z, x, y, z[j]);
                zs_project(cam, x, y, z[j], ur, vr);
                const bool in_frame = (ur >= 0.f) && (ur < fW) && (vr >= 0.f) && (vr < fH);
                const size_t pix = in_frame ? (size_t)(int)vr * cam.W + (int)ur : 0;
                if (kMask && in_frame) im[j] = __ldg(mask + pix) != 0;                 // no z test: zephyr_utils.py:58-66
                need_d[j] = in_frame && (z[j] > 0.f) && (z[j] <= kFltMax);
                if (need_d[j]) d[j] = __ldg(frame_d + 4 * pix);
            }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const bool fs = need_d[j] && (d[j] > 0.f) && (d[j] <= kFltMax) && (xsub(d[j], z[j]) > ZS_DEPTH_MARGIN);
            viol += __popc(__ballot_sync(0xffffffffu, fs));
            if (kMask) in_mask += __popc(__ballot_sync(0xffffffffu, im[j]));
        }
        if (kMask && in_mask + max(N - (p0 + 64), 0) < mask_min) return ZS_VIOL_MASKED;   // warp-uniform early-out
    }
    return viol;
}

// kMask: the hypothesis must also project at least `mask_min` of its points onto non-zero pixels of `mask`
// (filterHypoByMask, python/ossid/utils/zephyr_utils.py:49-71; bounds predicate :61-62).  The warp stops working on a
// hypothesis as soon as the points still to come cannot lift it over that bar (early-out: a hypothesis far from the
// detector's box is dropped after about half of its projections and without the rest of its depth gathers) and
// reports ZS_VIOL_MASKED instead of a count.
template <bool kSmem, bool kMask>
__global__ void __launch_bounds__(kMaxThreads, 1)
zs_k_violations(obj_view o, zs_cam cam, const float4* __restrict__ frame, const float* __restrict__ poses,
                int n, const uint8_t* __restrict__ mask, int mask_min, int32_t* __restrict__ viol_out) {
    extern __shared__ __align__(16) char smem[];
    float4 *sA, *sB;
    float* sV;
    stage_cloud<kSmem>(o, sA, sB, sV, smem);
    const int lane = threadIdx.x & 31;
    const int warps_per_cta = blockDim.x >> 5;
    const int warp = blockIdx.x * warps_per_cta + (threadIdx.x >> 5);
    const int n_warps = gridDim.x * warps_per_cta;
    for (int h = warp; h < n; h += n_warps) {
        const int v = viol_hypothesis<kSmem, kMask>(o.n_pts, cam, reinterpret_cast<const float*>(frame), zs_load_pose(poses, h),
                                                    sA, mask, mask_min, lane);
        if (lane == 0) viol_out[h] = v;
    }
}

// The pre-filter pass of a whole frame in one launch: segment i = (cloud, poses, count[, mask]) -> viol_out[i].
struct viol_seg {
    obj_view o;
    const float* poses;
    const uint8_t* mask;
    int32_t* viol_out;
    int n, mask_min;
};
struct viol_segs { viol_seg s[kMaxSegsPerLaunch]; };

template <bool kSmem>
__global__ void __launch_bounds__(kMaxThreads, 1)
zs_k_violations_multi(const __grid_constant__ viol_segs segs, int n_seg, zs_cam cam, const float4* __restrict__ frame) {
    extern __shared__ __align__(16) char smem[];
    const int lane = threadIdx.x & 31;
    const int warps_per_cta = blockDim.x >> 5;
    const int n_warps = gridDim.x * warps_per_cta;
    const int stride_ctas = max(1, (int)gridDim.x / n_seg);
    for (int sgi = 0; sgi < n_seg; ++sgi) {
        const viol_seg& sg = segs.s[sgi];
        const int cta_first = (int)(((long long)blockIdx.x + gridDim.x - (long long)(sgi * stride_ctas) % gridDim.x) % gridDim.x);
        if (cta_first * warps_per_cta >= sg.n) continue;                          // CTA-uniform
        float4 *sA, *sB;
        float* sV;
        __syncthreads();                                                           // previous segment's cloud no longer in use
        stage_cloud<kSmem>(sg.o, sA, sB, sV, smem);
        for (int h = cta_first * warps_per_cta + (threadIdx.x >> 5); h < sg.n; h += n_warps) {
            const zs_pose T = zs_load_pose(sg.poses, h);
            const int v = sg.mask ? viol_hypothesis<kSmem, true>(sg.o.n_pts, cam, reinterpret_cast<const float*>(frame), T, sA,
                                                                 sg.mask, sg.mask_min, lane)
                                  : viol_hypothesis<kSmem, false>(sg.o.n_pts, cam, reinterpret_cast<const float*>(frame), T, sA,
                                                                  nullptr, 0, lane);
            if (lane == 0) sg.viol_out[h] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------
// Raw projection (projectPointsUv) and mask-overlap count (filterHypoByMask).  Model points
// come straight from the caller's (n_pts,3) float32 array.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int raw_round(float f) {
    // non-finite or |f| >= 2^30 -> -2^30 (always outside any frame), as oracle project_raw
    return (fabsf(f) < 1073741824.f) ? (int)f : -1073741824;
}

template <bool kCount>
__global__ void __launch_bounds__(kThreads)
zs_k_project(const float* __restrict__ poses, int n, const float* __restrict__ pts, int n_pts, zs_cam cam,
             int32_t* __restrict__ uv_out, const uint8_t* __restrict__ mask, int32_t* __restrict__ count_out) {
    extern __shared__ __align__(16) char smem[];
    float* sp = reinterpret_cast<float*>(smem);
    for (int i = threadIdx.x; i < 3 * n_pts; i += blockDim.x) sp[i] = __ldg(pts + i);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warp = blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    const int n_warps = gridDim.x * kWarpsPerCta;
    for (int h = warp; h < n; h += n_warps) {
        const zs_pose T = zs_load_pose(poses, h);
        int cnt = 0;
        for (int p0 = 0; p0 < n_pts; p0 += 32) {
            const int p = p0 + lane;
            bool in = false;
            if (p < n_pts) {
                float x, y, z, ur, vr;
                zs_transform(T, sp[3 * p], sp[3 * p + 1], sp[3 * p + 2], x, y, z);
                zs_project(cam, x, y, z, ur, vr);
                const int u = raw_round(ur), v = raw_round(vr);
                if (kCount) {
                    // bounds predicate of python/ossid/utils/zephyr_utils.py:61-62
                    const bool invalid = (v >= cam.H) || (v < 0) || (u >= cam.W) || (u < 0);
                    in = !invalid && (__ldg(mask + (size_t)v * cam.W + u) != 0);
                } else {
                    reinterpret_cast<int2*>(uv_out)[(size_t)h * n_pts + p] = make_int2(u, v);
                }
            }
            if (kCount) cnt += __popc(__ballot_sync(0xffffffffu, in));
        }
        if (kCount && lane == 0) count_out[h] = cnt;
    }
}

// ---------------------------------------------------------------------------------------
// Hypothesis pre-filter: stable compaction of {h : viol[h]*100/n_pts < th}; single CTA.
// ---------------------------------------------------------------------------------------
struct filt_seg { const int32_t* viol; int32_t* keep_idx; int32_t* n_keep_out; int32_t* info_out; int n; float n_pts_f; };
struct filt_segs { filt_seg s[ZS_MAX_OBJECTS]; };

__device__ __forceinline__ void filter_body(const int32_t* __restrict__ viol, int n, float n_pts_f, float th,
                                            int32_t* __restrict__ keep_idx, int32_t* __restrict__ n_keep_out,
                                            int32_t* __restrict__ info_out);

__global__ void __launch_bounds__(1024)
zs_k_filter(const int32_t* __restrict__ viol, int n, float n_pts_f, float th, int32_t* __restrict__ keep_idx,
            int32_t* __restrict__ n_keep_out, int32_t* __restrict__ info_out) {
    filter_body(viol, n, n_pts_f, th, keep_idx, n_keep_out, info_out);
}

// one CTA per object of the frame
__global__ void __launch_bounds__(1024)
zs_k_filter_multi(const __grid_constant__ filt_segs segs, float th) {
    const filt_seg& g = segs.s[blockIdx.x];
    filter_body(g.viol, g.n, g.n_pts_f, th, g.keep_idx, g.n_keep_out, g.info_out);
}

__device__ __forceinline__ void filter_body(const int32_t* __restrict__ viol, int n, float n_pts_f, float th,
                                            int32_t* __restrict__ keep_idx, int32_t* __restrict__ n_keep_out,
                                            int32_t* __restrict__ info_out) {
    __shared__ int s_warp[32];
    __shared__ int s_total;
    __shared__ unsigned long long s_min;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_min = ~0ull;
    int base_out = 0;                  // kept so far (same value in every thread)
    unsigned long long best = ~0ull;   // (viol << 32 | index): minimum = first minimum-violation hypothesis
    for (int base = 0; base < n; base += 1024) {
        const int h = base + threadIdx.x;
        bool keep = false;
        if (h < n) {
            const int v = viol[h];
            if (v != ZS_VIOL_MASKED) {         // dropped by the mask test: never kept, not even by the never-empty rule
                keep = (th >= 100.f) || (xdiv(xmul((float)v, 100.f), n_pts_f) < th);
                const unsigned long long key = ((unsigned long long)(uint32_t)v << 32) | (uint32_t)h;
                best = key < best ? key : best;
            }
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, keep);
        __syncthreads();               // previous iteration's readers of s_warp / s_total are done
        if (lane == 0) s_warp[wid] = __popc(bal);
        __syncthreads();
        if (wid == 0) {
            const int c = s_warp[lane];
            int incl = c;
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            s_warp[lane] = incl - c;   // exclusive prefix over warps
            if (lane == 31) s_total = incl;
        }
        __syncthreads();
        if (keep) keep_idx[base_out + s_warp[wid] + __popc(bal & ((1u << lane) - 1u))] = h;
        base_out += s_total;
    }
    for (int d = 16; d; d >>= 1) {
        const unsigned long long t = __shfl_xor_sync(0xffffffffu, best, d);
        best = t < best ? t : best;
    }
    if (lane == 0) atomicMin(&s_min, best);
    __syncthreads();
    if (threadIdx.x == 0) {
        int nk = base_out;
        const bool have_fallback = s_min != ~0ull;                    // some hypothesis survived the mask test
        if (info_out) {                // what a multi-GPU merge needs to apply the never-empty rule globally
            info_out[0] = base_out;                                   // hypotheses that really passed the test
            info_out[1] = have_fallback ? (int)(s_min >> 32) : 0x7fffffff;   // violation count of the fallback candidate
        }
        if (nk == 0 && have_fallback) {        // never empty (among the hypotheses the mask test left)
            keep_idx[0] = (int)(s_min & 0xffffffffull);
            nk = 1;
        }
        *n_keep_out = nk;
    }
}

// CTA shape for the cloud-staging kernels: ONE CTA of 32 warps per SM that shares one staged copy of the cloud.
// Four 8-warp CTAs (four copies, 144 KB at 1,000 points) measured 6 % slower: the frame gathers are what bounds
// these kernels and every KB of shared memory is a KB less L1 for them.  `per_warp` = extra shared memory per warp.
struct cta_shape { int threads, ctas_per_sm; size_t smem; };
#ifndef ZS_MAX_CTAS_PER_SM
#define ZS_MAX_CTAS_PER_SM 1
#endif
cta_shape shape_for(size_t cloud_bytes, size_t per_warp) {
    for (int ctas = ZS_MAX_CTAS_PER_SM; ctas >= 1; --ctas) {
        const int warps = 32 / ctas;                                  // 8, 10->8.., keep multiples that divide 32
        if (32 % ctas) continue;
        const size_t smem = cloud_bytes + per_warp * warps;
        if ((smem + 1024) * ctas <= 220 * 1024) return cta_shape{warps * 32, ctas, smem};
    }
    return cta_shape{1024, 1, cloud_bytes + per_warp * 32};
}

int grid_for(const zs_ctx* ctx, long long n, int ctas_per_sm, int warps_per_cta = kWarpsPerCta) {
    long long need = (n + warps_per_cta - 1) / warps_per_cta;
    int full = ctx->sm_count * ctas_per_sm;
    if (need >= full) return full;
    return need < 1 ? 1 : (int)need;
}

constexpr size_t kCloudSmemMax = 200 * 1024;

template <typename K>
int opt_in_smem(zs_ctx* ctx, K kernel, size_t bytes) {
    if (bytes > 48 * 1024)
        ZS_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    // ask for the smallest shared-memory carve-out that holds one CTA: the rest of the 256 KB stays L1 for the gathers
    const int pct = (int)(((bytes + 2048) * 100 + 228 * 1024 - 1) / (228 * 1024));
    ZS_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct > 100 ? 100 : pct));
    return ZS_OK;
}

int check_obj(zs_ctx* ctx, int slot, const float* poses, obj_view& o, zs_cam& cam) {
    if (!ctx) return ZS_ERR_INVALID;
    if (slot < 0 || slot >= ZS_MAX_OBJECTS || ctx->obj[slot].n_pts == 0)
        return zs_fail(ctx, ZS_ERR_STATE, "object slot %d not set", slot);
    if (!ctx->frame.set) return zs_fail(ctx, ZS_ERR_STATE, "frame not set");
    if (!poses || ((uintptr_t)poses & 15)) return zs_fail(ctx, ZS_ERR_INVALID, "poses must be non-null, 16-byte aligned");
    const zs_object& ob = ctx->obj[slot];
    o = obj_view{ob.pA, ob.pB, ob.pV, ob.n_pts};
    const zs_frame& f = ctx->frame;
    cam = zs_cam{f.fx, f.fy, f.cx, f.cy, f.inv_fx, f.inv_fy, f.H, f.W};
    return ZS_OK;
}

}  // namespace

extern "C" int zs_features(zs_ctx* ctx, int obj_slot, const float* poses, const int32_t* keep_idx, int n_keep,
                           void* feat_out, int feat_dtype, int32_t* uv_out, uint8_t* mask_out,
                           int32_t* viol_out, void* stream) {
    obj_view o;
    zs_cam cam;
    if (ctx && n_keep == 0) return ZS_OK;
    int rc = check_obj(ctx, obj_slot, poses, o, cam);
    if (rc) return rc;
    if (n_keep < 0 || (feat_dtype != ZS_F32 && feat_dtype != ZS_BF16 && feat_dtype != ZS_BF16_SPLIT))
        return zs_fail(ctx, ZS_ERR_INVALID, "n_keep %d, feat_dtype %d", n_keep, feat_dtype);
    if (feat_dtype == ZS_BF16_SPLIT && (uv_out || mask_out || viol_out))
        return zs_fail(ctx, ZS_ERR_UNSUPPORTED, "split-bf16 features have no side outputs: request them with ZS_F32 features");
    if (!feat_out || ((uintptr_t)feat_out & 15) || (uv_out && ((uintptr_t)uv_out & 7)))
        return zs_fail(ctx, ZS_ERR_INVALID, "feat_out must be 16-byte aligned (uv_out 8)");
    ZS_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool in_smem = cloud_smem(o.n_pts) <= kCloudSmemMax;
    const bool aux = uv_out || mask_out || viol_out;
    const cta_shape cs = shape_for(in_smem ? cloud_smem(o.n_pts) : 0, (aux && feat_dtype == ZS_F32) ? 1024 : 0);   // fp32 store staging (side-output kernel): 1 KB/warp
    size_t smem = cs.smem;
    crop_view cv{nullptr, 0, 0, 0, 0};
#ifdef ZS_CROP_STAGE
    if (const char* e = aux ? nullptr : getenv("ZS_CROP_RECT")) {      // the hot (features-only) kernels only
        if (sscanf(e, "%d,%d,%d,%d", &cv.x0, &cv.y0, &cv.w, &cv.h) != 4 || cv.x0 < 0 || cv.y0 < 0 || cv.w <= 0 || cv.h <= 0 ||
            cv.x0 + cv.w > cam.W || cv.y0 + cv.h > cam.H || smem + (size_t)cv.w * cv.h * 16 > 220 * 1024)
            return zs_fail(ctx, ZS_ERR_INVALID, "ZS_CROP_RECT=%s does not fit the frame / %zu bytes of shared memory", e, (size_t)220 * 1024 - smem);
        smem += (size_t)cv.w * cv.h * 16;
    }
#endif
    const int threads = cs.threads;
    const long long units = aux ? n_keep : (long long)n_keep * ((o.n_pts + 255) / 256);
    const int grid = grid_for(ctx, units, cs.ctas_per_sm, threads / 32);
    cudaStream_t st = (cudaStream_t)stream;
    const float4* frame = ctx->frame.packed;
#define ZS_LAUNCH_FEAT(FMT, SM)                                                                         \
    do {                                                                                                \
        if (aux) {                                                                                      \
            rc = opt_in_smem(ctx, zs_k_features<(FMT) == ZS_BF16, SM, true>, smem);                     \
            if (rc) return rc;                                                                          \
            zs_k_features<(FMT) == ZS_BF16, SM, true><<<grid, threads, smem, st>>>(                     \
                o, cam, frame, poses, keep_idx, n_keep, feat_out, uv_out, mask_out, viol_out,           \
                ctx->dyn_n, ctx->dyn_off);                                                              \
        } else {                                                                                        \
            rc = opt_in_smem(ctx, zs_k_features_hot<FMT, SM>, smem);                                    \
            if (rc) return rc;                                                                          \
            zs_k_features_hot<FMT, SM><<<grid, threads, smem, st>>>(o, cam, frame, poses,               \
                                                                   keep_idx, n_keep, feat_out,          \
                                                                   (((uintptr_t)feat_out & 31) == 0),   \
                                                                   ctx->dyn_n, ctx->dyn_off, cv);       \
        }                                                                                               \
    } while (0)
    if (feat_dtype == ZS_BF16) { if (in_smem) ZS_LAUNCH_FEAT(ZS_BF16, true); else ZS_LAUNCH_FEAT(ZS_BF16, false); }
    else if (feat_dtype == ZS_BF16_SPLIT) { if (in_smem) ZS_LAUNCH_FEAT(ZS_BF16_SPLIT, true); else ZS_LAUNCH_FEAT(ZS_BF16_SPLIT, false); }
    else                       { if (in_smem) ZS_LAUNCH_FEAT(ZS_F32, true); else ZS_LAUNCH_FEAT(ZS_F32, false); }
#undef ZS_LAUNCH_FEAT
    ZS_LAUNCHED(ctx);
    return ZS_OK;
}

extern "C" int zs_features_multi(zs_ctx* ctx, int n_seg, const int32_t* obj_slots, const float* const* poses,
                                 const int32_t* const* keep_idx, const int32_t* n_keep, const int32_t* const* n_dev,
                                 const int32_t* n_off, void* const* feat_out, int feat_dtype, void* stream) {
    if (!ctx) return ZS_ERR_INVALID;
    if (n_seg == 0) return ZS_OK;
    if (n_seg < 0 || !obj_slots || !poses || !n_keep || !feat_out ||
        (feat_dtype != ZS_F32 && feat_dtype != ZS_BF16 && feat_dtype != ZS_BF16_SPLIT))
        return zs_fail(ctx, ZS_ERR_INVALID, "zs_features_multi arguments");
    ZS_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    for (int s0 = 0; s0 < n_seg; s0 += kMaxSegsPerLaunch) {
        feat_segs segs;
        zs_cam cam;
        int m = 0, max_pts = 0;
        long long units = 0;
        for (int i = s0; i < n_seg && m < kMaxSegsPerLaunch; ++i) {
            if (n_keep[i] == 0) continue;
            obj_view o;
            int rc = check_obj(ctx, obj_slots[i], poses[i], o, cam);
            if (rc) return rc;
            if (n_keep[i] < 0 || !feat_out[i] || ((uintptr_t)feat_out[i] & 15))
                return zs_fail(ctx, ZS_ERR_INVALID, "segment %d: n_keep %d / feat_out alignment", i, n_keep[i]);
            feat_seg& g = segs.s[m++];
            g.o = o;
            g.poses = poses[i];
            g.keep_idx = keep_idx ? keep_idx[i] : nullptr;
            g.n_dev = n_dev ? n_dev[i] : nullptr;
            g.n_off = (n_dev && n_off) ? n_off[i] : 0;
            g.out = feat_out[i];
            g.n_keep = n_keep[i];
            g.aligned32 = (((uintptr_t)feat_out[i] & 31) == 0);
            g.pad_ = 0;
            max_pts = o.n_pts > max_pts ? o.n_pts : max_pts;
            units += (long long)n_keep[i] * ((o.n_pts + 255) / 256);
        }
        if (m == 0) continue;
        const bool in_smem = cloud_smem(max_pts) <= kCloudSmemMax;
        const cta_shape cs = shape_for(in_smem ? cloud_smem(max_pts) : 0, 0);
        // enough CTAs for every segment to start on its own one; a big launch fills the machine
        int grid = grid_for(ctx, units, cs.ctas_per_sm, cs.threads / 32);
        if (grid < m) grid = m < ctx->sm_count ? m : ctx->sm_count;
        const float4* frame = ctx->frame.packed;
        int rc;
#define ZS_LAUNCH_MULTI(FMT, SM)                                                                        \
    do {                                                                                                \
        rc = opt_in_smem(ctx, zs_k_features_multi<FMT, SM>, cs.smem);                                   \
        if (rc) return rc;                                                                              \
        zs_k_features_multi<FMT, SM><<<grid, cs.threads, cs.smem, st>>>(segs, m, cam, frame);           \
    } while (0)
        if (feat_dtype == ZS_BF16) { if (in_smem) ZS_LAUNCH_MULTI(ZS_BF16, true); else ZS_LAUNCH_MULTI(ZS_BF16, false); }
        else if (feat_dtype == ZS_BF16_SPLIT) { if (in_smem) ZS_LAUNCH_MULTI(ZS_BF16_SPLIT, true); else ZS_LAUNCH_MULTI(ZS_BF16_SPLIT, false); }
        else { if (in_smem) ZS_LAUNCH_MULTI(ZS_F32, true); else ZS_LAUNCH_MULTI(ZS_F32, false); }
#undef ZS_LAUNCH_MULTI
        ZS_LAUNCHED(ctx);
    }
    return ZS_OK;
}

static int mask_min_count(double mask_th, int n_pts) {
    // smallest in-mask count c that the reference keeps: c / N > th in float64 (zephyr_utils.py:68-70)
    int c = (int)floor(mask_th * (double)n_pts) - 1;
    if (c < 0) c = 0;
    while (c <= n_pts && !((double)c / (double)n_pts > mask_th)) ++c;
    return c;
}

extern "C" int zs_violations(zs_ctx* ctx, int obj_slot, const float* poses, int n, const uint8_t* mask, double mask_th,
                             int32_t* viol_out, void* stream) {
    obj_view o;
    zs_cam cam;
    if (ctx && n == 0) return ZS_OK;
    int rc = check_obj(ctx, obj_slot, poses, o, cam);
    if (rc) return rc;
    if (n < 0 || !viol_out) return zs_fail(ctx, ZS_ERR_INVALID, "n %d", n);
    ZS_CUDA(ctx, cudaSetDevice(ctx->device));
    const int mask_min = mask ? mask_min_count(mask_th, o.n_pts) : 0;
    const bool in_smem = cloud_smem(o.n_pts) <= kCloudSmemMax;
    const cta_shape cs = shape_for(in_smem ? cloud_smem(o.n_pts) : 0, 0);
    const int grid = grid_for(ctx, n, cs.ctas_per_sm, cs.threads / 32);
    cudaStream_t st = (cudaStream_t)stream;
#define ZS_LAUNCH_VIOL(SM, MK)                                                                                       \
    do {                                                                                                             \
        if (SM) { rc = opt_in_smem(ctx, zs_k_violations<SM, MK>, cs.smem); if (rc) return rc; }                      \
        zs_k_violations<SM, MK><<<grid, cs.threads, SM ? cs.smem : 0, st>>>(o, cam, ctx->frame.packed, poses, n,     \
                                                                           mask, mask_min, viol_out);                \
    } while (0)
    if (in_smem) { if (mask) ZS_LAUNCH_VIOL(true, true); else ZS_LAUNCH_VIOL(true, false); }
    else         { if (mask) ZS_LAUNCH_VIOL(false, true); else ZS_LAUNCH_VIOL(false, false); }
#undef ZS_LAUNCH_VIOL
    ZS_LAUNCHED(ctx);
    return ZS_OK;
}

extern "C" int zs_prefilter(zs_ctx* ctx, int n_seg, const int32_t* obj_slots, const float* const* poses, const int32_t* n_hyp,
                            const uint8_t* const* masks, double mask_th, float inconst_ratio_th, int32_t* const* viol_out,
                            int32_t* const* keep_idx_out, int32_t* const* n_keep_out, int32_t* const* info_out, void* stream) {
    if (!ctx) return ZS_ERR_INVALID;
    if (n_seg == 0) return ZS_OK;
    if (n_seg < 0 || n_seg > ZS_MAX_OBJECTS || !obj_slots || !poses || !n_hyp || !viol_out || !keep_idx_out || !n_keep_out)
        return zs_fail(ctx, ZS_ERR_INVALID, "zs_prefilter arguments (%d segments, at most %d)", n_seg, ZS_MAX_OBJECTS);
    ZS_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    filt_segs fsegs;
    zs_cam cam;
    for (int s0 = 0; s0 < n_seg; s0 += kMaxSegsPerLaunch) {
        viol_segs vs;
        int m = 0, max_pts = 0;
        long long units = 0;
        for (int i = s0; i < n_seg && i < s0 + kMaxSegsPerLaunch; ++i) {
            if (n_hyp[i] < 0 || !n_keep_out[i] || (n_hyp[i] > 0 && (!viol_out[i] || !keep_idx_out[i])))
                return zs_fail(ctx, ZS_ERR_INVALID, "segment %d: n_hyp %d / null output", i, n_hyp[i]);
            if (n_hyp[i] == 0) continue;
            obj_view o;
            int rc = check_obj(ctx, obj_slots[i], poses[i], o, cam);
            if (rc) return rc;
            viol_seg& g = vs.s[m++];
            g.o = o;
            g.poses = poses[i];
            g.mask = masks ? masks[i] : nullptr;
            g.viol_out = viol_out[i];
            g.n = n_hyp[i];
            g.mask_min = g.mask ? mask_min_count(mask_th, o.n_pts) : 0;
            max_pts = o.n_pts > max_pts ? o.n_pts : max_pts;
            units += n_hyp[i];
        }
        if (m == 0) continue;
        const bool in_smem = cloud_smem(max_pts) <= kCloudSmemMax;
        const cta_shape cs = shape_for(in_smem ? cloud_smem(max_pts) : 0, 0);
        int grid = grid_for(ctx, units, cs.ctas_per_sm, cs.threads / 32);
        if (grid < m) grid = m < ctx->sm_count ? m : ctx->sm_count;
        if (in_smem) {
            int rc = opt_in_smem(ctx, zs_k_violations_multi<true>, cs.smem);
            if (rc) return rc;
            zs_k_violations_multi<true><<<grid, cs.threads, cs.smem, st>>>(vs, m, cam, ctx->frame.packed);
        } else {
            zs_k_violations_multi<false><<<grid, cs.threads, 0, st>>>(vs, m, cam, ctx->frame.packed);
        }
        ZS_LAUNCHED(ctx);
    }
    for (int i = 0; i < n_seg; ++i) {
        const int npts = (obj_slots[i] >= 0 && obj_slots[i] < ZS_MAX_OBJECTS) ? ctx->obj[obj_slots[i]].n_pts : 0;
        fsegs.s[i] = filt_seg{viol_out[i], keep_idx_out[i], n_keep_out[i], info_out ? info_out[i] : nullptr, n_hyp[i],
                              (float)(npts > 0 ? npts : 1)};
    }
    zs_k_filter_multi<<<n_seg, 1024, 0, st>>>(fsegs, inconst_ratio_th);
    ZS_LAUNCHED(ctx);
    return ZS_OK;
}

// DTOID detections -> binary mask (python/ossid/scripts/online_learning.py:389-405): boxes are visited in order; a box
// with score < 0.5 is skipped once the mask already covers a pixel with depth; every other box is grown by
// expandBox (python/ossid/utils/__init__.py:11-16) and filled.  "The mask covers a pixel with depth" only changes when a
// box is filled, so the single CTA carries it as a flag and tests just the pixels of the box it has filled.
namespace {
constexpr int kMaxBoxes = 64;
struct box_list { int n; int x1[kMaxBoxes], y1[kMaxBoxes], x2[kMaxBoxes], y2[kMaxBoxes]; int low_score[kMaxBoxes]; };

__global__ void __launch_bounds__(1024)
zs_k_boxes_to_mask(const __grid_constant__ box_list bl, const float4* __restrict__ frame, int H, int W, uint8_t* __restrict__ mask) {
    __shared__ int s_any;
    if (threadIdx.x == 0) s_any = 0;
    for (int i = threadIdx.x; i < H * W; i += blockDim.x) mask[i] = 0;
    __syncthreads();
    for (int b = 0; b < bl.n; ++b) {
        if (bl.low_score[b] && s_any) continue;                        // block-uniform (s_any only changes behind a barrier)
        const int bw = bl.x2[b] - bl.x1[b], bh = bl.y2[b] - bl.y1[b];
        int any = 0;
        if (bw > 0 && bh > 0) {
            for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) {
                const int px = (bl.y1[b] + i / bw) * W + bl.x1[b] + i % bw;
                mask[px] = 1;
                any |= frame[px].x > 0.f;                              // .x = depth / camera_scale
            }
        }
        __syncthreads();
        if (any) atomicOr(&s_any, 1);
        __syncthreads();
    }
}
}  // namespace

extern "C" int zs_boxes_to_mask(zs_ctx* ctx, const double* boxes, const double* scores, int n_boxes, double expand_ratio,
                                uint8_t* mask_out, void* stream) {
    if (!ctx) return ZS_ERR_INVALID;
    if (!ctx->frame.set) return zs_fail(ctx, ZS_ERR_STATE, "frame not set");
    if (n_boxes < 0 || n_boxes > kMaxBoxes || !mask_out || (n_boxes > 0 && (!boxes || !scores)))
        return zs_fail(ctx, ZS_ERR_INVALID, "zs_boxes_to_mask: %d boxes (at most %d)", n_boxes, kMaxBoxes);
    ZS_CUDA(ctx, cudaSetDevice(ctx->device));
    const int H = ctx->frame.H, W = ctx->frame.W;
    box_list bl;
    bl.n = n_boxes;
    for (int i = 0; i < n_boxes; ++i) {
        // expandBox, float64 as in Python; int() truncates toward zero; numpy slicing clips to the array
        const double x1 = boxes[4 * i], y1 = boxes[4 * i + 1], x2 = boxes[4 * i + 2], y2 = boxes[4 * i + 3];
        const double cx = (x1 + x2) / 2, cy = (y1 + y2) / 2, w = x2 - x1, h = y2 - y1;
        const double ex1 = fmax(0.0, cx - w / 2 * expand_ratio), ex2 = fmin((double)(W - 1), cx + w / 2 * expand_ratio);
        const double ey1 = fmax(0.0, cy - h / 2 * expand_ratio), ey2 = fmin((double)(H - 1), cy + h / 2 * expand_ratio);
        auto clip = [](double v, int hi) { long long t = (long long)v; if (t < 0) t += hi; return (int)(t < 0 ? 0 : (t > hi ? hi : t)); };
        bl.x1[i] = clip(ex1, W); bl.x2[i] = clip(ex2, W); bl.y1[i] = clip(ey1, H); bl.y2[i] = clip(ey2, H);
        bl.low_score[i] = scores[i] < 0.5;
    }
    zs_k_boxes_to_mask<<<1, 1024, 0, (cudaStream_t)stream>>>(bl, ctx->frame.packed, H, W, mask_out);
    ZS_LAUNCHED(ctx);
    return ZS_OK;
}

extern "C" int zs_filter(zs_ctx* ctx, const int32_t* viol, int n, int n_pts, float inconst_ratio_th,
                         int32_t* keep_idx_out, int32_t* n_keep_out, int32_t* info_out, void* stream) {
    if (!ctx) return ZS_ERR_INVALID;
    if (n < 0 || n_pts <= 0 || !n_keep_out || (n > 0 && (!viol || !keep_idx_out)))
        return zs_fail(ctx, ZS_ERR_INVALID, "zs_filter arguments");
    ZS_CUDA(ctx, cudaSetDevice(ctx->device));
    zs_k_filter<<<1, 1024, 0, (cudaStream_t)stream>>>(viol, n, (float)n_pts, inconst_ratio_th, keep_idx_out, n_keep_out,
                                                      info_out);
    ZS_LAUNCHED(ctx);
    return ZS_OK;
}

static int project_common(zs_ctx* ctx, const float* poses, int n, const float* pts, int n_pts,
                          float fx, float fy, float cx, float cy, int H, int W, int32_t* uv_out,
                          const uint8_t* mask, int32_t* count_out, void* stream) {
    if (!ctx) return ZS_ERR_INVALID;
    if (n == 0) return ZS_OK;
    if (n < 0 || n_pts <= 0 || !pts || !poses || ((uintptr_t)poses & 15))
        return zs_fail(ctx, ZS_ERR_INVALID, "projection arguments");
    if ((size_t)n_pts * 12 > kCloudSmemMax) return zs_fail(ctx, ZS_ERR_UNSUPPORTED, "%d model points", n_pts);
    ZS_CUDA(ctx, cudaSetDevice(ctx->device));
    zs_cam cam{fx, fy, cx, cy, 0.f, 0.f, H, W};
    const size_t smem = (size_t)n_pts * 12;
    const int grid = grid_for(ctx, n, 4);
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if (mask) {
        rc = opt_in_smem(ctx, zs_k_project<true>, smem);
        if (rc) return rc;
        zs_k_project<true><<<grid, kThreads, smem, st>>>(poses, n, pts, n_pts, cam, nullptr, mask, count_out);
    } else {
        rc = opt_in_smem(ctx, zs_k_project<false>, smem);
        if (rc) return rc;
        zs_k_project<false><<<grid, kThreads, smem, st>>>(poses, n, pts, n_pts, cam, uv_out, nullptr, nullptr);
    }
    ZS_LAUNCHED(ctx);
    return ZS_OK;
}

extern "C" int zs_project_uv(zs_ctx* ctx, const float* poses, int n, const float* pts, int n_pts,
                             float fx, float fy, float cx, float cy, int32_t* uv_out, void* stream) {
    if (ctx && n > 0 && (!uv_out || ((uintptr_t)uv_out & 7))) return zs_fail(ctx, ZS_ERR_INVALID, "uv_out");
    return project_common(ctx, poses, n, pts, n_pts, fx, fy, cx, cy, 0, 0, uv_out, nullptr, nullptr, stream);
}

extern "C" int zs_mask_count(zs_ctx* ctx, const float* poses, int n, const float* pts, int n_pts,
                             float fx, float fy, float cx, float cy, const uint8_t* mask, int H, int W,
                             int32_t* count_out, void* stream) {
    if (ctx && (!mask || H <= 0 || W <= 0 || (n > 0 && !count_out))) return zs_fail(ctx, ZS_ERR_INVALID, "mask arguments");
    return project_common(ctx, poses, n, pts, n_pts, fx, fy, cx, cy, H, W, nullptr, mask, count_out, stream);
}
