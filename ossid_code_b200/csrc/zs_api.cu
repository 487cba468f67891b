// Context life cycle, frame / model-cloud / weight upload.  ABI: include/zs.h.
#include <stdarg.h>
#include <new>

#include "zs_common.cuh"

int zs_fail(zs_ctx* ctx, int code, const char* fmt, ...) {
    if (ctx) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(ctx->err, sizeof(ctx->err), fmt, ap);
        va_end(ap);
    }
    return code;
}

int zs_reserve_ws(zs_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->ws_bytes) return ZS_OK;
    // Grow-only scratch; cudaFree synchronises the device, which orders it after pending users.  Steady-state callers
    // size it once up front with zs_reserve() so that no allocation happens between the kernels of a frame.
    ctx->alloc_gen++;
    if (ctx->ws) ZS_CUDA(ctx, cudaFree(ctx->ws));
    ctx->ws = nullptr;
    ctx->ws_bytes = 0;
    size_t want = bytes + (bytes >> 2);
    if (cudaMalloc(&ctx->ws, want) != cudaSuccess) {
        cudaGetLastError();
        return zs_fail(ctx, ZS_ERR_NOMEM, "workspace of %zu bytes", want);
    }
    ctx->ws_bytes = want;
    return ZS_OK;
}

extern "C" int zs_version(void) { return 200; }

extern "C" int zs_reserve(zs_ctx* ctx, int max_hypotheses) {
    if (!ctx) return ZS_ERR_INVALID;
    if (max_hypotheses < 0) return zs_fail(ctx, ZS_ERR_INVALID, "max_hypotheses %d", max_hypotheses);
    ZS_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t chunk = max_hypotheses < ZS_SCORE_CHUNK ? max_hypotheses : ZS_SCORE_CHUNK;
    return zs_reserve_ws(ctx, chunk * ZS_HEAD_WS_FLOATS * sizeof(float));
}

// transforms (n,4,4) float32 or float64, row-major -> poses [n][12] float32 rows of (R | t): the one cast of the
// hand-over (python/ossid/utils/zephyr_utils.py:16, float64) to the kernels' float32, IEEE round-to-nearest.
template <typename T>
__global__ void zs_k_pack_poses(const T* __restrict__ src, int n, float* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n * 12) dst[i] = (float)src[(size_t)(i / 12) * 16 + (i % 12)];     // the first 12 of 16 are rows 0-2
}

extern "C" int zs_pack_poses(zs_ctx* ctx, const void* transforms, int dtype, int n, float* poses_out, void* stream) {
    if (!ctx) return ZS_ERR_INVALID;
    if (n == 0) return ZS_OK;
    if (n < 0 || n > (1 << 27) || !transforms || !poses_out || (dtype != ZS_F32 && dtype != ZS_F64))
        return zs_fail(ctx, ZS_ERR_INVALID, "zs_pack_poses n %d dtype %d", n, dtype);
    ZS_CUDA(ctx, cudaSetDevice(ctx->device));
    const int grid = (n * 12 + 255) / 256;
    if (dtype == ZS_F32) zs_k_pack_poses<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)transforms, n, poses_out);
    else zs_k_pack_poses<double><<<grid, 256, 0, (cudaStream_t)stream>>>((const double*)transforms, n, poses_out);
    ZS_LAUNCHED(ctx);
    return ZS_OK;
}

extern "C" const char* zs_strerror(int status) {
    switch (status) {
        case ZS_OK: return "ok";
        case ZS_ERR_INVALID: return "invalid argument";
        case ZS_ERR_CUDA: return "CUDA error";
        case ZS_ERR_STATE: return "frame, object or weights not set";
        case ZS_ERR_UNSUPPORTED: return "unsupported shape";
        case ZS_ERR_NOMEM: return "out of device memory";
        default: return "unknown status";
    }
}

extern "C" int zs_create(zs_ctx** out, int device) {
    if (!out) return ZS_ERR_INVALID;
    *out = nullptr;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev) {
        cudaGetLastError();
        return ZS_ERR_CUDA;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return ZS_ERR_CUDA;
    if (prop.major != 10) return ZS_ERR_UNSUPPORTED;    // sm_100a only; no other code path exists
    if (cudaSetDevice(device) != cudaSuccess) return ZS_ERR_CUDA;
    zs_ctx* ctx = new (std::nothrow) zs_ctx();
    if (!ctx) return ZS_ERR_NOMEM;
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    float lut[256];
    for (int i = 0; i < 256; ++i) lut[i] = (float)((double)i / 255.0);   // zephyr_utils.py:14, then one f32 cast
    if (cudaMalloc(&ctx->lut255, sizeof(lut)) != cudaSuccess ||
        cudaMemcpy(ctx->lut255, lut, sizeof(lut), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(ctx->lut255);
        delete ctx;
        return ZS_ERR_NOMEM;
    }
    *out = ctx;
    return ZS_OK;
}

extern "C" void zs_destroy(zs_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    zs_tc_destroy(ctx);
    cudaFree(ctx->frame.packed);
    for (auto& o : ctx->obj) { cudaFree(o.pA); cudaFree(o.pB); cudaFree(o.pV); }
    for (auto& w : ctx->w) { cudaFree(w.f32); cudaFree(w.f32t); cudaFree(w.bf16); cudaFree(w.bf16x2); cudaFree(w.head_lo); }
    cudaFree(ctx->ws);
    cudaFree(ctx->lut255);
    delete ctx;
}

extern "C" const char* zs_last_error(const zs_ctx* ctx) { return ctx ? ctx->err : "null context"; }
extern "C" int64_t zs_launch_count(const zs_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" int64_t zs_alloc_generation(const zs_ctx* ctx) { return ctx ? ctx->alloc_gen : 0; }

// ---------------------------------------------------------------------------------------
// Frame packing: {depth/camera_scale, H, S, V} per pixel (one 16-byte gather per projected point).
// ---------------------------------------------------------------------------------------
__global__ void zs_k_pack_frame(const float* __restrict__ rgb, const float* __restrict__ depth,
                                float4* __restrict__ out, int n_px, float cam_scale) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_px; i += gridDim.x * blockDim.x) {
        float h, s, v;
        zs_rgb_to_hsv(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2], h, s, v);
        out[i] = make_float4(xdiv(depth[i], cam_scale), h, s, v);
    }
}

// 5x5 Gaussian of cv2.GaussianBlur(img,(5,5),0) on uint8: sigma = 0.3*((5-1)*0.5-1)+0.8 = 1.1,
// OpenCV picks its fixed small-kernel table {1,4,6,4,1}/16 for ksize 5 with sigma <= 0, applies
// it separably with BORDER_REFLECT_101 and rounds once at the end in 8.8 fixed point.  The /255
// of zephyr_utils.py:14 is an fp64 division rounded once to fp32; `lut` holds those 256 values.
__global__ void zs_k_blur_pack(const uint8_t* __restrict__ img, const float* __restrict__ depth,
                               float4* __restrict__ out, int H, int W, float cam_scale, int blur,
                               const float* __restrict__ lut) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const int kw[5] = {1, 4, 6, 4, 1};
    int c[3];
    if (blur) {
        int acc[3] = {0, 0, 0};
        for (int dy = -2; dy <= 2; ++dy) {
            int yy = y + dy;
            yy = yy < 0 ? -yy : (yy >= H ? 2 * H - 2 - yy : yy);
            for (int dx = -2; dx <= 2; ++dx) {
                int xx = x + dx;
                xx = xx < 0 ? -xx : (xx >= W ? 2 * W - 2 - xx : xx);
                int wgt = kw[dy + 2] * kw[dx + 2];
                const uint8_t* p = img + ((size_t)yy * W + xx) * 3;
                acc[0] += wgt * p[0]; acc[1] += wgt * p[1]; acc[2] += wgt * p[2];
            }
        }
        for (int k = 0; k < 3; ++k) c[k] = (acc[k] + 128) >> 8;
    } else {
        const uint8_t* p = img + ((size_t)y * W + x) * 3;
        c[0] = p[0]; c[1] = p[1]; c[2] = p[2];
    }
    float h, s, v;
    zs_rgb_to_hsv(lut[c[0]], lut[c[1]], lut[c[2]], h, s, v);
    size_t i = (size_t)y * W + x;
    out[i] = make_float4(xdiv(depth[i], cam_scale), h, s, v);
}

static int zs_frame_common(zs_ctx* ctx, int H, int W, float fx, float fy, float cx, float cy, float cam_scale) {
    if (!ctx) return ZS_ERR_INVALID;
    if (H <= 0 || W <= 0 || (int64_t)H * W > (int64_t)1 << 28)
        return zs_fail(ctx, ZS_ERR_INVALID, "frame %dx%d", H, W);
    if (!(cam_scale > 0.f)) return zs_fail(ctx, ZS_ERR_INVALID, "camera_scale %f", cam_scale);
    ZS_CUDA(ctx, cudaSetDevice(ctx->device));
    size_t n_px = (size_t)H * W;
    if (n_px > ctx->frame.cap_px) {
        ctx->frame.set = false;             // stays unset if the regrow fails (zs_features then reports ZS_ERR_STATE)
        ctx->alloc_gen++;
        ZS_CUDA(ctx, cudaFree(ctx->frame.packed));
        ctx->frame.packed = nullptr;
        ctx->frame.cap_px = 0;
        if (cudaMalloc(&ctx->frame.packed, n_px * sizeof(float4)) != cudaSuccess) {
            cudaGetLastError();
            return zs_fail(ctx, ZS_ERR_NOMEM, "frame %dx%d", H, W);
        }
        ctx->frame.cap_px = n_px;
    }
    if (ctx->frame.H != H || ctx->frame.W != W || ctx->frame.fx != fx || ctx->frame.fy != fy || ctx->frame.cx != cx ||
        ctx->frame.cy != cy)
        ctx->alloc_gen++;                        // captured launches carry the camera as a parameter
    ctx->frame.H = H; ctx->frame.W = W;
    ctx->frame.fx = fx; ctx->frame.fy = fy; ctx->frame.cx = cx; ctx->frame.cy = cy;
    ctx->frame.inv_fx = 1.0f / fx;      // fp32 reciprocal, as the oracle computes it
    ctx->frame.inv_fy = 1.0f / fy;
    return ZS_OK;
}

extern "C" int zs_set_frame(zs_ctx* ctx, const float* rgb, const float* depth, int H, int W,
                            float fx, float fy, float cx, float cy, float camera_scale, void* stream) {
    int rc = zs_frame_common(ctx, H, W, fx, fy, cx, cy, camera_scale);
    if (rc) return rc;
    if (!rgb || !depth) return zs_fail(ctx, ZS_ERR_INVALID, "null frame pointer");
    int n_px = H * W;
    int grid = min((n_px + 255) / 256, ctx->sm_count * 8);
    zs_k_pack_frame<<<grid, 256, 0, (cudaStream_t)stream>>>(rgb, depth, ctx->frame.packed, n_px, camera_scale);
    ZS_LAUNCHED(ctx);
    ctx->frame.set = true;
    return ZS_OK;
}

extern "C" int zs_set_frame_u8(zs_ctx* ctx, const uint8_t* img, const float* depth, int H, int W,
                               float fx, float fy, float cx, float cy, float camera_scale, int blur,
                               void* stream) {
    int rc = zs_frame_common(ctx, H, W, fx, fy, cx, cy, camera_scale);
    if (rc) return rc;
    if (!img || !depth) return zs_fail(ctx, ZS_ERR_INVALID, "null frame pointer");
    if (blur && (H < 3 || W < 3)) return zs_fail(ctx, ZS_ERR_UNSUPPORTED, "blur needs H,W >= 3");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 blk(32, 8), grd((W + 31) / 32, (H + 7) / 8);
    zs_k_blur_pack<<<grd, blk, 0, st>>>(img, depth, ctx->frame.packed, H, W, camera_scale, blur,
                                        ctx->lut255);
    ZS_LAUNCHED(ctx);
    ctx->frame.set = true;
    return ZS_OK;
}

// ---------------------------------------------------------------------------------------
// Model cloud packing: {px,py,pz,Hm} {nx,ny,nz,Sm} Vm.
// ---------------------------------------------------------------------------------------
__global__ void zs_k_pack_object(const float* __restrict__ pts, const float* __restrict__ cols,
                                 const float* __restrict__ nrms, float4* __restrict__ pA,
                                 float4* __restrict__ pB, float* __restrict__ pV, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float h, s, v;
    zs_rgb_to_hsv(cols[3 * i], cols[3 * i + 1], cols[3 * i + 2], h, s, v);
    pA[i] = make_float4(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], h);
    pB[i] = make_float4(nrms[3 * i], nrms[3 * i + 1], nrms[3 * i + 2], s);
    pV[i] = v;
}

extern "C" int zs_set_object(zs_ctx* ctx, int slot, const float* pts, const float* cols,
                             const float* nrms, int n_pts, void* stream) {
    if (!ctx) return ZS_ERR_INVALID;
    if (slot < 0 || slot >= ZS_MAX_OBJECTS) return zs_fail(ctx, ZS_ERR_INVALID, "object slot %d", slot);
    if (n_pts <= 0 || n_pts > (1 << 20) || !pts || !cols || !nrms)
        return zs_fail(ctx, ZS_ERR_INVALID, "object with %d points", n_pts);
    ZS_CUDA(ctx, cudaSetDevice(ctx->device));
    zs_object& o = ctx->obj[slot];
    if (n_pts > o.cap) {
        ctx->alloc_gen++;
        ZS_CUDA(ctx, cudaFree(o.pA)); ZS_CUDA(ctx, cudaFree(o.pB)); ZS_CUDA(ctx, cudaFree(o.pV));
        o.pA = o.pB = nullptr; o.pV = nullptr; o.cap = 0; o.n_pts = 0;
        if (cudaMalloc(&o.pA, n_pts * sizeof(float4)) != cudaSuccess ||
            cudaMalloc(&o.pB, n_pts * sizeof(float4)) != cudaSuccess ||
            cudaMalloc(&o.pV, n_pts * sizeof(float)) != cudaSuccess) {
            cudaGetLastError();
            return zs_fail(ctx, ZS_ERR_NOMEM, "object with %d points", n_pts);
        }
        o.cap = n_pts;
    }
    zs_k_pack_object<<<(n_pts + 255) / 256, 256, 0, (cudaStream_t)stream>>>(pts, cols, nrms, o.pA, o.pB, o.pV, n_pts);
    ZS_LAUNCHED(ctx);
    if (o.n_pts != n_pts) ctx->alloc_gen++;      // captured launches carry the point count as a parameter
    o.n_pts = n_pts;
    return ZS_OK;
}

extern "C" int zs_set_dynamic_count(zs_ctx* ctx, const int32_t* n_dev, int n_offset) {
    if (!ctx) return ZS_ERR_INVALID;
    if (n_dev && n_offset < 0) return zs_fail(ctx, ZS_ERR_INVALID, "n_offset %d", n_offset);
    ctx->dyn_n = n_dev;
    ctx->dyn_off = n_dev ? n_offset : 0;
    return ZS_OK;
}

extern "C" int zs_set_weights(zs_ctx* ctx, int slot, const float* blob, size_t n_floats, void* stream) {
    if (!ctx) return ZS_ERR_INVALID;
    if (slot < 0 || slot >= ZS_MAX_WEIGHT_SLOTS) return zs_fail(ctx, ZS_ERR_INVALID, "weight slot %d", slot);
    if (!blob || n_floats != (size_t)ZS_WEIGHT_FLOATS)
        return zs_fail(ctx, ZS_ERR_INVALID, "weight blob has %zu floats, expected %d", n_floats, ZS_WEIGHT_FLOATS);
    ZS_CUDA(ctx, cudaSetDevice(ctx->device));
    zs_weights& w = ctx->w[slot];
    if (!w.f32 && cudaMalloc(&w.f32, ZS_WEIGHT_FLOATS * sizeof(float)) != cudaSuccess) {
        cudaGetLastError();
        return zs_fail(ctx, ZS_ERR_NOMEM, "weights");
    }
    cudaStream_t st = (cudaStream_t)stream;
    ZS_CUDA(ctx, cudaMemcpyAsync(w.f32, blob, ZS_WEIGHT_FLOATS * sizeof(float), cudaMemcpyDeviceToDevice, st));
    int rc = zs_f32_prepare_weights(ctx, slot, st);
    if (rc) return rc;
    rc = zs_tc_prepare_weights(ctx, slot, st);
    if (rc) return rc;
    rc = zs_tc3_prepare_weights(ctx, slot, st);
    if (rc) return rc;
    rc = zs_head_prepare_weights(ctx, slot, st);
    if (rc) return rc;
    w.set = true;
    return ZS_OK;
}
