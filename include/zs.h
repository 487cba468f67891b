/*
 * zs.h -- C ABI of the B200-native Zephyr hypothesis-scoring extension (libzs.so).
 *
 * The reference (r-pad/OSSID_code) has no native interface for this path: scoring is
 * reached through Python objects of the un-vendored `zephyr` package.  Each entry
 * point below names the reference call it stands behind (paths relative to the
 * reference checkout).  Bindings: ossid_code_b200/_lib.py (ctypes); the stub a
 * reference maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain C, no C++ exceptions cross the boundary; every call returns ZS_OK (0) or a
 *     negative zs_status; zs_last_error(ctx) gives the message of the last failure.
 *   - pointers marked [dev] are CUDA device pointers on the context's device
 *     (torch: tensor.data_ptr()); [host] are host pointers.  The library never frees or
 *     keeps caller memory beyond a call; frames, model clouds and weights are copied
 *     into context-owned buffers.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  All work
 *     is asynchronous and stream-ordered unless stated.
 *   - one context per device; a context is not thread-safe.
 *   - poses are float32 [n][12] = rows of (R | t), i.e. transforms[:, :3, :4] of the
 *     reference's (n,4,4) float64 hand-over cast once to float32.
 */
#ifndef ZS_H_
#define ZS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define ZS_API __attribute__((visibility("default")))
#else
#define ZS_API
#endif

typedef struct zs_ctx zs_ctx;

typedef enum {
    ZS_OK = 0,
    ZS_ERR_INVALID = -1,      /* bad argument */
    ZS_ERR_CUDA = -2,         /* CUDA runtime/driver error */
    ZS_ERR_STATE = -3,        /* frame / object / weights not set */
    ZS_ERR_UNSUPPORTED = -4,  /* shape outside what the kernels support */
    ZS_ERR_NOMEM = -5
} zs_status;

/* element types: ZS_F32 / ZS_BF16 name feature and precision arguments; ZS_BF16_SPLIT features are two bf16 planes
 * [2][n][n_pts][8] (hi = bf16(x), lo = bf16(x - hi)) feeding the 3-term fp32-accurate tensor-core scorer;
 * ZS_F64 only names the source type of zs_pack_poses. */
enum { ZS_F32 = 0, ZS_BF16 = 1, ZS_BF16_SPLIT = 2, ZS_F64 = 3 };

#define ZS_DIM_POINT 8          /* channels of point_x: u_n v_n dH dS dV dD ncos 0 */
#define ZS_MAX_OBJECTS 64
#define ZS_MAX_WEIGHT_SLOTS 4
#define ZS_MAX_TOPK 64
#define ZS_WEIGHT_FLOATS (64*8+64 + 128*64+128 + 1024*128+1024 + 512*1024+512 + 256*512+256 + 256+1)

/* mask byte per (hypothesis, point) */
#define ZS_BIT_VALID_PROJ  1
#define ZS_BIT_VALID_DEPTH 2
#define ZS_BIT_FRONT       4
#define ZS_BIT_FREE_SPACE  8
#define ZS_BIT_OCCLUDED    16

ZS_API int zs_version(void);
ZS_API const char* zs_strerror(int status);

/* Context life cycle.  Stands behind `ScoreDataset([], "", name, args, mode='test')` and
 * `PointNet2SSG(dim_point, args, num_class=1).to(0).eval()`
 * (python/ossid/scripts/online_learning.py:206-227). */
ZS_API int zs_create(zs_ctx** out, int device);
ZS_API void zs_destroy(zs_ctx* ctx);
ZS_API const char* zs_last_error(const zs_ctx* ctx);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
ZS_API int64_t zs_launch_count(const zs_ctx* ctx);
/* Changes whenever something a recorded launch sequence depends on has changed inside the context: a context-owned
 * buffer moved (bigger frame, bigger cloud, grown scratch) or a launch parameter kept by the context changed (frame
 * size, intrinsics, a slot's point count).  A caller that replays captured launches (CUDA graph) re-captures when it
 * differs from the value at capture time. */
ZS_API int64_t zs_alloc_generation(const zs_ctx* ctx);

/* Sizes the context's scratch for scoring calls of up to max_hypotheses hypotheses, so that no later call allocates
 * or frees device memory between the kernels of a frame (allocation synchronises the device).  Optional: without
 * it the scratch grows on demand. */
ZS_API int zs_reserve(zs_ctx* ctx, int max_hypotheses);

/* Pose hand-over: transforms [dev] (n,4,4) row-major, dtype ZS_F32 or ZS_F64 (the reference hands float64 over,
 * python/ossid/utils/zephyr_utils.py:16) -> poses_out [dev] float32 [n][12] rows of (R | t): the single IEEE
 * round-to-nearest cast of `transforms[:, :3, :4]`. */
ZS_API int zs_pack_poses(zs_ctx* ctx, const void* transforms, int dtype, int n, float* poses_out, void* stream);

/* Frame upload.  Replaces the tensors built at python/ossid/utils/zephyr_utils.py:13-17.
 * zs_set_frame: rgb [dev] float32 H*W*3 in [0,1] (already blurred and divided by 255),
 * depth [dev] float32 H*W.  zs_set_frame_u8: img [dev] uint8 H*W*3 straight from the
 * camera; when blur != 0 the 5x5 Gaussian of cv2.GaussianBlur(img,(5,5),0) is applied on
 * the GPU with cv2's 8-bit fixed-point arithmetic before the /255.
 * Both pack {depth/camera_scale, H, S, V} per pixel into a context-owned frame. */
ZS_API int zs_set_frame(zs_ctx* ctx, const float* rgb, const float* depth, int H, int W,
                 float fx, float fy, float cx, float cy, float camera_scale, void* stream);
ZS_API int zs_set_frame_u8(zs_ctx* ctx, const uint8_t* img, const float* depth, int H, int W,
                    float fx, float fy, float cx, float cy, float camera_scale, int blur, void* stream);

/* Model cloud upload: pts/cols/nrms [dev] float32 (n_pts,3).  Replaces model_points /
 * model_colors / model_normals of the scoring dict (zephyr_utils.py:18-20). */
ZS_API int zs_set_object(zs_ctx* ctx, int slot, const float* pts, const float* cols, const float* nrms,
                  int n_pts, void* stream);

/* Scorer weights: blob [dev] float32, ZS_WEIGHT_FLOATS values = W1 b1 W2 b2 W3 b3 F1 c1 F2 c2
 * F3 c3 (BatchNorm folded), each weight (out,in) row-major.  Replaces
 * `model.load_state_dict(ckpt['state_dict'])` (online_learning.py:213-214). */
ZS_API int zs_set_weights(zs_ctx* ctx, int slot, const float* blob, size_t n_floats, void* stream);

/* `zephyr.utils.projectPointsUv(pose_hypos, model_points, meta_data)`
 * (used at zephyr_utils.py:58): raw rounded pixel indices, no z or bounds test.
 * uv_out [dev] int32 [n][n_pts][2], [..,0]=x/col, [..,1]=y/row. */
ZS_API int zs_project_uv(zs_ctx* ctx, const float* poses, int n, const float* pts, int n_pts,
                  float fx, float fy, float cx, float cy, int32_t* uv_out, void* stream);

/* Fused `filterHypoByMask` (zephyr_utils.py:49-71): per hypothesis, the number of model
 * points that project inside the frame onto a non-zero mask pixel.  mask [dev] uint8 H*W.
 * count_out [dev] int32 [n].  The caller applies `count / n_pts > th`. */
ZS_API int zs_mask_count(zs_ctx* ctx, const float* poses, int n, const float* pts, int n_pts,
                  float fx, float fy, float cx, float cy, const uint8_t* mask, int H, int W,
                  int32_t* count_out, void* stream);

/* First half of `ScoreDataset.getPointNetData` (call site zephyr_utils.py:31): number of
 * free-space-violating points per hypothesis.  viol_out [dev] int32 [n].
 * mask [dev] uint8 H*W (nullable): additionally apply `filterHypoByMask(model_points, meta, poses, mask, mask_th)`
 * (zephyr_utils.py:49-71; mask_th is a float64 as in Python, the comparison `count / n_pts > th` is made in float64)
 * in the same pass: a hypothesis that does not project more than mask_th of its points onto
 * non-zero mask pixels gets viol_out = ZS_VIOL_MASKED (0x7fffffff), which zs_filter never keeps; the kernel abandons
 * such a hypothesis as soon as its remaining points cannot reach the bar (early-out before the remaining gathers). */
#define ZS_VIOL_MASKED 0x7fffffff
ZS_API int zs_violations(zs_ctx* ctx, int obj_slot, const float* poses, int n, const uint8_t* mask, double mask_th,
                  int32_t* viol_out, void* stream);

/* DTOID detections -> binary mask (python/ossid/scripts/online_learning.py:389-405) against the resident frame's depth:
 * boxes [host] float64 [n_boxes][4] = x1 y1 x2 y2, scores [host] float64 [n_boxes], visited in order; a box with score
 * < 0.5 is skipped once the mask covers a pixel with depth > 0; the others are grown by expandBox(.., expand_ratio)
 * (python/ossid/utils/__init__.py:11-16) and filled.  mask_out [dev] uint8 H*W in {0, 1}.  At most 64 boxes. */
ZS_API int zs_boxes_to_mask(zs_ctx* ctx, const double* boxes, const double* scores, int n_boxes, double expand_ratio,
                     uint8_t* mask_out, void* stream);

/* Hypothesis pre-filter of getPointNetData (its effect is visible at zephyr_utils.py:39-43;
 * thresholds online_learning.py:174,184): keep h iff viol[h]*100/n_pts < th (th >= 100
 * keeps all); never empty (first minimum kept; entries equal to ZS_VIOL_MASKED are never kept, so the result is
 * empty when the mask test dropped everything).  keep_idx_out [dev] int32 [n] ascending,
 * n_keep_out [dev] int32[1].  info_out [dev] int32[2] (nullable) = {hypotheses that really passed the test (0 when
 * the never-empty rule supplied the single kept one), violation count of that fallback}: what zs_merge_topk needs to
 * apply the never-empty rule to the object's WHOLE list when the list is sharded over GPUs. */
ZS_API int zs_filter(zs_ctx* ctx, const int32_t* viol, int n, int n_pts, float inconst_ratio_th,
              int32_t* keep_idx_out, int32_t* n_keep_out, int32_t* info_out, void* stream);

/* zs_violations + zs_filter for ALL objects of a frame in two launches (one projection / depth-gather pass over every
 * segment, then one CTA per object for the compaction): segment i = n_hyp[i] hypotheses poses[i] of the cloud in
 * obj_slots[i], optional masks[i] (nullable array / entries) tested with mask_th as in zs_violations, kept by
 * inconst_ratio_th as in zs_filter.  viol_out[i] [dev] int32 [n_hyp[i]], keep_idx_out[i] [dev] int32 [n_hyp[i]],
 * n_keep_out[i] [dev] int32[1], info_out[i] [dev] int32[2] (nullable).  All array arguments are [host] arrays of n_seg
 * (<= ZS_MAX_OBJECTS) entries. */
ZS_API int zs_prefilter(zs_ctx* ctx, int n_seg, const int32_t* obj_slots, const float* const* poses, const int32_t* n_hyp,
                 const uint8_t* const* masks, double mask_th, float inconst_ratio_th, int32_t* const* viol_out,
                 int32_t* const* keep_idx_out, int32_t* const* n_keep_out, int32_t* const* info_out, void* stream);

/* Device-side hypothesis count for the calls that follow (no host read-back of zs_filter's count, so a filtered frame
 * stays asynchronous): while n_dev [dev] int32[1] is set, zs_features (n_keep) and zs_pool with bf16 features (n)
 * treat their count argument as a CAPACITY and process entries [n_offset, n_offset + capacity) of a list that has
 * *n_dev entries, i.e. min(capacity, max(0, *n_dev - n_offset)) hypotheses; rows beyond that are left untouched.
 * zs_score and the fp32 zs_pool return ZS_ERR_UNSUPPORTED while it is set; zs_head and zs_topk_segments are
 * unaffected (run zs_head over the capacity and pass the device count through the segment table).
 * n_dev = NULL restores host counts. */
ZS_API int zs_set_dynamic_count(zs_ctx* ctx, const int32_t* n_dev, int n_offset);

/* Second half of getPointNetData: features for hypotheses keep_idx[0..n_keep) of `poses`
 * (keep_idx NULL = all of 0..n_keep).  feat_out [dev] [n_keep][n_pts][8] float32 or bf16, or for ZS_BF16_SPLIT
 * [n_keep][2][n_pts][8] bf16 = per hypothesis a plane of bf16(x) and a plane of bf16(x - bf16(x)) (no side outputs
 * with that format);
 * uv_out [dev] int32 [n_keep][n_pts][2] (nullable); mask_out [dev] uint8 [n_keep][n_pts]
 * (nullable); viol_out [dev] int32 [n_keep] (nullable). */
ZS_API int zs_features(zs_ctx* ctx, int obj_slot, const float* poses, const int32_t* keep_idx, int n_keep,
                void* feat_out, int feat_dtype, int32_t* uv_out, uint8_t* mask_out,
                int32_t* viol_out, void* stream);

/* zs_features (no side outputs) for several objects of the frame in ONE launch: segment i featurises hypotheses
 * keep_idx[i][0..n_keep[i]) (keep_idx NULL or keep_idx[i] NULL = 0..n_keep[i]) of poses[i] against the cloud in
 * obj_slots[i] and writes feat_out[i].  All array arguments are [host] arrays of n_seg entries holding [dev] pointers;
 * n_dev / n_off (nullable) are the per-segment device-side counts of zs_set_dynamic_count (which this call ignores). */
ZS_API int zs_features_multi(zs_ctx* ctx, int n_seg, const int32_t* obj_slots, const float* const* poses,
                      const int32_t* const* keep_idx, const int32_t* n_keep, const int32_t* const* n_dev,
                      const int32_t* n_off, void* const* feat_out, int feat_dtype, void* stream);

/* Scorer forward, `model({"point_x": point_x})` (zephyr_utils.py:34).  feat [dev] [n][n_pts][8] in feat_dtype
 * ([n][2][n_pts][8] bf16 for ZS_BF16_SPLIT).  precision ZS_BF16 (feat ZS_BF16): bf16 tcgen05 path, 1e-2;
 * precision ZS_F32: fp32-accurate, 1e-4 -- feat ZS_BF16_SPLIT: tcgen05 with 3-term bf16-split products and the fp32
 * head; feat ZS_F32: the CUDA-core fp32 kernel (reference implementation of the same arithmetic).
 * scores_out [dev] float32 [n]. */
ZS_API int zs_score(zs_ctx* ctx, int weight_slot, const void* feat, int feat_dtype, int n, int n_pts,
             int precision, float* scores_out, void* stream);

/* point_x handed over as float32 (the reference's dtype at zephyr_utils.py:34) -> the split-bf16 planes the fp32-accurate
 * tensor-core scorer reads: feat [dev] float32 [n][n_pts][8] -> split_out [dev] bf16 [n][2][n_pts][8]. */
ZS_API int zs_split_features(zs_ctx* ctx, const float* feat, int n, int n_pts, void* split_out, void* stream);

/* The two halves of zs_score, exposed so that callers can keep the pooled vectors and so that
 * each stage can be timed alone: zs_pool = shared per-point MLP + max over points
 * (pooled_out [dev] float32 [n][1024]; feat_dtype ZS_F32 / ZS_BF16 / ZS_BF16_SPLIT picks the kernel as in zs_score);
 * zs_head = 1024 -> 512 -> 256 -> 1, precision ZS_BF16_SPLIT: fp32-accurate on the tensor cores (3-term tf32 products:
 * the head of the fp32-accurate path; falls back to the CUDA-core kernels below 1,024 rows), ZS_F32: fp32 on
 * CUDA cores (1e-4 parity path), ZS_BF16: tf32 tensor cores (what zs_score uses after the bf16 MLP). */
ZS_API int zs_pool(zs_ctx* ctx, int weight_slot, const void* feat, int feat_dtype, int n, int n_pts,
            float* pooled_out, void* stream);
ZS_API int zs_head(zs_ctx* ctx, int weight_slot, const float* pooled, int n, int precision, float* scores_out,
            void* stream);

/* zs_features + zs_pool (bf16 tensor-core path) in ONE kernel: the producer warps of the MLP kernel project, gather and
 * featurise each 128-point tile themselves and hand it to the tensor cores through shared memory, so the feature rows
 * never travel through HBM.  The hypothesis list is the concatenation of n_seg segments: segment i = n_hyp[i]
 * hypotheses of the cloud in obj_slots[i]; hypothesis j of the segment is pose keep_idx[i][j] of poses[i] [dev] float32
 * [..][12] (keep_idx NULL or keep_idx[i] NULL: pose j).  n_dev (nullable, entries nullable): [dev] int32[1] per segment,
 * the device-side count of zs_filter; then n_hyp[i] is the segment's CAPACITY and min(n_hyp[i], *n_dev[i]) hypotheses
 * are scored.  obj_slots / poses / keep_idx / n_hyp / n_dev are [host] arrays of n_seg entries.  All clouds of a call
 * have the same number of points (>= 128).  pooled_out [dev] float32 [sum n_hyp][1024]: segment i's rows start at
 * sum(n_hyp[0..i)) (rows beyond a device-side count are left untouched); bit-identical to zs_features(ZS_BF16)
 * followed by zs_pool. */
ZS_API int zs_pool_fused(zs_ctx* ctx, int weight_slot, int n_seg, const int32_t* obj_slots, const float* const* poses,
                  const int32_t* const* keep_idx, const int32_t* n_hyp, const int32_t* const* n_dev,
                  float* pooled_out, void* stream);

/* Diagnostic twin of zs_pool for ZS_BF16 / ZS_BF16_SPLIT features: additionally dumps the activations of layers 1
 * and 2 as the next layer reads them (h1_out [dev] float32 [n*n_pts][64], h2_out [dev] float32 [n*n_pts][128];
 * either may be NULL).  Used by the parity tests to localise a tensor-core mismatch. */
ZS_API int zs_pool_debug(zs_ctx* ctx, int weight_slot, const void* feat, int feat_dtype, int n, int n_pts, float* pooled_out,
                  float* h1_out, float* h2_out, void* stream);

/* Per-object top-k, ordered by (score desc, index asc); k <= ZS_MAX_TOPK.  Generalises
 * `scores.argmax()` (online_learning.py:466-467; first maximum wins ties).
 * s_out [dev] float32 [k], i_out [dev] int32 [k] = index_map[i] + index_base (index_map [dev] int32 [n],
 * NULL = identity; pass zs_filter's keep_idx to get indices into the unfiltered hypothesis list; an entry whose
 * index_map value is negative is an empty slot and is skipped); a NaN score never beats a real one and is reported
 * as -inf; output slots beyond the candidates are (-inf, -1). */
ZS_API int zs_topk(zs_ctx* ctx, const float* scores, int n, int k, int index_base, const int32_t* index_map,
            float* s_out, int32_t* i_out, void* stream);

/* The same for all objects of a frame in one launch (one CTA per object).  segments [dev] int32 [n_segments][4] =
 * {first score, count, index_base, map_delta}; scores [dev] are indexed by first score + i, index_map [dev]
 * (nullable) by first score + map_delta + i (map_delta = 0: the two arrays share one layout).
 * s_out [dev] float32 [n_segments][k], i_out [dev] int32 [n_segments][k]. */
ZS_API int zs_topk_segments(zs_ctx* ctx, const float* scores, const int32_t* segments, int n_segments, int k,
                     const int32_t* index_map, float* s_out, int32_t* i_out, void* stream);

/* Multi-GPU merge (the reduction the reference does with one argmax, online_learning.py:466-467, after the hypothesis
 * list has been sharded): gathered [dev] int32 [world][rec_ints] = the all-gathered per-rank records, each
 *   [0, n_obj*k) score bits | [n_obj*k, 2*n_obj*k) global hypothesis indices (-1 = empty) |
 *   [2*n_obj*k, 2*n_obj*k + 2*n_obj) zs_filter's info_out per object ({1, 0} when no pre-filter ran) |
 *   (only when poses_out != NULL; starts at the next multiple of 4 ints) n_obj*k*12 floats: the candidates'
 *   poses (zs_gather_poses),
 * rank r holding hypotheses of a lower index range than rank r+1.  Output per object: the k best of all ranks by
 * (score desc, index asc), with the never-empty rule of zs_filter applied to the whole list (fallback candidates are
 * dropped when any rank kept a hypothesis; otherwise only the first minimum-violation one survives).  Identical on
 * every rank and identical to the unsharded result.  s_out [dev] float32 [n_obj][k], i_out [dev] int32 [n_obj][k],
 * poses_out [dev] float32 [n_obj][k][12] (nullable): the winners' poses, for the fp32 re-rank. */
ZS_API int zs_merge_topk(zs_ctx* ctx, const int32_t* gathered, int world, int rec_ints, int n_obj, int k,
                  float* s_out, int32_t* i_out, float* poses_out, void* stream);

/* Poses of the top-k candidates: out [dev] float32 [n_seg][k][12] = poses[segments[g].first_row + idx[g][j] -
 * segments[g].index_base] for idx [dev] int32 [n_seg][k] >= 0 (zeros otherwise); segments [dev] int32 [n_seg][4] =
 * {first pose row of the object, index_base, 0, 0}.  The k winners' poses travel with the all-gathered record so that
 * any rank can re-score any candidate (the fp32-accurate re-rank that makes top-1 independent of the bf16 rounding). */
ZS_API int zs_gather_poses(zs_ctx* ctx, const float* poses, const int32_t* idx, const int32_t* segments, int n_seg, int k,
                    float* out, void* stream);

/* Batched ADD / ADI pose error of every hypothesis against one ground-truth pose; replaces the Python loop
 * `[err_func(R, t, R_gt, t_gt, model_points) for mat in poses_all]` (online_learning.py:452, err_func = add | adi
 * from zephyr.utils.metrics, :32,337-339).  gt_pose [dev] float32 [12] (R|t rows), pts [dev] float32 (n_pts,3),
 * symmetric != 0 selects ADI (nearest neighbour).  err_out [dev] float32 [n], metres. */
ZS_API int zs_pose_errors(zs_ctx* ctx, const float* poses, int n, const float* gt_pose, const float* pts, int n_pts,
                   int symmetric, float* err_out, void* stream);

/* Batched point-to-point ICP of n poses against the depth image; replaces
 * `icpRefinement(depth, uv_original[pred_idx], pred_pose, cam_K, model_points, inpaint_depth=False, icp_max_dist=0.01)`
 * (online_learning.py:476-479; zephyr.utils.icp, which runs Open3D's registration_icp on the CPU for the single
 * winning pose).  Target cloud = depth back-projected at the pixels uv [dev] int32 ([n][n_src][2] when uv_per_pose != 0,
 * else one [n_src][2] shared by all poses; [..,0]=x/col, [..,1]=y/row; pixels outside the image or without depth are
 * ignored); source = src_pts [dev] float32 (n_src,3) under each pose.  depth [dev] float32 H*W metres with intrinsics
 * fx..cy, or NULL = the context's resident frame (zs_set_frame*; H, W and intrinsics arguments are then ignored).
 * Open3D's loop and default stopping rule (relative fitness / rmse 1e-6).  poses_out [dev] float32 [n][12];
 * stats_out [dev] float32 [n][4] = {fitness, inlier_rmse, iterations, correspondences} (nullable). */
ZS_API int zs_icp_refine(zs_ctx* ctx, const float* poses, int n, const float* src_pts, int n_src, const int32_t* uv,
                  int uv_per_pose, const float* depth, int H, int W, float fx, float fy, float cx, float cy,
                  float max_dist, int max_iter, float* poses_out, float* stats_out, void* stream);

/* `estimate_visib_mask_gt(depth, pred_depth, 15/1000.)` (online_learning.py:500; bop_toolkit_lib.visibility):
 * bop19 rule  visible = (d_model - d_test <= delta || d_test == 0) && d_model > 0;  bop18 != 0 selects
 * (d_model - d_test <= delta) && d_test > 0 && d_model > 0.  d_test, d_model [dev] float32 [n]; mask_out [dev] uint8 [n]. */
ZS_API int zs_visib_mask(zs_ctx* ctx, const float* d_test, const float* d_model, size_t n, float delta, int bop18,
                  uint8_t* mask_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ZS_H_ */
