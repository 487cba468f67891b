"""CPU restatement of the post-scoring refinement step.  TEST INFRASTRUCTURE ONLY: nothing under
ossid_code_b200/ imports this module; tests/ and the smoke check use it as the checker.

Reference call sites (python/ossid/scripts/online_learning.py):
  :474-479   pred_pose, _ = icpRefinement(depth, uv_original[pred_idx], pred_pose, cam_K, model_points,
                                          inpaint_depth=False, icp_max_dist=0.01)
  :497       pred_mask_visib = estimate_visib_mask_gt(depth, pred_depth, 15/1000.)

PARITY UNPINNED.  `icpRefinement` lives in the un-vendored `zephyr` package and runs Open3D's point-to-point ICP;
`estimate_visib_mask_gt` lives in the un-vendored `bop_toolkit_lib.visibility`.  Neither is on disk, neither is
version-pinned (readme.md:36-50).  What is restated here is the PUBLISHED algorithm of each:

* Open3D `registration_icp` with `TransformationEstimationPointToPoint` and the default `ICPConvergenceCriteria`
  (relative_fitness 1e-6, relative_rmse 1e-6, max_iteration 30): evaluate nearest-neighbour correspondences within
  `max_correspondence_distance`; loop {closed-form rigid update from the correspondences (Umeyama / Kabsch, proper
  rotation), apply it to the source, re-evaluate, stop when |d fitness| and |d inlier_rmse| both fall below the
  thresholds}.  fitness = #correspondences / #source points, inlier_rmse = sqrt(sum d^2 / #correspondences).
* BOP toolkit `visibility._estimate_visib_mask`, mode 'bop19':
  visible = (d_model - d_test <= delta  OR  d_test == 0) AND d_model > 0.

Frozen here (no authority on disk): the target cloud is the depth image back-projected at the pixels `uv` (x = col, y =
row, as everywhere in this repo) with depth > 0; the source cloud is `model_points` under the current pose estimate.
"""
from __future__ import annotations

import numpy as np
from scipy.spatial import cKDTree


def backproject(depth, uv, cam_K):
    """Depth (H,W) metres, pixels uv (N,2) [x,y], cam_K (3,3) -> (n_valid,3) camera-frame points with depth > 0."""
    depth = np.asarray(depth, np.float32)
    uv = np.asarray(uv).astype(np.int64)
    K = np.asarray(cam_K, np.float64)
    H, W = depth.shape
    inb = (uv[:, 0] >= 0) & (uv[:, 0] < W) & (uv[:, 1] >= 0) & (uv[:, 1] < H)
    u, v = uv[inb, 0], uv[inb, 1]
    d = depth[v, u].astype(np.float32)
    ok = np.isfinite(d) & (d > 0)
    u, v, d = u[ok].astype(np.float32), v[ok].astype(np.float32), d[ok]
    fx, fy, cx, cy = (np.float32(K[0, 0]), np.float32(K[1, 1]), np.float32(K[0, 2]), np.float32(K[1, 2]))
    return np.stack([(u - cx) * d / fx, (v - cy) * d / fy, d], axis=1).astype(np.float32)


def rigid_fit(p, q):
    """Proper rotation R and translation t minimising sum ||R p_i + t - q_i||^2 (Kabsch / Umeyama without scale)."""
    p, q = np.asarray(p, np.float64), np.asarray(q, np.float64)
    pm, qm = p.mean(0), q.mean(0)
    S = (q - qm).T @ (p - pm)
    U, _, Vt = np.linalg.svd(S)
    D = np.diag([1.0, 1.0, np.sign(np.linalg.det(U) * np.linalg.det(Vt)) or 1.0])
    R = U @ D @ Vt
    return R, qm - R @ pm


def _evaluate(src, tree, tgt, max_dist):
    d, j = tree.query(src, k=1, distance_upper_bound=max_dist)
    ok = np.isfinite(d)
    n = int(ok.sum())
    fitness = n / max(len(src), 1)
    rmse = float(np.sqrt((d[ok] ** 2).sum() / n)) if n else 0.0
    return ok, j, fitness, rmse


def icp_point_to_point(source, target, init_pose, max_dist=0.01, max_iter=30, rel_fitness=1e-6, rel_rmse=1e-6):
    """Open3D-style point-to-point ICP.  Returns (pose (4,4) float64, dict(fitness, inlier_rmse, iterations, n_corr))."""
    T = np.array(init_pose, np.float64).copy()
    src0, tgt = np.asarray(source, np.float64), np.asarray(target, np.float64)
    info = dict(fitness=0.0, inlier_rmse=0.0, iterations=0, n_corr=0)
    if len(tgt) == 0 or len(src0) == 0:
        return T, info
    tree = cKDTree(tgt)
    src = src0 @ T[:3, :3].T + T[:3, 3]
    ok, j, fit, rmse = _evaluate(src, tree, tgt, max_dist)
    it = 0
    for it in range(1, max_iter + 1):
        if ok.sum() < 3:
            it -= 1
            break
        R, t = rigid_fit(src[ok], tgt[j[ok]])
        U = np.eye(4)
        U[:3, :3], U[:3, 3] = R, t
        T = U @ T
        src = src @ R.T + t
        ok, j, fit2, rmse2 = _evaluate(src, tree, tgt, max_dist)
        done = abs(fit - fit2) < rel_fitness and abs(rmse - rmse2) < rel_rmse
        fit, rmse = fit2, rmse2
        if done:
            break
    info.update(fitness=fit, inlier_rmse=rmse, iterations=it, n_corr=int(ok.sum()))
    return T, info


def icp_refinement(depth, uv, pose, cam_K, model_points, inpaint_depth=False, icp_max_dist=0.01):
    """Call shape of zephyr's icpRefinement as used at online_learning.py:476-479 -> (pose (4,4), info)."""
    if inpaint_depth:
        raise NotImplementedError("inpaint_depth is never enabled by the reference (online_learning.py:478)")
    return icp_point_to_point(model_points, backproject(depth, uv, cam_K), pose, max_dist=icp_max_dist)


def estimate_visib_mask(d_test, d_model, delta, visib_mode="bop19"):
    """bop_toolkit_lib.visibility._estimate_visib_mask (estimate_visib_mask_gt is the same function, online_learning.py:497)."""
    d_test, d_model = np.asarray(d_test, np.float32), np.asarray(d_model, np.float32)
    d_diff = d_model - d_test
    if visib_mode == "bop18":
        return (d_diff <= delta) & (d_test > 0) & (d_model > 0)
    return ((d_diff <= delta) | (d_test == 0)) & (d_model > 0)
