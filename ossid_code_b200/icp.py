"""Post-scoring refinement with the call shapes the reference uses.

* ``icpRefinement(depth, uv, pose, cam_K, model_points, inpaint_depth=False, icp_max_dist=0.01) -> (pose, info)``
  -- ``zephyr.utils.icp.icpRefinement`` as called at python/ossid/scripts/online_learning.py:476-479 (Open3D
  point-to-point ICP of the single winning pose, on the CPU).  Here: ``zs_icp_refine``, one CTA per pose, so
  ``icp_refine_batch`` refines the top-k candidates of every object of a frame in one launch.
* ``estimate_visib_mask_gt(d_test, d_gt, delta, visib_mode='bop19')`` -- ``bop_toolkit_lib.visibility`` as called at
  online_learning.py:500.  The rendered depth ``d_gt`` comes from the reference's renderer, which stays in the
  reference stack.

Parity: against oracle/icp_oracle.py (published algorithms restated; both callees are un-vendored and unpinned).
"""
from __future__ import annotations

import numpy as np
import torch

from .engine import get_context, poses_to_rt12
from .zephyr_utils import K2meta


def _to44(p12: torch.Tensor) -> np.ndarray:
    p = p12.detach().cpu().numpy().astype(np.float64).reshape(-1, 3, 4)
    out = np.tile(np.eye(4), (p.shape[0], 1, 1))
    out[:, :3, :] = p
    return out


def icp_refine_batch(depth, uv, poses, cam_K, model_points, icp_max_dist=0.01, max_iter=30, device=0):
    """(M,4,4) poses, uv (N,2) or (M,N,2) -> (refined (M,4,4) float64, stats (M,4) float32 numpy:
    fitness, inlier_rmse, iterations, correspondences).  ``depth`` None = the frame already resident in the context."""
    ctx = get_context(device)
    meta = None if depth is None else {k: float(v) for k, v in K2meta(np.asarray(cam_K)).items()}
    p12 = poses_to_rt12(torch.as_tensor(np.asarray(poses)).reshape(-1, 4, 4), ctx.device)
    out, stats = ctx.icp_refine(p12, model_points, uv, depth=depth, meta=meta, max_dist=icp_max_dist, max_iter=max_iter)
    return _to44(out), stats.cpu().numpy()


def icpRefinement(depth, uv, pose, cam_K, model_points, inpaint_depth=False, icp_max_dist=0.01, device=0):
    """Drop-in for the reference call (online_learning.py:476-479): returns ``(pose (4,4) float64, info dict)``."""
    if inpaint_depth:
        raise NotImplementedError("inpaint_depth is never enabled by the reference (online_learning.py:478)")
    T, st = icp_refine_batch(depth, np.asarray(uv).reshape(-1, 2), np.asarray(pose).reshape(1, 4, 4), cam_K, model_points,
                             icp_max_dist=icp_max_dist, device=device)
    info = dict(fitness=float(st[0, 0]), inlier_rmse=float(st[0, 1]), iterations=int(st[0, 2]), n_corr=int(st[0, 3]))
    return T[0], info


def estimate_visib_mask_gt(d_test, d_gt, delta, visib_mode="bop19", device=0):
    """Boolean visibility mask of the rendered depth ``d_gt`` in the observed depth ``d_test`` (numpy in, numpy out)."""
    if visib_mode not in ("bop18", "bop19"):
        raise ValueError("Unknown visibility mode.")
    return get_context(device).visib_mask(d_test, d_gt, float(delta), bop18=visib_mode == "bop18").cpu().numpy()
