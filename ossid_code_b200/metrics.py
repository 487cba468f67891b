"""Pose-error metrics with the call shape of ``zephyr.utils.metrics`` (imported at
python/ossid/scripts/online_learning.py:32 and used at :337-339,452,482).

``add`` / ``adi`` keep the single-pose signature ``(R_est, t_est, R_gt, t_gt, pts) -> float``; ``pose_errors`` is the
batched form that replaces the reference's per-hypothesis Python loop with one kernel launch (``zs_pose_errors``).
"""
from __future__ import annotations

import numpy as np
import torch

from .engine import get_context, poses_to_rt12


def _mat(R, t):
    m = np.eye(4)
    m[:3, :3], m[:3, 3] = np.asarray(R, np.float64), np.asarray(t, np.float64).reshape(3)
    return m


def pose_errors(pose_hypos, mat_gt, model_points, symmetric: bool = False, device=0) -> np.ndarray:
    """(M,4,4) hypotheses, (4,4) ground truth, (N,3) points -> (M,) ADD (or ADI when ``symmetric``) in metres."""
    ctx = get_context(device)
    err = ctx.pose_errors(poses_to_rt12(pose_hypos, ctx.device), mat_gt, model_points, symmetric)
    return err.cpu().numpy()


def add(R_est, t_est, R_gt, t_gt, pts, device=0) -> float:
    """Average distance of model points (BOP ADD)."""
    return float(pose_errors(_mat(R_est, t_est)[None], _mat(R_gt, t_gt), pts, False, device)[0])


def adi(R_est, t_est, R_gt, t_gt, pts, device=0) -> float:
    """Average distance to the nearest ground-truth model point (BOP ADI, symmetric objects)."""
    return float(pose_errors(_mat(R_est, t_est)[None], _mat(R_gt, t_gt), pts, True, device)[0])
