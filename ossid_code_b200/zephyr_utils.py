"""Mirror of the reference's scoring glue, ``ossid.utils.zephyr_utils``.

Same function names, argument meaning, return tuples and in-place contract as
python/ossid/utils/zephyr_utils.py:10-71, written against this package.  The reference's own
file also works unmodified on top of ``zephyr_shim.install()``; this mirror exists so that the
path runs where /root/reference is absent, and adds one fast path: a featuriser that declares
``gpu_frontend`` receives the raw uint8 frame and blurs / normalises it on the GPU (bit-identical
to ``cv2.GaussianBlur(img,(5,5),0)/255.``), so no float64 image is built on the host.
"""
from __future__ import annotations

import time

import numpy as np
import torch


def K2meta(cam_K):
    """python/ossid/utils/__init__.py:148-156."""
    return {"camera_fx": cam_K[0, 0], "camera_fy": cam_K[1, 1],
            "camera_cx": cam_K[0, 2], "camera_cy": cam_K[1, 2], "camera_scale": 1.0}


def to_np(x):
    """python/ossid/utils/__init__.py:166-173."""
    if isinstance(x, (np.ndarray, float, int)):
        return x
    return x.detach().cpu().numpy()


def networkInference(model, dataset, data, return_time=False):
    """Score every pose hypothesis of one (object, frame); see zephyr_utils.py:10-47.

    Returns ``(poses (n,4,4) ndarray, scores ndarray, pp_err, uv_original[, seconds])`` for the
    hypotheses that survive the featuriser's free-space pre-filter, in input order.
    """
    scoring_data = {}
    if getattr(dataset, "gpu_frontend", False):
        scoring_data["img_u8"] = torch.from_numpy(np.ascontiguousarray(data["img"]))
    else:
        import cv2
        scoring_data["img"] = torch.from_numpy(cv2.GaussianBlur(data["img"], (5, 5), 0) / 255.)
    scoring_data["depth"] = torch.from_numpy(data["depth"])
    scoring_data["transforms"] = torch.from_numpy(data["pose_hypos"])
    scoring_data["meta_data"] = K2meta(data["cam_K"])
    for key in ("model_points", "model_colors", "model_normals"):
        scoring_data[key] = torch.from_numpy(data[key])
    scoring_data["pp_err"] = data["pp_err"] if "pp_err" in data else torch.zeros(len(data["pose_hypos"]))

    with torch.no_grad():
        t1 = time.time()
        point_x, uv_original = dataset.getPointNetData(scoring_data, return_uv_original=True)
        pred_score = to_np(model({"point_x": point_x.to(model.device)}))   # to_np synchronises
        inference_time = time.time() - t1

    out = (to_np(scoring_data["transforms"]), pred_score, scoring_data["pp_err"], uv_original)
    return out + (inference_time,) if return_time else out


def filterHypoByMask(model_points, meta_data, pose_hypos, mask, th=0.5, device=0):
    """Keep hypotheses that project more than ``th`` of the model points onto ``mask`` (zephyr_utils.py:49-71).

    Projection, bounds test, mask gather and the per-hypothesis count are one kernel
    (``zs_mask_count``); only the (M,) counts come back.
    """
    from .engine import get_context, poses_to_rt12
    ctx = get_context(device)
    meta = {k: float(np.asarray(v)) for k, v in meta_data.items()}
    cnt = ctx.mask_count(poses_to_rt12(pose_hypos, ctx.device), model_points, meta, mask)
    ratio = cnt.cpu().numpy() / np.asarray(model_points).shape[0]
    return ratio > th
