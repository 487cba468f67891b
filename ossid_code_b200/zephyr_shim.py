"""Drop-in stand-ins for the three ``zephyr`` symbols OSSID's scoring path imports.

The reference reaches scoring through
  * ``zephyr.datasets.score_dataset.ScoreDataset(...).getPointNetData(data, return_uv_original=True)``
    (call python/ossid/utils/zephyr_utils.py:31, constructed python/ossid/scripts/online_learning.py:206),
  * ``zephyr.models.pointnet2.PointNet2SSG(dim_point, args, num_class=1)`` (online_learning.py:212-227,
    called zephyr_utils.py:34),
  * ``zephyr.utils.projectPointsUv(pose_hypos, model_points, meta_data)`` (zephyr_utils.py:8,58).
The classes here keep those names, constructor arguments, attributes and in-place side effects
and run the work in libzs.so.  ``install()`` registers them under the ``zephyr.*`` module paths so
that the reference's unmodified ``networkInference`` / ``online_learning.py`` import them.

Scorer substitution (SURVEY.md §8 note 2): the reference's PointNet2SSG is a PointNet++ whose code
and weights are not on disk; the class of that name here is the point-wise MLP + max-pool scorer
BASELINE.json specifies.
"""
from __future__ import annotations

import sys
import types

import numpy as np
import torch

from . import weights as W
from .engine import get_context, poses_to_rt12

_next_weight_slot = [0]


def _meta_f(meta):
    return {k: float(np.asarray(v)) for k, v in meta.items()}


class UvOriginal:
    """``uv_original`` as the reference consumes it, without moving it: the reference keeps the (n, N_pts, 2) pixel
    indices of ALL hypotheses only to read ``to_np(uv_original)[pred_idx]`` for the winner
    (python/ossid/scripts/online_learning.py:474-478).  This object stays on the device (int32); ``.detach()``,
    ``.cpu()`` and ``.numpy()`` -- the chain ``to_np`` applies (python/ossid/utils/__init__.py:166-173) -- are no-ops,
    indexing copies just the requested rows to the host as int64, and ``np.asarray(...)`` / ``.to(...)`` /
    ``.tensor()`` materialise the whole array for callers that really want it."""

    dtype = np.dtype(np.int64)

    def __init__(self, dev_i32: torch.Tensor):
        self._dev = dev_i32

    shape = property(lambda self: tuple(self._dev.shape))
    ndim = property(lambda self: self._dev.dim())
    device = property(lambda self: self._dev.device)

    def __len__(self):
        return self._dev.shape[0]

    def detach(self):
        return self

    def cpu(self):
        return self

    def numpy(self):
        return self

    def tensor(self) -> torch.Tensor:
        return self._dev.to(torch.int64)

    def to(self, *args, **kwargs):
        return self.tensor().to(*args, **kwargs)

    def __getitem__(self, idx):
        if torch.is_tensor(idx):
            idx = idx.to(self._dev.device)
        elif isinstance(idx, np.ndarray):
            idx = torch.from_numpy(idx).to(self._dev.device)
        elif isinstance(idx, np.generic):
            idx = idx.item()
        return self._dev[idx].to(torch.int64).cpu().numpy()

    def __array__(self, dtype=None, copy=None):
        a = self.tensor().cpu().numpy()
        return a if dtype is None else a.astype(dtype)


class ScoreDataset:
    """``ScoreDataset(datapoints, dataset_root, dataset_name, args, mode='test')`` -- featuriser only.

    Honoured ``args`` fields: ``inconst_ratio_th`` (online_learning.py:195), plus two of this
    build: ``zs_precision`` ("fp32" (default) -> float32 features, scored to 1e-4 by the 3-term bf16-split tcgen05
    scorer, so that ``scores.argmax()`` (online_learning.py:466-467) is the fp32 argmax; "bf16" -> bfloat16 features +
    bf16 tcgen05 scorer, 1e-2, three times the throughput) and ``zs_device``.
    """

    dim_point = W.DIM_POINT
    gpu_frontend = True      # accepts data['img_u8'] (uint8 camera frame) instead of data['img']

    def __init__(self, datapoints=None, dataset_root="", dataset_name="", args=None, mode="test"):
        self.args, self.mode, self.dataset_name = args, mode, dataset_name
        th = getattr(args, "inconst_ratio_th", None)
        self.inconst_ratio_th = 100.0 if th is None else float(th)
        prec = getattr(args, "zs_precision", "fp32")
        if prec not in ("fp32", "bf16"):
            raise ValueError(f"zs_precision must be 'fp32' or 'bf16', got {prec!r}")
        self.feature_dtype = torch.float32 if prec == "fp32" else torch.bfloat16
        self.device = getattr(args, "zs_device", 0)
        self.return_masks = False
        self.lazy_uv = bool(getattr(args, "zs_lazy_uv", True))     # False: uv_original is a torch.int64 CUDA tensor
        self.last_mask = self.last_keep = None

    def getPointNetData(self, data, return_uv_original=False):
        ctx = get_context(self.device)
        meta = _meta_f(data["meta_data"])
        if "img_u8" in data:
            ctx.set_frame_u8(data["img_u8"], data["depth"], meta, blur=True)
        else:
            ctx.set_frame(data["img"], data["depth"], meta)
        ctx.set_object(0, data["model_points"], data["model_colors"], data["model_normals"])
        transforms = torch.as_tensor(data["transforms"])
        poses12 = poses_to_rt12(transforms, ctx.device)
        M, N = poses12.shape[0], ctx.obj_npts[0]
        keep = None
        if self.inconst_ratio_th < 100 and M > 0:
            keep = ctx.filter(ctx.violations(0, poses12), N, self.inconst_ratio_th)
        feat, uv, mask, _ = ctx.features(0, poses12, keep_idx=keep, dtype=self.feature_dtype,
                                         want_uv=return_uv_original, want_mask=self.return_masks)
        self.last_mask, self.last_keep = mask, keep
        if keep is not None:
            # the reference reads the scored subset back from the dict (zephyr_utils.py:39-43)
            kc = keep.to(torch.int64).cpu()
            data["transforms"] = transforms[kc.to(transforms.device)]
            pe = data.get("pp_err")
            if pe is not None:
                data["pp_err"] = pe[kc.to(pe.device)] if torch.is_tensor(pe) else np.asarray(pe)[kc.numpy()]
        if return_uv_original:
            return feat, (UvOriginal(uv) if self.lazy_uv else uv.to(torch.int64))
        return feat


class PointNet2SSG(torch.nn.Module):
    """``PointNet2SSG(dim_point, args, num_class=1)``: holds the scorer's parameters; forward runs in libzs.so."""

    def __init__(self, dim_point=W.DIM_POINT, args=None, num_class=1):
        super().__init__()
        if dim_point != W.DIM_POINT or num_class != 1:
            raise ValueError(f"kernels are built for dim_point={W.DIM_POINT}, num_class=1")
        dims = W.layer_dims(dim_point, num_class)
        nn = torch.nn
        self.conv1, self.conv2, self.conv3 = (nn.Conv1d(ci, co, 1) for (co, ci) in dims[:3])
        self.fc1, self.fc2, self.fc3 = (nn.Linear(ci, co) for (co, ci) in dims[3:])
        self.bn1, self.bn2, self.bn3 = (nn.BatchNorm1d(co) for (co, _) in dims[:3])
        self.bn_fc1, self.bn_fc2 = (nn.BatchNorm1d(co) for (co, _) in dims[3:5])
        self._slot = _next_weight_slot[0] % 4
        _next_weight_slot[0] += 1
        self._uploaded = None
        self._token = object()
        # float32 point_x: scored by the 3-term bf16-split tcgen05 kernel (1e-4); zs_fp32_kernel="cuda" selects the
        # CUDA-core fp32 kernel that implements the same arithmetic without tensor cores
        self.fp32_on_cuda_cores = getattr(args, "zs_fp32_kernel", "tensor") == "cuda"
        self.eval()

    @property
    def device(self):
        return self.conv1.weight.device

    def load_state_dict(self, state_dict, strict=True):
        sd = {k: (v.reshape(v.shape[0], -1, 1) if k.startswith("conv") and k.endswith("weight") and v.ndim == 2 else v)
              for k, v in state_dict.items()}
        self._uploaded = None
        return super().load_state_dict(sd, strict=strict)

    def _sync_weights(self, ctx):
        # re-upload when the parameters changed or when another user of the context (a model instance, a FrameScorer)
        # has taken this slot meanwhile: the context credits each slot to the token of its last uploader
        key = (ctx.index, tuple(p._version for p in self.state_dict().values()))
        if self._uploaded != key or ctx.weight_owner.get(self._slot) is not self._token:
            ctx.set_weights(self._slot, W.fold_state_dict(self.state_dict()), owner=self._token)
            self._uploaded = key

    def forward(self, batch):
        if self.training:
            raise RuntimeError("inference only: call .eval() (BatchNorm is folded)")
        x = batch["point_x"]
        if not x.is_cuda:
            raise RuntimeError("point_x must be a CUDA tensor: there is no CPU scoring path")
        ctx = get_context(x.device)
        self._sync_weights(ctx)
        if x.dtype == torch.float32 and x.ndim == 3 and not self.fp32_on_cuda_cores:
            x = ctx.split_features(x)          # fp32-accurate path on the tensor cores (3-term bf16 split)
        return ctx.score(self._slot, x).reshape(-1, 1)


def projectPointsUv(pose_hypos, model_points, meta_data, device=0):
    """(M,4,4), (N,3), K2meta dict -> (M,N,2) int64 numpy, [...,0]=x/col, [...,1]=y/row (zephyr_utils.py:58-66)."""
    ctx = get_context(device)
    uv = ctx.project_uv(poses_to_rt12(pose_hypos, ctx.device), model_points, _meta_f(meta_data))
    return uv.to(torch.int64).cpu().numpy()


def install():
    """Register ``zephyr``, ``zephyr.utils`` (+ ``.metrics``, ``.icp``), ``zephyr.datasets.score_dataset``,
    ``zephyr.models.pointnet2`` -- the symbols online_learning.py:28-36 imports that this build provides."""
    def mod(name):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            sys.modules[name] = m
        return m

    z, zu = mod("zephyr"), mod("zephyr.utils")
    zd, zds = mod("zephyr.datasets"), mod("zephyr.datasets.score_dataset")
    zm, zmp = mod("zephyr.models"), mod("zephyr.models.pointnet2")
    from . import icp as _icp, metrics as _metrics
    from .zephyr_utils import K2meta
    zu.projectPointsUv = projectPointsUv
    zu.K2meta = K2meta
    zum, zui = mod("zephyr.utils.metrics"), mod("zephyr.utils.icp")
    zum.add, zum.adi = _metrics.add, _metrics.adi                  # online_learning.py:32
    zui.icpRefinement = _icp.icpRefinement                         # online_learning.py:36
    zu.metrics, zu.icp = zum, zui
    zds.ScoreDataset = ScoreDataset
    zmp.PointNet2SSG = PointNet2SSG
    z.utils, z.datasets, z.models = zu, zd, zm
    zd.score_dataset, zm.pointnet2 = zds, zmp
    return z
