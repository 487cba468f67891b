"""Torch-facing wrapper over the C ABI: one ``ZsContext`` per CUDA device.

PyTorch is used for device memory and streams only; every computation is a call
into libzs.so on ``torch.cuda.current_stream()``.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import ZS_BF16, ZS_BF16_SPLIT, ZS_F32, ZS_F64, check

FEAT_DTYPES = {torch.float32: ZS_F32, torch.bfloat16: ZS_BF16}
POSE_DTYPES = {torch.float32: ZS_F32, torch.float64: ZS_F64}


def feat_code(feat: torch.Tensor) -> int:
    """Feature tensors: (n,N,8) float32 / bfloat16, or (n,2,N,8) bfloat16 = split-bf16 planes (hi, lo)."""
    if feat.ndim == 4:
        if feat.dtype != torch.bfloat16 or feat.shape[1] != 2 or feat.shape[3] != 8:
            raise ValueError(f"split features must be (n,2,N,8) bfloat16, got {tuple(feat.shape)} {feat.dtype}")
        return ZS_BF16_SPLIT
    if feat.ndim != 3 or feat.shape[2] != 8 or feat.dtype not in FEAT_DTYPES:
        raise ValueError(f"point_x must be (n,N,8) float32/bfloat16, got {tuple(feat.shape)} {feat.dtype}")
    return FEAT_DTYPES[feat.dtype]


def split_bf16(x: torch.Tensor) -> torch.Tensor:
    """(n,N,8) float32 -> (n,2,N,8) bfloat16: hi = bf16(x), lo = bf16(x - hi) (what zs_features writes for ZS_BF16_SPLIT)."""
    hi = x.to(torch.bfloat16)
    lo = (x - hi.to(torch.float32)).to(torch.bfloat16)
    return torch.stack([hi, lo], dim=1).contiguous()


def _dev_f32(x, device) -> torch.Tensor:
    """Single cast to float32 (on the host when the source is host memory), then to the device."""
    t = torch.as_tensor(x)
    if t.device.type == "cpu":
        t = t.to(torch.float32).contiguous()
        return t.to(device, non_blocking=True)
    return t.to(device=device, dtype=torch.float32).contiguous()


def poses_to_rt12(transforms, device) -> torch.Tensor:
    """(M,4,4) any float dtype -> (M,12) float32 rows of (R|t) on ``device`` (include/zs.h).

    The whole (M,16) block is copied as is (asynchronously when the source is pinned) and sliced / cast
    once to float32 by ``zs_pack_poses`` on the device; IEEE round-to-nearest either side, so the values equal a
    host cast.
    """
    t = torch.as_tensor(transforms)
    if t.ndim != 3 or t.shape[1:] != (4, 4):
        raise ValueError(f"pose hypotheses must be (M,4,4), got {tuple(t.shape)}")
    if t.dtype not in (torch.float32, torch.float64):
        t = t.to(torch.float64)
    dev = torch.device(device)
    if dev.type == "cpu":
        return t[:, :3, :4].to(torch.float32).reshape(t.shape[0], 12).contiguous()
    t = t.contiguous().to(dev, non_blocking=True)
    return get_context(dev).pack_poses(t)


class ZsContext:
    """Owns a ``zs_ctx`` on one device."""

    def __init__(self, device: int = 0):
        if not torch.cuda.is_available():
            raise _lib.ZsError("CUDA device required: this path has no CPU fallback")
        self.lib = _lib.load()
        self.index = torch.device("cuda", device).index if not isinstance(device, int) else device
        self.device = torch.device("cuda", self.index)
        h = C.c_void_p()
        rc = self.lib.zs_create(C.byref(h), self.index)
        if rc != 0:
            raise _lib.ZsError(f"zs_create(device={self.index}) failed: {self.lib.zs_strerror(rc).decode()}")
        self.h = h
        self.obj_token = {}
        self.frame_hw = None
        self.obj_npts = {}
        self.weight_owner = {}        # slot -> token of whoever uploaded last (see set_weights)
        self.graph_launches = 0       # kernels executed by CUDA-graph replays (FrameScorer), not seen by zs_launch_count

    def close(self):
        if getattr(self, "h", None):
            self.lib.zs_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- plumbing ---------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _ck(self, rc, what):
        check(self.h, rc, what)

    @property
    def launches(self) -> int:
        """Kernels launched through this context: the library's count plus the kernels of replayed CUDA graphs."""
        return int(self.lib.zs_launch_count(self.h)) + self.graph_launches

    @property
    def generation(self) -> int:
        """``zs_alloc_generation``: changes when a replayed launch sequence would no longer be valid."""
        return int(self.lib.zs_alloc_generation(self.h))

    # -- uploads ----------------------------------------------------------------------
    def set_frame(self, img01, depth, meta):
        """img01 (H,W,3) in [0,1] (already blurred, /255), depth (H,W), meta = K2meta dict."""
        rgb, dep = _dev_f32(img01, self.device), _dev_f32(depth, self.device)
        H, W = dep.shape
        if rgb.shape != (H, W, 3):
            raise ValueError(f"img {tuple(rgb.shape)} does not match depth {tuple(dep.shape)}")
        self._ck(self.lib.zs_set_frame(self.h, rgb.data_ptr(), dep.data_ptr(), H, W,
                                       float(meta["camera_fx"]), float(meta["camera_fy"]),
                                       float(meta["camera_cx"]), float(meta["camera_cy"]),
                                       float(meta.get("camera_scale", 1.0)), self._stream()), "zs_set_frame")
        self.frame_hw, self._keep = (H, W), (rgb, dep)

    def set_frame_u8(self, img_u8, depth, meta, blur: bool = True):
        """img_u8 (H,W,3) uint8 straight from the camera; blur + /255 + HSV run on the GPU."""
        img = torch.as_tensor(img_u8)
        if img.dtype != torch.uint8:
            raise ValueError("img_u8 must be uint8")
        img = img.contiguous().to(self.device, non_blocking=True)
        dep = _dev_f32(depth, self.device)
        H, W = dep.shape
        if img.shape != (H, W, 3):
            raise ValueError(f"img {tuple(img.shape)} does not match depth {tuple(dep.shape)}")
        self._ck(self.lib.zs_set_frame_u8(self.h, img.data_ptr(), dep.data_ptr(), H, W,
                                          float(meta["camera_fx"]), float(meta["camera_fy"]),
                                          float(meta["camera_cx"]), float(meta["camera_cy"]),
                                          float(meta.get("camera_scale", 1.0)), int(bool(blur)), self._stream()),
                 "zs_set_frame_u8")
        self.frame_hw, self._keep = (H, W), (img, dep)

    def set_object(self, slot: int, points, colors, normals, token=None):
        """Upload a model cloud into ``slot``.  ``token``: any object identifying the uploaded asset; a later call with
        the same token (``is``) while the slot still holds it is skipped - clouds are static assets that the reference
        loads once (online_learning.py:303-311)."""
        if token is not None and self.obj_token.get(slot) is token:
            return
        self.obj_token.pop(slot, None)          # the slot is only credited with the new cloud once the upload succeeded
        p, c, n = (_dev_f32(t, self.device) for t in (points, colors, normals))
        if p.ndim != 2 or p.shape[1] != 3 or c.shape != p.shape or n.shape != p.shape:
            raise ValueError("model_points/colors/normals must all be (N,3)")
        self._ck(self.lib.zs_set_object(self.h, slot, p.data_ptr(), c.data_ptr(), n.data_ptr(), p.shape[0],
                                        self._stream()), "zs_set_object")
        self.obj_npts[slot] = p.shape[0]
        self.obj_token[slot] = token
        self._keep_obj = (p, c, n)

    def set_weights(self, slot: int, folded: dict, owner=None):
        """Upload folded weights into ``slot``.  ``owner``: token of the uploader; users that share a context check
        ``weight_owner[slot] is their token`` before every launch and re-upload when someone else took the slot."""
        from .weights import FOLDED_KEYS
        self.weight_owner.pop(slot, None)
        blob = torch.cat([folded[k].detach().to(torch.float32).reshape(-1).cpu() for k in FOLDED_KEYS])
        if blob.numel() != _lib.ZS_WEIGHT_FLOATS:
            raise ValueError(f"folded weights have {blob.numel()} values, expected {_lib.ZS_WEIGHT_FLOATS}")
        blob = blob.to(self.device)
        self._ck(self.lib.zs_set_weights(self.h, slot, blob.data_ptr(), blob.numel(), self._stream()), "zs_set_weights")
        torch.cuda.current_stream(self.device).synchronize()   # blob may be freed after return
        self.weight_owner[slot] = owner

    def reserve(self, max_hypotheses: int):
        """Size the library's scratch once (no allocation between the kernels of later frames)."""
        self._ck(self.lib.zs_reserve(self.h, int(max_hypotheses)), "zs_reserve")

    def pack_poses(self, transforms: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Device (n,4,4) float32 / float64 -> (n,12) float32 rows of (R|t), one kernel."""
        n = transforms.shape[0]
        if not transforms.is_cuda or not transforms.is_contiguous() or transforms.dtype not in POSE_DTYPES:
            raise ValueError("pack_poses needs a contiguous float32/float64 CUDA tensor (n,4,4)")
        if out is None:
            out = torch.empty((n, 12), dtype=torch.float32, device=self.device)
        self._ck(self.lib.zs_pack_poses(self.h, transforms.data_ptr() if n else None, POSE_DTYPES[transforms.dtype], n,
                                        out.data_ptr() if n else None, self._stream()), "zs_pack_poses")
        return out

    # -- kernels ----------------------------------------------------------------------
    def project_uv(self, poses12, points, meta) -> torch.Tensor:
        pts = _dev_f32(points, self.device)
        n = poses12.shape[0]
        uv = torch.empty((n, pts.shape[0], 2), dtype=torch.int32, device=self.device)
        self._ck(self.lib.zs_project_uv(self.h, poses12.data_ptr(), n, pts.data_ptr(), pts.shape[0],
                                        float(meta["camera_fx"]), float(meta["camera_fy"]),
                                        float(meta["camera_cx"]), float(meta["camera_cy"]),
                                        uv.data_ptr(), self._stream()), "zs_project_uv")
        return uv

    def mask_count(self, poses12, points, meta, mask) -> torch.Tensor:
        pts = _dev_f32(points, self.device)
        m = torch.as_tensor(mask)
        m = (m != 0).to(torch.uint8).contiguous().to(self.device)
        n = poses12.shape[0]
        cnt = torch.empty((n,), dtype=torch.int32, device=self.device)
        self._ck(self.lib.zs_mask_count(self.h, poses12.data_ptr(), n, pts.data_ptr(), pts.shape[0],
                                        float(meta["camera_fx"]), float(meta["camera_fy"]),
                                        float(meta["camera_cx"]), float(meta["camera_cy"]),
                                        m.data_ptr(), m.shape[0], m.shape[1], cnt.data_ptr(), self._stream()),
                 "zs_mask_count")
        return cnt

    def violations(self, slot: int, poses12, out: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None,
                   mask_th: float = 0.5) -> torch.Tensor:
        """Free-space violation count per hypothesis; with ``mask`` (device uint8 (H,W)) the mask-overlap test of
        ``filterHypoByMask`` runs in the same pass and failing hypotheses report ``ZS_VIOL_MASKED``."""
        n = poses12.shape[0]
        viol = out if out is not None else torch.empty((n,), dtype=torch.int32, device=self.device)
        if mask is not None and (not mask.is_cuda or mask.dtype != torch.uint8 or tuple(mask.shape) != tuple(self.frame_hw)
                                 or not mask.is_contiguous()):
            raise ValueError("mask must be a contiguous uint8 CUDA tensor of the frame's shape")
        self._ck(self.lib.zs_violations(self.h, slot, poses12.data_ptr(), n, mask.data_ptr() if mask is not None else None,
                                        float(mask_th), viol.data_ptr(), self._stream()), "zs_violations")
        return viol

    def boxes_to_mask(self, boxes, scores, expand_ratio: float = 1.2, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """DTOID boxes (n,4) x1 y1 x2 y2 + scores (n,) -> uint8 (H,W) mask on the device (online_learning.py:389-405),
        against the resident frame's depth."""
        b = np.ascontiguousarray(np.asarray(boxes, dtype=np.float64).reshape(-1, 4))
        sc = np.ascontiguousarray(np.asarray(scores, dtype=np.float64).reshape(-1))
        if len(b) != len(sc):
            raise ValueError("boxes and scores differ in length")
        H, W = self.frame_hw
        mask = out if out is not None else torch.empty((H, W), dtype=torch.uint8, device=self.device)
        self._ck(self.lib.zs_boxes_to_mask(self.h, b.ctypes.data_as(C.c_void_p), sc.ctypes.data_as(C.c_void_p), len(b),
                                           float(expand_ratio), mask.data_ptr(), self._stream()), "zs_boxes_to_mask")
        return mask

    def prefilter(self, segments, th: float, mask_th: float = 0.5):
        """``zs_prefilter``: violations + mask test + compaction for every object of a frame in two launches.
        ``segments``: ``(slot, poses12, mask or None, viol_out, keep_out, n_keep_out, info_out or None)`` per object
        (all outputs caller-owned contiguous int32 device views)."""
        segs = list(segments)
        n = len(segs)
        if n == 0:
            return
        ptr = lambda t: t.data_ptr() if t is not None else None
        arr = lambda ctype, vals: (ctype * n)(*vals)
        has_mask = any(sg[2] is not None for sg in segs)
        self._ck(self.lib.zs_prefilter(
            self.h, n, arr(C.c_int32, [sg[0] for sg in segs]), arr(C.c_void_p, [sg[1].data_ptr() for sg in segs]),
            arr(C.c_int32, [sg[1].shape[0] for sg in segs]),
            arr(C.c_void_p, [ptr(sg[2]) for sg in segs]) if has_mask else None, float(mask_th), float(th),
            arr(C.c_void_p, [ptr(sg[3]) for sg in segs]), arr(C.c_void_p, [ptr(sg[4]) for sg in segs]),
            arr(C.c_void_p, [ptr(sg[5]) for sg in segs]), arr(C.c_void_p, [ptr(sg[6]) for sg in segs]),
            self._stream()), "zs_prefilter")

    def filter(self, viol, n_pts: int, th: float) -> torch.Tensor:
        """Kept hypothesis indices (ascending, int32).  Reads the count back: one 4-byte sync."""
        keep, n_keep = self.filter_async(viol, n_pts, th)
        return keep[: int(n_keep.item())]

    def filter_async(self, viol, n_pts: int, th: float, info: Optional[torch.Tensor] = None,
                     keep_out: Optional[torch.Tensor] = None, n_keep_out: Optional[torch.Tensor] = None):
        """As ``filter`` without the read-back: (keep indices, full length; kept count as a device int32[1] tensor).
        ``info``: optional device int32[2] that receives {really kept, fallback violation count} (zs_merge_topk);
        ``keep_out`` / ``n_keep_out``: caller-owned outputs (contiguous int32 views)."""
        n = viol.shape[0]
        keep = keep_out if keep_out is not None else torch.empty((max(n, 1),), dtype=torch.int32, device=self.device)
        n_keep = n_keep_out if n_keep_out is not None else torch.empty((1,), dtype=torch.int32, device=self.device)
        self._ck(self.lib.zs_filter(self.h, viol.data_ptr(), n, n_pts, float(th), keep.data_ptr(),
                                    n_keep.data_ptr(), info.data_ptr() if info is not None else None, self._stream()),
                 "zs_filter")
        return keep, n_keep

    def dynamic_count(self, n_dev: Optional[torch.Tensor], offset: int = 0):
        """``zs_set_dynamic_count``: while set, ``features`` / ``pool`` (bf16) take their count from the device.
        Use as ``with ctx.dynamic_count(n_dev, offset): ...`` so that the host-count mode is always restored."""
        return _DynamicCount(self, n_dev, offset)

    def features(self, slot: int, poses12, keep_idx: Optional[torch.Tensor] = None, n_keep: Optional[int] = None,
                 dtype=torch.float32, want_uv: bool = False, want_mask: bool = False, want_viol: bool = False,
                 out: Optional[torch.Tensor] = None, split: bool = False):
        """``split=True`` (or a 4-D ``out``): split-bf16 features (n,2,N,8) for the fp32-accurate tensor-core scorer."""
        N = self.obj_npts[slot]
        if keep_idx is not None:
            n_keep = keep_idx.shape[0]
        elif n_keep is None:
            n_keep = poses12.shape[0]
        if out is not None:
            feat = out
        elif split:
            feat = torch.empty((n_keep, 2, N, 8), dtype=torch.bfloat16, device=self.device)
        else:
            feat = torch.empty((n_keep, N, 8), dtype=dtype, device=self.device)
        uv = torch.empty((n_keep, N, 2), dtype=torch.int32, device=self.device) if want_uv else None
        mask = torch.empty((n_keep, N), dtype=torch.uint8, device=self.device) if want_mask else None
        viol = torch.empty((n_keep,), dtype=torch.int32, device=self.device) if want_viol else None
        self._ck(self.lib.zs_features(self.h, slot, poses12.data_ptr(),
                                      keep_idx.data_ptr() if keep_idx is not None else None, n_keep,
                                      feat.data_ptr(), feat_code(feat),
                                      uv.data_ptr() if want_uv else None, mask.data_ptr() if want_mask else None,
                                      viol.data_ptr() if want_viol else None, self._stream()), "zs_features")
        return feat, uv, mask, viol

    def score(self, wslot: int, feat: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """feat (n,N,8) bfloat16 -> bf16 tcgen05 path (1e-2); (n,2,N,8) split-bf16 -> fp32-accurate tcgen05 path;
        (n,N,8) float32 -> fp32 CUDA-core path (both 1e-4)."""
        code = feat_code(feat)
        feat = feat.contiguous()
        n, N = feat.shape[0], feat.shape[-2]
        scores = out if out is not None else torch.empty((n,), dtype=torch.float32, device=self.device)
        self._ck(self.lib.zs_score(self.h, wslot, feat.data_ptr(), code, n, N, ZS_BF16 if code == ZS_BF16 else ZS_F32,
                                   scores.data_ptr(), self._stream()), "zs_score")
        return scores

    def features_multi(self, segments):
        """``zs_features_multi``: featurise several objects of the frame in one launch.  ``segments``: iterable of
        ``(slot, poses12, out)`` or ``(slot, poses12, out, keep_idx, n_dev, n_off)``; every ``out`` has the same format
        ((n,N,8) float32 / bfloat16 or split (n,2,N,8)) and as many rows as hypotheses (its capacity with ``n_dev``)."""
        segs = [tuple(sg) + (None,) * (6 - len(sg)) for sg in segments if sg[2].shape[0] > 0]
        if not segs:
            return
        n = len(segs)
        code = feat_code(segs[0][2])
        ptr = lambda t: t.data_ptr() if t is not None else None
        slots = (C.c_int32 * n)(*[sg[0] for sg in segs])
        poses = (C.c_void_p * n)(*[sg[1].data_ptr() for sg in segs])
        outs = (C.c_void_p * n)(*[sg[2].data_ptr() for sg in segs])
        counts = (C.c_int32 * n)(*[sg[2].shape[0] for sg in segs])
        keeps = (C.c_void_p * n)(*[ptr(sg[3]) for sg in segs])
        ndev = (C.c_void_p * n)(*[ptr(sg[4]) for sg in segs])
        noff = (C.c_int32 * n)(*[int(sg[5] or 0) for sg in segs])
        for sg in segs:
            if feat_code(sg[2]) != code:
                raise ValueError("features_multi: every segment must use the same feature format")
        self._ck(self.lib.zs_features_multi(self.h, n, slots, poses, keeps, counts, ndev, noff, outs, code, self._stream()),
                 "zs_features_multi")

    def split_features(self, feat: torch.Tensor) -> torch.Tensor:
        """(n,N,8) float32 CUDA features -> split-bf16 planes (n,2,N,8), one kernel (``zs_split_features``)."""
        if feat.ndim != 3 or feat.shape[2] != 8 or feat.dtype != torch.float32:
            raise ValueError(f"expected (n,N,8) float32 features, got {tuple(feat.shape)} {feat.dtype}")
        feat = feat.contiguous()
        n, N = feat.shape[0], feat.shape[1]
        out = torch.empty((n, 2, N, 8), dtype=torch.bfloat16, device=self.device)
        self._ck(self.lib.zs_split_features(self.h, feat.data_ptr() if n else None, n, N, out.data_ptr() if n else None,
                                            self._stream()), "zs_split_features")
        return out

    def pool(self, wslot: int, feat: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Shared per-point MLP + max over points: (n,N,8) or split (n,2,N,8) -> (n,1024) float32."""
        n, N = feat.shape[0], feat.shape[-2]
        pooled = out if out is not None else torch.empty((n, 1024), dtype=torch.float32, device=self.device)
        self._ck(self.lib.zs_pool(self.h, wslot, feat.data_ptr(), feat_code(feat), n, N, pooled.data_ptr(),
                                  self._stream()), "zs_pool")
        return pooled

    def pool_fused(self, wslot: int, segments, out: torch.Tensor) -> torch.Tensor:
        """``zs_pool_fused``: featurise + shared MLP + max-pool in one kernel.  ``segments``: ``(slot, poses12)`` or
        ``(slot, poses12, keep_idx, n_dev)`` per object, all of one cloud size; with ``keep_idx`` (int32 kept list) the
        segment's capacity is ``len(keep_idx)`` and ``n_dev`` (device int32[1]) says how many of them are live.
        ``out`` (sum of capacities, 1024) float32."""
        segs = [tuple(sg) + (None,) * (4 - len(sg)) for sg in segments]
        segs = [sg for sg in segs if (sg[2] if sg[2] is not None else sg[1]).shape[0] > 0]
        n = len(segs)
        if n == 0:
            return out
        ptr = lambda t: t.data_ptr() if t is not None else None
        caps = [(sg[2] if sg[2] is not None else sg[1]).shape[0] for sg in segs]
        slots = (C.c_int32 * n)(*[sg[0] for sg in segs])
        poses = (C.c_void_p * n)(*[sg[1].data_ptr() for sg in segs])
        keeps = (C.c_void_p * n)(*[ptr(sg[2]) for sg in segs])
        counts = (C.c_int32 * n)(*caps)
        ndev = (C.c_void_p * n)(*[ptr(sg[3]) for sg in segs])
        if out.shape[0] < sum(caps):
            raise ValueError("pool_fused: output has fewer rows than hypotheses")
        self._ck(self.lib.zs_pool_fused(self.h, wslot, n, slots, poses, keeps, counts, ndev, out.data_ptr(), self._stream()),
                 "zs_pool_fused")
        return out

    def head(self, wslot: int, pooled: torch.Tensor, tensor_cores, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """(n,1024) pooled -> (n,) scores; tensor_cores=True / "tf32": tf32 tcgen05 GEMMs (the bf16 path's head);
        "fp32_tc": fp32-accurate 3-term tf32 tcgen05 GEMMs; False / "fp32": fp32 CUDA cores."""
        n = pooled.shape[0]
        scores = out if out is not None else torch.empty((n,), dtype=torch.float32, device=self.device)
        code = {True: ZS_BF16, "tf32": ZS_BF16, "fp32_tc": ZS_BF16_SPLIT, False: ZS_F32, "fp32": ZS_F32}[tensor_cores]
        self._ck(self.lib.zs_head(self.h, wslot, pooled.data_ptr(), n, code, scores.data_ptr(), self._stream()), "zs_head")
        return scores

    def pool_debug(self, wslot: int, feat: torch.Tensor):
        """Tensor-core pool plus the layer-1 / layer-2 activations as the next layer reads them (diagnostic)."""
        n, N = feat.shape[0], feat.shape[-2]
        pooled = torch.zeros((n, 1024), dtype=torch.float32, device=self.device)
        h1 = torch.zeros((n * N, 64), dtype=torch.float32, device=self.device)
        h2 = torch.zeros((n * N, 128), dtype=torch.float32, device=self.device)
        self._ck(self.lib.zs_pool_debug(self.h, wslot, feat.data_ptr(), feat_code(feat), n, N, pooled.data_ptr(), h1.data_ptr(),
                                        h2.data_ptr(), self._stream()), "zs_pool_debug")
        return pooled, h1.view(n, N, 64), h2.view(n, N, 128)

    def pose_errors(self, poses12, gt_pose, points, symmetric: bool) -> torch.Tensor:
        """ADD (symmetric=False) or ADI (True) of every hypothesis against ``gt_pose`` (4,4) -> (n,) float32 metres."""
        pts = _dev_f32(points, self.device)
        gt = poses_to_rt12(torch.as_tensor(gt_pose).reshape(1, 4, 4), self.device)
        n = poses12.shape[0]
        err = torch.empty((n,), dtype=torch.float32, device=self.device)
        self._ck(self.lib.zs_pose_errors(self.h, poses12.data_ptr() if n else None, n, gt.data_ptr(), pts.data_ptr(),
                                         pts.shape[0], int(bool(symmetric)), err.data_ptr(), self._stream()),
                 "zs_pose_errors")
        return err

    def topk(self, scores: torch.Tensor, k: int, index_base: int = 0, index_map: Optional[torch.Tensor] = None):
        """Top-k by (score desc, index asc).  Returned index = index_map[i] (if given) + index_base."""
        s = torch.empty((k,), dtype=torch.float32, device=self.device)
        i = torch.empty((k,), dtype=torch.int32, device=self.device)
        self._ck(self.lib.zs_topk(self.h, scores.data_ptr() if scores.numel() else None, scores.shape[0], k, index_base,
                                  index_map.data_ptr() if index_map is not None and index_map.numel() else None,
                                  s.data_ptr(), i.data_ptr(), self._stream()), "zs_topk")
        return s, i

    def icp_refine(self, poses12: torch.Tensor, src_pts, uv: torch.Tensor, depth=None, meta=None, max_dist: float = 0.01,
                   max_iter: int = 30):
        """Point-to-point ICP of every pose against the depth image (``depth`` None = the resident frame).

        ``uv``: int32 (n_src,2) shared by all poses or (n,n_src,2) per pose.  Returns (poses (n,12) float32,
        stats (n,4) = fitness, inlier_rmse, iterations, correspondences)."""
        src = _dev_f32(src_pts, self.device)
        uv = torch.as_tensor(uv).to(device=self.device, dtype=torch.int32).contiguous()
        n, n_src = poses12.shape[0], src.shape[0]
        per_pose = uv.dim() == 3
        if uv.shape[-2:] != (n_src, 2) or (per_pose and uv.shape[0] != n):
            raise ValueError(f"uv shape {tuple(uv.shape)} does not match {n} poses x {n_src} points")
        out = torch.empty((n, 12), dtype=torch.float32, device=self.device)
        stats = torch.empty((n, 4), dtype=torch.float32, device=self.device)
        H = W = 0
        fx = fy = cx = cy = 0.0
        dptr = None
        if depth is not None:
            d = _dev_f32(depth, self.device)
            H, W = d.shape
            fx, fy, cx, cy = (float(meta[k]) for k in ("camera_fx", "camera_fy", "camera_cx", "camera_cy"))
            scale = float(meta.get("camera_scale", 1.0))
            if scale != 1.0:
                d = d / scale
            dptr = d.data_ptr()
        self._ck(self.lib.zs_icp_refine(self.h, poses12.data_ptr() if n else None, n, src.data_ptr(), n_src, uv.data_ptr(),
                                        int(per_pose), dptr, H, W, fx, fy, cx, cy, float(max_dist), int(max_iter),
                                        out.data_ptr(), stats.data_ptr(), self._stream()), "zs_icp_refine")
        return out, stats

    def visib_mask(self, d_test, d_model, delta: float, bop18: bool = False) -> torch.Tensor:
        """bop_toolkit visibility rule, elementwise -> bool tensor of d_test's shape."""
        a, b = _dev_f32(d_test, self.device), _dev_f32(d_model, self.device)
        if a.shape != b.shape:
            raise ValueError("d_test and d_model must have the same shape")
        out = torch.empty(a.shape, dtype=torch.uint8, device=self.device)
        self._ck(self.lib.zs_visib_mask(self.h, a.data_ptr(), b.data_ptr(), a.numel(), float(delta), int(bool(bop18)),
                                        out.data_ptr(), self._stream()), "zs_visib_mask")
        return out.to(torch.bool)

    def merge_topk(self, gathered: torch.Tensor, n_obj: int, k: int, out=None, poses_out: Optional[torch.Tensor] = None):
        """``zs_merge_topk`` over all-gathered records (world, rec_ints) int32 -> (n_obj,k) scores, indices;
        ``poses_out`` (n_obj,k,12) float32 also receives the winners' poses when the records carry them."""
        world, rec_ints = gathered.shape
        if out is not None:
            s, i = out
        else:
            s = torch.empty((n_obj, k), dtype=torch.float32, device=self.device)
            i = torch.empty((n_obj, k), dtype=torch.int32, device=self.device)
        self._ck(self.lib.zs_merge_topk(self.h, gathered.data_ptr(), world, rec_ints, n_obj, k, s.data_ptr(), i.data_ptr(),
                                        poses_out.data_ptr() if poses_out is not None else None, self._stream()),
                 "zs_merge_topk")
        return s, i

    def gather_poses(self, poses12: torch.Tensor, idx: torch.Tensor, pose_seg: torch.Tensor, out: torch.Tensor):
        """Poses of the candidates: out[o][j] = poses12[pose_seg[o].first + idx[o][j] - pose_seg[o].lo] (zeros for
        idx < 0).  ``idx`` (n_obj,k) int32 global indices, ``pose_seg`` (n_obj,4) int32 {first row, lo, 0, 0}."""
        n_obj, k = idx.shape
        self._ck(self.lib.zs_gather_poses(self.h, poses12.data_ptr(), idx.data_ptr(), pose_seg.data_ptr(), n_obj, k,
                                          out.data_ptr(), self._stream()), "zs_gather_poses")
        return out

    # fp32-accurate scoring of a handful of hypotheses (the re-rank of the top-k candidates): split-bf16 features
    # (n,2,N,8) and the 3-term tcgen05 scorer
    def features_f32a(self, slot: int, poses12: torch.Tensor, out: torch.Tensor):
        return self.features(slot, poses12, out=out)[0]

    def pool_f32a(self, wslot: int, feat: torch.Tensor, out: torch.Tensor):
        return self.pool(wslot, feat, out=out)

    def topk_segments(self, scores: torch.Tensor, segments: torch.Tensor, k: int, index_map: Optional[torch.Tensor] = None,
                      out=None):
        """Per-segment top-k in one launch.  ``segments``: int32 (n_seg,4) device tensor of {first, count, index_base, 0}.
        Returns (n_seg,k) scores and indices (index = index_map[first+i] (if given) + index_base); ``out`` = (s, i)
        preallocated outputs."""
        n_seg = segments.shape[0]
        if out is not None:
            s, i = out
        else:
            s = torch.empty((n_seg, k), dtype=torch.float32, device=self.device)
            i = torch.empty((n_seg, k), dtype=torch.int32, device=self.device)
        self._ck(self.lib.zs_topk_segments(self.h, scores.data_ptr() if scores.numel() else None, segments.data_ptr(), n_seg, k,
                                           index_map.data_ptr() if index_map is not None and index_map.numel() else None,
                                           s.data_ptr(), i.data_ptr(), self._stream()), "zs_topk_segments")
        return s, i


class _DynamicCount:
    def __init__(self, ctx, n_dev, offset):
        self.ctx, self.n_dev, self.offset = ctx, n_dev, int(offset)

    def _set(self, t, off):
        self.ctx._ck(self.ctx.lib.zs_set_dynamic_count(self.ctx.h, t.data_ptr() if t is not None else None, off),
                     "zs_set_dynamic_count")

    def __enter__(self):
        self._set(self.n_dev, self.offset)
        return self

    def __exit__(self, *exc):
        self._set(None, 0)
        return False


_contexts = {}


def get_context(device=0) -> ZsContext:
    idx = torch.device(device).index if not isinstance(device, int) else device
    idx = 0 if idx is None else idx
    if idx not in _contexts:
        _contexts[idx] = ZsContext(idx)
    return _contexts[idx]
