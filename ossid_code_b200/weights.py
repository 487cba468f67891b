"""Scorer weights: layer shapes, seeded initialisation, BatchNorm folding, state_dict ingest.

The reference loads scorer weights as a Lightning ``state_dict``
(python/ossid/scripts/online_learning.py:213-214) for ``PointNet2SSG``; no
checkpoint is on disk (python/ossid/ckpts/.gitignore:1-3).  The scorer built
here is the point-wise MLP + max-pool network ``BASELINE.json`` names
(SURVEY.md §8 note 2): shared MLP ``dim_point -> 64 -> 128 -> 1024`` (1x1 conv +
BatchNorm + ReLU), max over points, head ``1024 -> 512 -> 256 -> num_class``.
BatchNorm is folded into the preceding affine map for inference (eval mode).
"""
from __future__ import annotations

from collections import OrderedDict

import torch

DIM_POINT = 8
POINT_CHANNELS = (64, 128, 1024)
HEAD_CHANNELS = (512, 256)
BN_EPS = 1e-5

# (name, has_bn) in forward order; weight shape is (out, in).
_LAYERS = (("conv1", True), ("conv2", True), ("conv3", True), ("fc1", True), ("fc2", True), ("fc3", False))
FOLDED_KEYS = ("W1", "b1", "W2", "b2", "W3", "b3", "F1", "c1", "F2", "c2", "F3", "c3")


def layer_dims(dim_point: int = DIM_POINT, num_class: int = 1):
    dims = [dim_point, *POINT_CHANNELS, *HEAD_CHANNELS, num_class]
    return [(dims[i + 1], dims[i]) for i in range(len(dims) - 1)]


def seeded_state_dict(seed: int = 0, dim_point: int = DIM_POINT, num_class: int = 1) -> "OrderedDict[str, torch.Tensor]":
    """Deterministic trained-looking weights (He-normal affine maps, non-trivial BN statistics)."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for (name, has_bn), (co, ci) in zip(_LAYERS, layer_dims(dim_point, num_class)):
        sd[f"{name}.weight"] = torch.randn(co, ci, generator=g) * (2.0 / ci) ** 0.5
        sd[f"{name}.bias"] = torch.randn(co, generator=g) * 0.05
        if has_bn:
            bn = name.replace("conv", "bn").replace("fc", "bn_fc")
            sd[f"{bn}.weight"] = 1.0 + 0.1 * torch.randn(co, generator=g)
            sd[f"{bn}.bias"] = 0.05 * torch.randn(co, generator=g)
            sd[f"{bn}.running_mean"] = 0.1 * torch.randn(co, generator=g)
            sd[f"{bn}.running_var"] = 1.0 + 0.2 * torch.rand(co, generator=g)
            sd[f"{bn}.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    return sd


def fold_state_dict(sd) -> "OrderedDict[str, torch.Tensor]":
    """Fold eval-mode BatchNorm into each affine map -> the 12 fp32 tensors the kernels consume.

    Accepts keys with an arbitrary common prefix (Lightning checkpoints prefix
    module paths) and 1x1-conv weights shaped (out, in, 1).
    """
    def find(suffix):
        hits = [k for k in sd if k == suffix or k.endswith("." + suffix)]
        if len(hits) != 1:
            raise KeyError(f"state_dict needs exactly one key ending in '{suffix}', found {hits}")
        return sd[hits[0]].detach().to(torch.float64).cpu()

    out = OrderedDict()
    names = (("W1", "b1"), ("W2", "b2"), ("W3", "b3"), ("F1", "c1"), ("F2", "c2"), ("F3", "c3"))
    for (name, has_bn), (wk, bk) in zip(_LAYERS, names):
        W = find(f"{name}.weight")
        W = W.reshape(W.shape[0], -1)
        b = find(f"{name}.bias")
        if has_bn:
            bn = name.replace("conv", "bn").replace("fc", "bn_fc")
            scale = find(f"{bn}.weight") / torch.sqrt(find(f"{bn}.running_var") + BN_EPS)
            W = W * scale[:, None]
            b = (b - find(f"{bn}.running_mean")) * scale + find(f"{bn}.bias")
        out[wk] = W.to(torch.float32).contiguous()
        out[bk] = b.to(torch.float32).contiguous()
    return out


def seeded_folded(seed: int = 0, dim_point: int = DIM_POINT, num_class: int = 1):
    return fold_state_dict(seeded_state_dict(seed, dim_point, num_class))
