"""ctypes binding of libzs.so (C ABI: include/zs.h).  No CPU fallback: a missing library raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ZS_LIB") or os.path.join(HERE, "libzs.so")   # ZS_LIB: A/B builds in tools/

ZS_F32, ZS_BF16, ZS_BF16_SPLIT, ZS_F64 = 0, 1, 2, 3
ZS_MAX_TOPK = 64
ZS_WEIGHT_FLOATS = 64 * 8 + 64 + 128 * 64 + 128 + 1024 * 128 + 1024 + 512 * 1024 + 512 + 256 * 512 + 256 + 256 + 1

_p, _i, _f = C.c_void_p, C.c_int, C.c_float

# name -> (restype, argtypes); mirrors include/zs.h one to one
SIGNATURES = {
    "zs_version": (_i, []),
    "zs_strerror": (C.c_char_p, [_i]),
    "zs_create": (_i, [C.POINTER(_p), _i]),
    "zs_destroy": (None, [_p]),
    "zs_last_error": (C.c_char_p, [_p]),
    "zs_launch_count": (C.c_int64, [_p]),
    "zs_alloc_generation": (C.c_int64, [_p]),
    "zs_set_frame": (_i, [_p, _p, _p, _i, _i, _f, _f, _f, _f, _f, _p]),
    "zs_set_frame_u8": (_i, [_p, _p, _p, _i, _i, _f, _f, _f, _f, _f, _i, _p]),
    "zs_set_object": (_i, [_p, _i, _p, _p, _p, _i, _p]),
    "zs_set_dynamic_count": (_i, [_p, _p, _i]),
    "zs_set_weights": (_i, [_p, _i, _p, C.c_size_t, _p]),
    "zs_project_uv": (_i, [_p, _p, _i, _p, _i, _f, _f, _f, _f, _p, _p]),
    "zs_mask_count": (_i, [_p, _p, _i, _p, _i, _f, _f, _f, _f, _p, _i, _i, _p, _p]),
    "zs_violations": (_i, [_p, _i, _p, _i, _p, C.c_double, _p, _p]),
    "zs_boxes_to_mask": (_i, [_p, _p, _p, _i, C.c_double, _p, _p]),
    "zs_prefilter": (_i, [_p, _i, _p, _p, _p, _p, C.c_double, _f, _p, _p, _p, _p, _p]),
    "zs_filter": (_i, [_p, _p, _i, _i, _f, _p, _p, _p, _p]),
    "zs_reserve": (_i, [_p, _i]),
    "zs_pack_poses": (_i, [_p, _p, _i, _i, _p, _p]),
    "zs_merge_topk": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p, _p]),
    "zs_gather_poses": (_i, [_p, _p, _p, _p, _i, _i, _p, _p]),
    "zs_features": (_i, [_p, _i, _p, _p, _i, _p, _i, _p, _p, _p, _p]),
    "zs_features_multi": (_i, [_p, _i, _p, _p, _p, _p, _p, _p, _p, _i, _p]),
    "zs_score": (_i, [_p, _i, _p, _i, _i, _i, _i, _p, _p]),
    "zs_split_features": (_i, [_p, _p, _i, _i, _p, _p]),
    "zs_pool": (_i, [_p, _i, _p, _i, _i, _i, _p, _p]),
    "zs_head": (_i, [_p, _i, _p, _i, _i, _p, _p]),
    "zs_pool_fused": (_i, [_p, _i, _i, _p, _p, _p, _p, _p, _p, _p]),
    "zs_pool_debug": (_i, [_p, _i, _p, _i, _i, _i, _p, _p, _p, _p]),
    "zs_pose_errors": (_i, [_p, _p, _i, _p, _p, _i, _i, _p, _p]),
    "zs_topk": (_i, [_p, _p, _i, _i, _i, _p, _p, _p, _p]),
    "zs_topk_segments": (_i, [_p, _p, _p, _i, _i, _p, _p, _p, _p]),
    "zs_icp_refine": (_i, [_p, _p, _i, _p, _i, _p, _i, _p, _i, _i, _f, _f, _f, _f, _f, _i, _p, _p, _p]),
    "zs_visib_mask": (_i, [_p, _p, _p, C.c_size_t, _f, _i, _p, _p]),
}

_lib = None


class ZsError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load libzs.so and bind every symbol of include/zs.h; raises if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ZsError(
            f"{LIB_PATH} is missing: build it with `python -m ossid_code_b200.build` "
            "(there is no CPU or PyTorch fallback for this path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export it
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(ctx, rc: int, what: str):
    if rc != 0:
        lib = load()
        msg = lib.zs_last_error(ctx).decode() if ctx else ""
        raise ZsError(f"{what} failed: {lib.zs_strerror(rc).decode()} ({rc}) {msg}")
