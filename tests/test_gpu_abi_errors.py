"""Error convention of the C ABI on a live device: negative status + message, no exceptions, no crashes.  pytest -m gpu."""
import ctypes as C

import pytest
import torch

from ossid_code_b200 import _lib, weights
from ossid_code_b200.engine import get_context

pytestmark = pytest.mark.gpu


def test_status_codes_and_messages():
    ctx = get_context(0)
    lib, h = ctx.lib, ctx.h
    st = C.c_void_p(0)
    x = torch.zeros(4, 16, 8, device=ctx.device)
    out = torch.zeros(4, device=ctx.device)
    # unset slots
    assert lib.zs_score(h, 3, x.data_ptr(), 0, 4, 16, 0, out.data_ptr(), st) == -3           # ZS_ERR_STATE
    assert b"weight slot 3" in lib.zs_last_error(h)
    assert lib.zs_violations(h, 63, x.data_ptr(), 1, None, 0.5, out.data_ptr(), st) == -3
    # bad arguments
    ctx.set_weights(0, weights.seeded_folded(0))
    assert lib.zs_score(h, 0, x.data_ptr(), 0, 4, 16, 1, out.data_ptr(), st) == -4           # precision / dtype mismatch
    assert lib.zs_score(h, 0, x.data_ptr() + 4, 0, 4, 16, 0, out.data_ptr(), st) == -1        # misaligned
    assert lib.zs_score(h, 0, None, 0, 4, 16, 0, out.data_ptr(), st) == -1
    assert lib.zs_topk(h, out.data_ptr(), 4, 65, 0, None, out.data_ptr(), out.data_ptr(), st) == -1   # k > ZS_MAX_TOPK
    assert lib.zs_set_weights(h, 0, x.data_ptr(), 7, st) == -1 and b"expected" in lib.zs_last_error(h)
    assert lib.zs_set_frame(h, None, None, 10, 10, 1.0, 1.0, 1.0, 1.0, 1.0, st) == -1
    assert lib.zs_set_frame(h, x.data_ptr(), x.data_ptr(), 2, 2, 1.0, 1.0, 1.0, 1.0, 0.0, st) == -1   # camera_scale <= 0
    assert lib.zs_set_object(h, 64, x.data_ptr(), x.data_ptr(), x.data_ptr(), 4, st) == -1
    # empty work is fine
    assert lib.zs_score(h, 0, None, 0, 0, 16, 0, None, st) == 0
    assert lib.zs_features(h, 0, None, None, 0, None, 0, None, None, None, st) == 0
    # null context never dereferences
    assert lib.zs_score(None, 0, x.data_ptr(), 0, 4, 16, 0, out.data_ptr(), st) == -1
    assert lib.zs_last_error(None) == b"null context"
    # the context still works afterwards
    assert torch.isfinite(ctx.score(0, x)).all()


def test_python_layer_raises_with_message():
    ctx = get_context(0)
    with pytest.raises(_lib.ZsError, match="weight slot"):
        ctx.score(3, torch.zeros(2, 8, 8, device=ctx.device))
    with pytest.raises(ValueError):
        ctx.score(0, torch.zeros(2, 8, 7, device=ctx.device))
