"""Out-of-bounds WRITE detection without compute-sanitizer (closed on this GPU pool, profiles/r2_sanitizer_unavailable.txt):
every output of the hot-path kernels is carved out of a larger allocation whose guard bands (a recognisable byte pattern
before and after) must come back untouched, on shapes that exercise tail tiles, shifted-back tiles, short hypotheses,
partial chunks and device-side counts.  pytest -m gpu."""
import numpy as np
import pytest
import torch

from ossid_code_b200 import scoring, synthetic as syn, weights, zephyr_utils as glue
from ossid_code_b200.engine import get_context, poses_to_rt12, split_bf16

pytestmark = pytest.mark.gpu
GUARD = 4096          # bytes on either side
PATTERN = 0xA5


class Guarded:
    """A tensor of `shape` / `dtype` inside a byte buffer with PATTERN-filled guard bands."""

    def __init__(self, shape, dtype, device):
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        self.raw = torch.full((GUARD + n + GUARD,), PATTERN, dtype=torch.uint8, device=device)
        self.t = self.raw[GUARD: GUARD + n].view(dtype).view(*shape)

    def intact(self):
        return bool((self.raw[:GUARD] == PATTERN).all()) and bool((self.raw[-GUARD:] == PATTERN).all())


@pytest.fixture(scope="module")
def ctx():
    return get_context(0)


@pytest.mark.parametrize("n_pts,n_hypo", [(1000, 37), (77, 50), (129, 9), (256, 3), (1, 5), (4000, 4)])
def test_feature_kernels_stay_inside_their_outputs(ctx, n_pts, n_hypo):
    sc = syn.make_scene(71, "lmo", n_obj=1, n_pts=n_pts, n_hypo=n_hypo)
    ob = sc["objects"][0]
    ctx.set_frame_u8(sc["img"], sc["depth"], glue.K2meta(sc["cam_K"]))
    ctx.set_object(0, ob["model_points"], ob["model_colors"], ob["model_normals"])
    p12 = poses_to_rt12(ob["pose_hypos"], ctx.device)
    keep = torch.arange(0, n_hypo, 2, dtype=torch.int32, device=ctx.device)
    for shape, dtype, kw in (((n_hypo, n_pts, 8), torch.bfloat16, {}), ((n_hypo, n_pts, 8), torch.float32, {}),
                             ((n_hypo, 2, n_pts, 8), torch.bfloat16, {}), ((len(keep), n_pts, 8), torch.bfloat16, dict(keep_idx=keep))):
        g = Guarded(shape, dtype, ctx.device)
        ctx.features(0, p12, out=g.t, **kw)
        torch.cuda.synchronize()
        assert g.intact(), (shape, dtype)
        g2 = Guarded(shape, dtype, ctx.device)
        ctx.features_multi([(0, p12, g2.t) + ((kw["keep_idx"],) if kw else ())])
        torch.cuda.synchronize()
        assert g2.intact() and torch.equal(g2.t, g.t), (shape, dtype)
    # side outputs
    gf, gu, gm, gv = (Guarded(s, d, ctx.device) for s, d in (((n_hypo, n_pts, 8), torch.float32), ((n_hypo, n_pts, 2), torch.int32),
                                                          ((n_hypo, n_pts), torch.uint8), ((n_hypo,), torch.int32)))
    rc = ctx.lib.zs_features(ctx.h, 0, p12.data_ptr(), None, n_hypo, gf.t.data_ptr(), 0, gu.t.data_ptr(), gm.t.data_ptr(),
                             gv.t.data_ptr(), ctx._stream())
    torch.cuda.synchronize()
    assert rc == 0 and gf.intact() and gu.intact() and gm.intact() and gv.intact()
    gviol = Guarded((n_hypo,), torch.int32, ctx.device)
    ctx.violations(0, p12, out=gviol.t)
    torch.cuda.synchronize()
    assert gviol.intact() and torch.equal(gviol.t, gv.t)


@pytest.mark.parametrize("n,N", [(1, 1), (3, 100), (5, 1000), (149, 257), (2, 127), (75, 384), (19, 1000)])
def test_scorer_kernels_stay_inside_their_outputs(ctx, n, N):
    g = torch.Generator().manual_seed(n * 13 + N)
    x = torch.randn(n, N, 8, generator=g) * 0.5
    ctx.set_weights(1, weights.seeded_folded(2))
    for feat, tc_head in ((x.to(torch.bfloat16).to(ctx.device), True), (split_bf16(x).to(ctx.device), False), (x.to(ctx.device), False)):
        pooled, scores = Guarded((n, 1024), torch.float32, ctx.device), Guarded((n,), torch.float32, ctx.device)
        ctx.pool(1, feat, out=pooled.t)
        ctx.head(1, pooled.t, tc_head, out=scores.t)
        torch.cuda.synchronize()
        assert pooled.intact() and scores.intact(), (feat.shape, feat.dtype)
        assert bool(torch.isfinite(pooled.t).all()) and bool(torch.isfinite(scores.t).all())
    top_s, top_i = Guarded((8,), torch.float32, ctx.device), Guarded((8,), torch.int32, ctx.device)
    ctx.lib.zs_topk(ctx.h, scores.t.data_ptr(), n, 8, 0, None, top_s.t.data_ptr(), top_i.t.data_ptr(), ctx._stream())
    torch.cuda.synchronize()
    assert top_s.intact() and top_i.intact()


@pytest.mark.parametrize("n_pts,n_obj,n_hypo", [(1000, 3, 41), (128, 2, 7), (300, 35, 5)])
def test_fused_kernel_stays_inside_its_output(ctx, n_pts, n_obj, n_hypo):
    sc = syn.make_scene(73, "lmo", n_obj=1, n_pts=n_pts, n_hypo=n_hypo)
    ob = sc["objects"][0]
    ctx.set_frame_u8(sc["img"], sc["depth"], glue.K2meta(sc["cam_K"]))
    ctx.set_weights(1, weights.seeded_folded(3))
    p12 = poses_to_rt12(ob["pose_hypos"], ctx.device)
    for s in range(n_obj):
        ctx.set_object(s, ob["model_points"], ob["model_colors"], ob["model_normals"])
    out = Guarded((n_obj * n_hypo, 1024), torch.float32, ctx.device)
    ctx.pool_fused(1, [(s, p12) for s in range(n_obj)], out=out.t)
    torch.cuda.synchronize()
    assert out.intact() and bool(torch.isfinite(out.t).all())
    assert torch.equal(out.t[:n_hypo], out.t[-n_hypo:])              # the same object and poses in the first and last segment


def test_frame_scorer_record_and_merge_stay_inside_their_buffers(ctx):
    """Top-k record, pose gather, merge and re-rank outputs (the multi-GPU tail) against guard bands."""
    k, n_obj, world = 5, 7, 3
    rec_ints = scoring.record_ints(n_obj, k, True)
    gathered = Guarded((world, rec_ints), torch.int32, ctx.device)
    gathered.t.zero_()
    gathered.t[:, n_obj * k: 2 * n_obj * k] = -1
    S, I, P = (Guarded(s, d, ctx.device) for s, d in (((n_obj, k), torch.float32), ((n_obj, k), torch.int32), ((n_obj, k, 12), torch.float32)))
    ctx.merge_topk(gathered.t, n_obj, k, out=(S.t, I.t), poses_out=P.t)
    torch.cuda.synchronize()
    assert S.intact() and I.intact() and P.intact() and gathered.intact()
    assert bool((I.t == -1).all()) and bool(torch.isinf(S.t).all())
    poses = torch.randn(40, 12, device=ctx.device)
    idx = torch.randint(-1, 10, (n_obj, k), dtype=torch.int32, device=ctx.device)
    seg = torch.tensor([[3 * o, 0, 0, 0] for o in range(n_obj)], dtype=torch.int32, device=ctx.device)
    out = Guarded((n_obj, k, 12), torch.float32, ctx.device)
    ctx.gather_poses(poses, idx, seg, out=out.t)
    torch.cuda.synchronize()
    assert out.intact()
    exp = torch.where(idx[..., None] >= 0, poses[(seg[:, 0:1] + idx.clamp(min=0)).long()], torch.zeros(1, device=ctx.device))
    assert torch.equal(out.t, exp)
