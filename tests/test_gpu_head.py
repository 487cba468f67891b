"""Tensor-core head (tf32 GEMMs fed by TMA) and the batched multi-object path, vs the oracle.  pytest -m gpu."""
import numpy as np
import pytest
import torch

from oracle import zephyr_oracle as zo
from ossid_code_b200 import scoring, synthetic as syn, weights, zephyr_utils as glue
from ossid_code_b200.engine import get_context

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return get_context(0)


def _head_ref(pooled, w, tf32):
    t = zo._tf32 if tf32 else (lambda x: x)
    g = torch.relu(t(pooled) @ t(w["F1"]).T + w["c1"])
    g = torch.relu(t(g) @ t(w["F2"]).T + w["c2"])
    return (g @ w["F3"].T + w["c3"])[:, 0]


@pytest.mark.parametrize("n", [1, 127, 128, 129, 1000, 1025, 2000, 33000])
def test_head_tensor_core_matches_oracle(ctx, n):
    g = torch.Generator().manual_seed(n)
    pooled = torch.relu(torch.randn(n, 1024, generator=g)) * 2.0          # pooled vectors are post-ReLU
    w = weights.seeded_folded(4)
    ctx.set_weights(2, w)
    dev = pooled.to(ctx.device)
    tc = ctx.head(2, dev, tensor_cores=True).cpu()
    f32 = ctx.head(2, dev, tensor_cores=False).cpu()
    ref32, reftf = _head_ref(pooled, w, False), _head_ref(pooled, w, True)
    scale = float(ref32.abs().max())
    assert float((f32 - ref32).abs().max()) <= 1e-4 * scale + 1e-6, "fp32 CUDA-core head"
    acc = ctx.head(2, dev, tensor_cores="fp32_tc").cpu()          # 3-term tf32 on the tensor cores (CUDA cores below 1,025 rows)
    e_acc = float((acc - ref32).abs().max())
    assert e_acc <= 2e-5 * scale + 1e-6, f"fp32-accurate tensor-core head: {e_acc:.3e} (scale {scale:.2f})"
    e_tf, e_32 = float((tc - reftf).abs().max()), float((tc - ref32).abs().max())
    print(f"n={n}: tf32 head vs tf32 oracle {e_tf:.3e}, vs fp32 oracle {e_32:.3e}, scale {scale:.2f}")
    assert e_tf <= 1e-3 * scale + 1e-6, "tf32 tensor-core head vs tf32-emulating oracle"
    assert e_32 <= 5e-3 * scale + 1e-6, "tf32 tensor-core head vs fp32 oracle"


def test_topk_index_map(ctx):
    s = torch.tensor([0.5, 3.0, 3.0, -1.0], device=ctx.device)
    keep = torch.tensor([10, 20, 30, 40], dtype=torch.int32, device=ctx.device)
    ts, ti = ctx.topk(s, 3, index_base=1000, index_map=keep)
    assert ti.tolist() == [1020, 1030, 1010] and ts.tolist() == [3.0, 3.0, 0.5]


def test_topk_segments_matches_per_object_topk(ctx):
    """One launch over all objects == zs_topk object by object: ties, an empty segment, a segment shorter than k."""
    g = torch.Generator().manual_seed(5)
    counts = [1000, 0, 3, 4097, 64]
    scores = torch.randn(sum(counts), generator=g)
    scores[10] = scores[500] = scores[999] = 9.0                 # ties inside segment 0: lowest index first
    keep = torch.randperm(sum(counts), generator=g).to(torch.int32)
    seg, first = [], 0
    for o, c in enumerate(counts):
        seg.append([first, c, 1000 * o, 0])
        first += c
    dev_s, dev_k = scores.to(ctx.device), keep.to(ctx.device)
    seg_t = torch.tensor(seg, dtype=torch.int32, device=ctx.device)
    for index_map in (None, dev_k):
        S, I = ctx.topk_segments(dev_s, seg_t, 8, index_map=index_map)
        for o, (f, c, base, _) in enumerate(seg):
            es, ei = ctx.topk(dev_s[f:f + c], 8, base, index_map=None if index_map is None else index_map[f:f + c])
            assert torch.equal(S[o], es) and torch.equal(I[o], ei), (o, S[o], es, I[o], ei)
    assert I[1].tolist() == [-1] * 8 and I[2][3:].tolist() == [-1] * 5
    S0, I0 = ctx.topk_segments(dev_s, seg_t, 8)
    assert I0[0][:3].tolist() == [10, 500, 999]


def test_score_frames_stream_equals_frame_by_frame(ctx):
    """BASELINE.json config 5 (multi-frame stream between finetune steps): the pipelined score_frames() over different
    frames returns exactly what score_frame() returns for each frame alone."""
    frames = []
    for seed in (3, 4, 5, 6):
        sc = syn.make_scene(seed, "lmo", n_obj=2, n_pts=384, n_hypo=300)
        frames.append(dict(img=sc["img"], depth=sc["depth"], cam_K=sc["cam_K"], objects=sc["objects"]))
    ws = [weights.seeded_folded(0), weights.seeded_folded(1)]
    fs = scoring.FrameScorer(ws, device=0, precision="bf16", inconst_ratio_th=100.0, k=4)
    stream = fs.score_frames(frames, weight_of=lambda o: o % 2, depth=2)
    assert len(stream) == len(frames)
    for fr, (S, I) in zip(frames, stream):
        S1, I1 = fs.score_frame(fr["img"], fr["depth"], fr["cam_K"], fr["objects"], weight_of=lambda o: o % 2)
        assert np.array_equal(I, I1) and np.array_equal(S, S1)
    assert len({tuple(I[:, 0].tolist()) + tuple(S[:, 0].tolist()) for S, I in stream}) > 1, "frames should differ"


@pytest.mark.parametrize("precision,rtol", [("fp32", 1e-4), ("bf16", 1e-2)])
def test_frame_scorer_batched_objects_two_scorers(ctx, precision, rtol):
    """3 objects, 2 scorers keyed on object parity (online_learning.py:461-463), pre-filter on: per-object
    top-k vs the oracle run object by object."""
    sc = syn.make_scene(23, "lmo", n_obj=3, n_pts=300, n_hypo=500)
    import cv2
    img01 = cv2.GaussianBlur(sc["img"], (5, 5), 0) / 255.
    meta = glue.K2meta(sc["cam_K"])
    ws = [weights.seeded_folded(0), weights.seeded_folded(1)]
    fs = scoring.FrameScorer(ws, device=0, precision=precision, inconst_ratio_th=10.0, k=5, chunk=170)
    S, I = fs.score_frame(sc["img"], sc["depth"], sc["cam_K"], sc["objects"], weight_of=lambda o: o % 2)
    for o, ob in enumerate(sc["objects"]):
        f = zo.features(img01, sc["depth"], ob["pose_hypos"], meta, ob["model_points"], ob["model_colors"], ob["model_normals"])
        keep = zo.violation_filter(f["viol"], 300, 10.0)
        ref = zo.scorer(f["point_x"][keep], ws[o % 2])
        es, ei = zo.topk(ref, 5)
        tol = rtol * float(ref.abs().max()) + 1e-6
        np.testing.assert_allclose(S[o, :len(es)], es.numpy(), rtol=0, atol=tol)
        assert int(I[o, 0]) == int(keep[ei[0]]), f"object {o}: top-1 differs"     # fp32 scorer / fp32 re-rank: unconditional
        assert set(I[o][I[o] >= 0].tolist()) <= set(keep.tolist())


@pytest.mark.parametrize("dtype,rtol", [(torch.float32, 1e-4), (torch.bfloat16, 2 ** -8)])
@pytest.mark.parametrize("intr,n_pts,n_hypo", [("lmo", 1000, 400), ("hd", 4000, 60), ("tiny", 77, 50)])
def test_features_without_side_outputs_match_oracle(ctx, dtype, rtol, intr, n_pts, n_hypo):
    """The variant that runs when no mask/uv/violation output is requested (FrameScorer's hot path)."""
    from ossid_code_b200.engine import poses_to_rt12
    sc = syn.make_scene(41, intr, n_obj=1, n_pts=n_pts, n_hypo=n_hypo)
    ob = sc["objects"][0]
    import cv2
    img01 = cv2.GaussianBlur(sc["img"], (5, 5), 0) / 255.
    meta = glue.K2meta(sc["cam_K"])
    ref = zo.features(img01, sc["depth"], ob["pose_hypos"], meta, ob["model_points"], ob["model_colors"], ob["model_normals"])
    ctx.set_frame(img01, sc["depth"], meta)
    ctx.set_object(0, ob["model_points"], ob["model_colors"], ob["model_normals"])
    p12 = poses_to_rt12(ob["pose_hypos"], ctx.device)
    hot, _, _, _ = ctx.features(0, p12, dtype=dtype)
    aux, uv, mask, _ = ctx.features(0, p12, dtype=dtype, want_uv=True, want_mask=True)
    assert torch.equal(uv.cpu(), ref["uv"]) and torch.equal(mask.cpu(), ref["mask"])
    got = hot.float().cpu()
    err = (got - ref["point_x"]).abs()
    assert not bool((err > 1e-6 + rtol * ref["point_x"].abs()).any()), f"worst {float(err.max()):.3e}"
    # zero pattern (invalid projections) must be identical to the exact variant
    assert torch.equal(hot.float().abs().sum(-1) == 0, aux.float().abs().sum(-1) == 0)


def _emulate_ranks(fs, sc, world, wof):
    """Run `world` emulated ranks on one device and merge their candidate records with zs_merge_topk (+ re-rank)."""
    recs = []
    for r in range(world):
        fs.forced_rank_world = (r, world)
        fs.upload(sc["img"], sc["depth"], sc["cam_K"], sc["objects"], wof)
        recs.append(fs.run_resident(local_record=True))
    S, I, P = fs.merge_records(torch.stack(recs))
    if fs.rerank:
        S, I = fs._rerank(S, I, P)
    S, I = S.cpu().numpy().copy(), I.cpu().numpy().copy()
    fs.forced_rank_world = None
    return S, I, recs


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("rerank", [False, True])
def test_sharded_scoring_top1_independent_of_gpu_count(ctx, world, rerank):
    """Emulates `world` ranks on one device: each scores its contiguous hypothesis slice, the candidate records are
    merged exactly as after the all-gather; the result must equal the unsharded run (same indices, same scores)."""
    sc = syn.make_scene(29, "lmo", n_obj=3, n_pts=200, n_hypo=257)
    sc["objects"][1]["pose_hypos"][100] = sc["objects"][1]["pose_hypos"][7]      # exact score tie across shards
    ws = [weights.seeded_folded(0), weights.seeded_folded(1)]
    wof = lambda o: o % 2
    fs = scoring.FrameScorer(ws, device=0, precision="bf16", inconst_ratio_th=10.0, k=6, rerank=rerank)
    S1, I1 = fs.score_frame(sc["img"], sc["depth"], sc["cam_K"], sc["objects"], weight_of=wof)
    Sm, Im, recs = _emulate_ranks(fs, sc, world, wof)
    assert np.array_equal(Im, I1), "sharded top-k indices differ from the single-GPU result"
    assert np.array_equal(Sm, S1)
    S_ref, I_ref = scoring.merge_records_reference(torch.stack(recs), 3, 6)       # kernel merge == torch restatement
    if not rerank:
        assert np.array_equal(I_ref.numpy(), Im) and np.array_equal(S_ref.numpy(), Sm)
    tie = [int(x) for x in I1[1] if x in (7, 100)]
    if len(tie) == 2:
        assert tie == [7, 100], "tie must resolve to the lower hypothesis index"


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_prefilter_never_empty_rule_is_global(ctx, world):
    """one shard's slice is rejected entirely by the free-space pre-filter (its never-empty fallback must
    not reach the merge), and one object is rejected on EVERY shard (exactly the single-GPU fallback must survive)."""
    sc = syn.make_scene(31, "lmo", n_obj=3, n_pts=200, n_hypo=240)
    far = np.tile(np.eye(4), (240, 1, 1)); far[:, 2, 3] = 0.05                  # in front of everything: all points violate
    per = -(-240 // world)
    sc["objects"][0]["pose_hypos"][per:2 * per] = far[:per]                      # object 0: rank 1's slice is all rejected
    sc["objects"][2]["pose_hypos"] = far.copy()                                  # object 2: rejected everywhere
    sc["objects"][2]["pose_hypos"][:, 0, 3] = np.linspace(-0.02, 0.02, 240)      # (different poses, so different counts)
    ws = [weights.seeded_folded(0), weights.seeded_folded(1)]
    wof = lambda o: o % 2
    for rerank in (False, True):
        fs = scoring.FrameScorer(ws, device=0, precision="bf16", inconst_ratio_th=10.0, k=5, rerank=rerank)
        S1, I1 = fs.score_frame(sc["img"], sc["depth"], sc["cam_K"], sc["objects"], weight_of=wof)
        Sm, Im, recs = _emulate_ranks(fs, sc, world, wof)
        assert np.array_equal(Im, I1) and np.array_equal(Sm, S1), (I1, Im)
        assert not any(per <= int(i) < 2 * per for i in I1[0]), "a rejected hypothesis reached object 0's top-k"
        assert (I1[2] >= 0).sum() == 1, "never-empty: exactly one hypothesis of a fully rejected object survives"
        info = torch.stack(recs)[:, 2 * 3 * 5: 2 * 3 * 5 + 6].reshape(world, 3, 2).cpu()
        assert int(info[1, 0, 0]) == 0 and int(info[0, 0, 0]) > 0 and int(info[:, 2, 0].sum()) == 0


def test_many_model_instances_do_not_share_stale_weights(ctx):
    """More PointNet2SSG instances than weight slots: each forward must use its own weights."""
    from ossid_code_b200 import zephyr_shim
    x = (torch.randn(6, 64, 8, generator=torch.Generator().manual_seed(0)) * 0.5).to(ctx.device)
    models, refs = [], []
    for seed in range(6):
        m = zephyr_shim.PointNet2SSG(8, None, 1)
        m.load_state_dict(weights.seeded_state_dict(seed))
        models.append(m.to(0).eval())
        refs.append(zo.scorer(x.cpu(), weights.seeded_folded(seed)))
    for rnd in range(2):
        for m, ref in zip(models, refs):
            got = m({"point_x": x}).reshape(-1).cpu()
            assert float((got - ref).abs().max()) <= 1e-4 * float(ref.abs().max()) + 1e-6


def test_uv_original_proxy_behaves_like_the_array_the_reference_reads(ctx):
    """to_np(uv_original)[pred_idx] (online_learning.py:474-478) through the lazy device-side proxy == the full array."""
    from ossid_code_b200 import zephyr_shim

    class Args:
        inconst_ratio_th = 100
    sc = syn.make_scene(17, "lmo", n_obj=1, n_pts=200, n_hypo=64)
    ob = sc["objects"][0]
    data = dict(img=sc["img"], depth=sc["depth"], cam_K=sc["cam_K"], model_points=ob["model_points"],
                model_colors=ob["model_colors"], model_normals=ob["model_normals"], pose_hypos=ob["pose_hypos"].copy())
    ds = zephyr_shim.ScoreDataset([], "", "lmo", Args(), mode="test")
    model = zephyr_shim.PointNet2SSG(8, Args(), num_class=1).to(0).eval()
    poses, scores, err, uv = glue.networkInference(model, ds, data)
    full = np.asarray(glue.to_np(uv))
    assert isinstance(uv, zephyr_shim.UvOriginal) and full.dtype == np.int64 and full.shape == (64, 200, 2) == uv.shape
    i = int(np.argmax(scores))
    assert np.array_equal(glue.to_np(uv)[i], full[i]) and np.array_equal(uv[np.int64(i)], full[i])
    assert np.array_equal(uv[torch.tensor([3, 1])], full[[3, 1]]) and torch.equal(uv.to("cpu"), torch.from_numpy(full))
    Args.zs_lazy_uv = False
    ds2 = zephyr_shim.ScoreDataset([], "", "lmo", Args(), mode="test")
    _, _, _, uv2 = glue.networkInference(model, ds2, data)
    assert torch.is_tensor(uv2) and uv2.dtype == torch.int64 and np.array_equal(uv2.cpu().numpy(), full)


def test_filtered_frame_with_device_side_counts_equals_the_synchronous_sequence(ctx):
    """Pre-filter on (YCB-V setting, online_learning.py:184), tensor-core path: FrameScorer keeps the kept counts on the
    device (zs_set_dynamic_count); the result must equal filter -> read count -> features -> pool -> head -> top-k run
    call by call, bit for bit, including an object whose hypotheses are all rejected but one (never-empty rule)."""
    sc = syn.make_scene(37, "lmo", n_obj=3, n_pts=300, n_hypo=700)
    far = np.tile(np.eye(4), (50, 1, 1)); far[:, 2, 3] = 0.05                  # object 2: everything violates free space
    sc["objects"][2]["pose_hypos"] = far
    ws = [weights.seeded_folded(0), weights.seeded_folded(1)]
    fs = scoring.FrameScorer(ws, device=0, precision="bf16", inconst_ratio_th=10.0, k=6, chunk=256, rerank=False)
    S, I = fs.score_frame(sc["img"], sc["depth"], sc["cam_K"], sc["objects"], weight_of=lambda o: o % 2)
    scored = fs.last_scored
    n_kept = 0
    for o, r in enumerate(fs._resident):
        keep = ctx.filter(ctx.violations(r["slot"], r["poses12"]), ctx.obj_npts[r["slot"]], 10.0)
        n_kept += keep.shape[0]
        feat, _, _, _ = ctx.features(r["slot"], r["poses12"], keep_idx=keep, dtype=torch.bfloat16)
        scores = ctx.head(r["wslot"], ctx.pool(r["wslot"], feat), tensor_cores=True)
        es, ei = ctx.topk(scores, 6, r["lo"], index_map=keep)
        assert np.array_equal(S[o], es.cpu().numpy()) and np.array_equal(I[o], ei.cpu().numpy()), (o, S[o], es, I[o], ei)
    assert scored == n_kept and 0 < n_kept < 1450


def test_kernel_merge_of_gathered_candidates_equals_torch_merge(ctx):
    """The one-launch merge after the all-gather == its torch restatement: ties across ranks, empty slots, a NaN,
    k > valid count, a genuine -inf score (keeps its index), poses travelling with the winners."""
    g = torch.Generator().manual_seed(11)
    world, k, n_obj = 8, 6, 5
    per = 1000
    recs = []
    at = scoring.record_pose_offset(n_obj, k)
    for r in range(world):                                   # every rank: k candidates ordered by (score desc, index asc)
        s = torch.randn(n_obj, per, generator=g)
        s[1, 7] = 5.0                                          # the same top score on every rank of object 1
        if r == 3:
            s[2, :] = float("nan")                             # a rank that only has NaNs for object 2
        if r == 5:
            s[3, :] = float("-inf")                            # genuine -inf scores: valid hypotheses, lowest rank
        top = torch.stack([torch.tensor(sorted(range(per), key=lambda j: (-float(torch.nan_to_num(s[o, j], nan=-1e30, neginf=-1e29)), j))[:k])
                           for o in range(n_obj)])
        gs, gi = torch.gather(s, 1, top), (top + r * per).to(torch.int32)
        if r > 0:
            gi[4, :] = -1                                      # object 4: only rank 0 has candidates ...
        else:
            gi[4, 3:] = -1                                     # ... and only three of them
        poses = (gi.to(torch.float32)[..., None] + torch.arange(12) / 16.0)       # recognisable pose rows
        rec = torch.zeros(scoring.record_ints(n_obj, k, True), dtype=torch.int32)
        rec[: n_obj * k] = gs.contiguous().view(torch.int32).reshape(-1)
        rec[n_obj * k: 2 * n_obj * k] = gi.reshape(-1)
        rec[2 * n_obj * k: 2 * n_obj * k + 2 * n_obj] = torch.tensor([[1, 0]] * n_obj, dtype=torch.int32).reshape(-1)
        rec[at:] = poses.contiguous().view(torch.int32).reshape(-1)
        recs.append(rec)
    gathered = torch.stack(recs)
    S0, I0 = scoring.merge_records_reference(gathered, n_obj, k)
    P1 = torch.zeros(n_obj, k, 12, device=ctx.device)
    S1, I1 = ctx.merge_topk(gathered.to(ctx.device), n_obj, k, poses_out=P1)
    assert torch.equal(I0, I1.cpu()), (I0, I1)
    assert torch.equal(S0, S1.cpu())
    assert I1[1, :3].tolist() == [7, 1007, 2007] and I1[4, 3:].tolist() == [-1, -1, -1]
    exp_p = torch.where(I1.cpu()[..., None] >= 0, I1.cpu().to(torch.float32)[..., None] + torch.arange(12) / 16.0, torch.zeros(1))
    assert torch.equal(P1.cpu(), exp_p)
    S2, I2 = ctx.merge_topk(gathered.to(ctx.device), n_obj, k)                  # without the pose output
    assert torch.equal(I2, I1) and torch.equal(S2, S1)


def test_pack_poses_equals_host_cast(ctx):
    """zs_pack_poses: (n,4,4) float64 / float32 -> (n,12) float32, the single cast of the hand-over, bit for bit."""
    from ossid_code_b200.engine import poses_to_rt12
    g = torch.Generator().manual_seed(3)
    T = torch.randn(1001, 4, 4, generator=g, dtype=torch.float64) * 3.0
    for src in (T, T.to(torch.float32)):
        got = poses_to_rt12(src, ctx.device).cpu()
        assert torch.equal(got, src[:, :3, :4].to(torch.float32).reshape(-1, 12))
    assert poses_to_rt12(T[:0], ctx.device).shape == (0, 12)


def test_frame_scorer_and_shim_models_do_not_clobber_each_others_weights(ctx):
    """FrameScorer and PointNet2SSG share the per-device context's weight slots; each must re-upload when
    the other has taken its slot, in both directions."""
    from ossid_code_b200 import zephyr_shim
    sc = syn.make_scene(19, "lmo", n_obj=2, n_pts=128, n_hypo=60)
    ws = [weights.seeded_folded(0), weights.seeded_folded(1)]
    fs = scoring.FrameScorer(ws, device=0, precision="bf16", k=4)
    S0, I0 = fs.score_frame(sc["img"], sc["depth"], sc["cam_K"], sc["objects"], weight_of=lambda o: o % 2)
    x = (torch.randn(5, 64, 8, generator=torch.Generator().manual_seed(0)) * 0.5).to(ctx.device)
    models = []
    for seed in (7, 8, 9, 10):                                    # four models: every slot of the context gets taken
        m = zephyr_shim.PointNet2SSG(8, None, 1)
        m.load_state_dict(weights.seeded_state_dict(seed))
        models.append((m.to(0).eval(), zo.scorer(x.cpu(), weights.seeded_folded(seed))))
    for m, ref in models:
        got = m({"point_x": x}).reshape(-1).cpu()
        assert float((got - ref).abs().max()) <= 1e-4 * float(ref.abs().max()) + 1e-6
    S1, I1 = fs.score_frame(sc["img"], sc["depth"], sc["cam_K"], sc["objects"], weight_of=lambda o: o % 2)
    assert np.array_equal(S0, S1) and np.array_equal(I0, I1), "FrameScorer scored with a shim model's weights"
    for m, ref in models:                                         # and the models re-upload after the FrameScorer ran
        got = m({"point_x": x}).reshape(-1).cpu()
        assert float((got - ref).abs().max()) <= 1e-4 * float(ref.abs().max()) + 1e-6


@pytest.mark.parametrize("fmt", ["bf16", "f32", "split"])
def test_features_multi_equals_per_object_launches(ctx, fmt):
    """One launch over several objects (different cloud sizes, a kept-index list, a device-side count, an empty
    segment, 40 segments = two launches) == zs_features object by object, bit for bit."""
    from ossid_code_b200.engine import poses_to_rt12
    sc = syn.make_scene(47, "lmo", n_obj=4, n_pts=300, n_hypo=90)
    for key in ("model_points", "model_colors", "model_normals"):
        sc["objects"][1][key] = sc["objects"][1][key][:77].copy()
        sc["objects"][2][key] = np.concatenate([sc["objects"][2][key]] * 5)[:1400].copy()
    ctx.set_frame_u8(sc["img"], sc["depth"], glue.K2meta(sc["cam_K"]))
    p12 = []
    for o, ob in enumerate(sc["objects"]):
        ctx.set_object(o, ob["model_points"], ob["model_colors"], ob["model_normals"])
        p12.append(poses_to_rt12(ob["pose_hypos"], ctx.device))

    def empty(n, N):
        if fmt == "split":
            return torch.zeros((n, 2, N, 8), dtype=torch.bfloat16, device=ctx.device)
        return torch.zeros((n, N, 8), dtype=torch.float32 if fmt == "f32" else torch.bfloat16, device=ctx.device)

    keep = torch.arange(1, 90, 4, dtype=torch.int32, device=ctx.device)
    n_dev = torch.tensor([13], dtype=torch.int32, device=ctx.device)
    outs = [empty(90, 300), empty(len(keep), 77), empty(90, 1400), empty(20, 300), empty(0, 300)]
    ctx.features_multi([(0, p12[0], outs[0]), (1, p12[1], outs[1], keep), (2, p12[2], outs[2]),
                        (3, p12[3][5:], outs[3], None, n_dev, 5), (0, p12[0][:0], outs[4])])
    ref0 = ctx.features(0, p12[0], out=empty(90, 300))[0]
    ref1 = ctx.features(1, p12[1], keep_idx=keep, out=empty(len(keep), 77))[0]
    ref2 = ctx.features(2, p12[2], out=empty(90, 1400))[0]
    ref3 = empty(20, 300)
    ctx.features(3, p12[3][5:13].contiguous(), out=ref3[:8])           # entries [5, 13) of a list of 13; rows beyond stay untouched
    for i, (got, ref) in enumerate(zip(outs, (ref0, ref1, ref2, ref3))):
        assert torch.equal(got, ref), f"segment {i}"
    many = [empty(3, 300) for _ in range(40)]
    ctx.features_multi([(0, p12[0][i: i + 3], many[i]) for i in range(40)])
    for i in range(40):
        assert torch.equal(many[i], ref0[i: i + 3]), f"segment {i} of 40"


@pytest.mark.parametrize("n_pts,n_obj,n_hypo", [(1000, 3, 200), (128, 2, 37), (300, 5, 150), (129, 40, 9), (4000, 1, 20)])
def test_fused_features_and_mlp_equal_the_two_kernel_sequence(ctx, n_pts, n_obj, n_hypo):
    """zs_pool_fused (producer warps of the tensor-core MLP kernel featurise each tile themselves) == zs_features(bf16)
    followed by zs_pool, bit for bit: several objects per launch, tail tiles shifted back, more than 32 segments."""
    from ossid_code_b200.engine import poses_to_rt12
    sc = syn.make_scene(53, "lmo", n_obj=min(n_obj, 5), n_pts=n_pts, n_hypo=n_hypo)
    ctx.set_frame_u8(sc["img"], sc["depth"], glue.K2meta(sc["cam_K"]))
    ctx.set_weights(1, weights.seeded_folded(3))
    segs, ref = [], []
    for s in range(n_obj):
        ob = sc["objects"][s % len(sc["objects"])]
        ctx.set_object(s, ob["model_points"], ob["model_colors"], ob["model_normals"])
        p12 = poses_to_rt12(np.roll(ob["pose_hypos"], s, axis=0), ctx.device)
        segs.append((s, p12))
        feat, _, _, _ = ctx.features(s, p12, dtype=torch.bfloat16)
        ref.append(ctx.pool(1, feat))
    ref = torch.cat(ref)
    out = torch.full((n_obj * n_hypo + 3, 1024), -7.0, device=ctx.device)
    ctx.pool_fused(1, segs, out=out)
    assert torch.equal(out[: n_obj * n_hypo], ref)
    assert bool((out[n_obj * n_hypo:] == -7.0).all()), "rows beyond the hypothesis list were written"


def test_fused_kernel_with_kept_lists_and_device_counts(ctx):
    """zs_pool_fused with a pre-filter's outputs: per segment a kept-index list and a device-side count (one segment
    empty, one full, one partial); rows are laid out by capacity and rows beyond a count stay untouched."""
    from ossid_code_b200.engine import poses_to_rt12
    sc = syn.make_scene(67, "lmo", n_obj=3, n_pts=384, n_hypo=120)
    ctx.set_frame_u8(sc["img"], sc["depth"], glue.K2meta(sc["cam_K"]))
    ctx.set_weights(1, weights.seeded_folded(3))
    g = torch.Generator().manual_seed(1)
    segs, ref_rows = [], []
    out = torch.full((3 * 120, 1024), -7.0, device=ctx.device)
    for s, (ob, live) in enumerate(zip(sc["objects"], (37, 0, 120))):
        ctx.set_object(s, ob["model_points"], ob["model_colors"], ob["model_normals"])
        p12 = poses_to_rt12(ob["pose_hypos"], ctx.device)
        keep = torch.randperm(120, generator=g).to(torch.int32).to(ctx.device)
        n_dev = torch.tensor([live], dtype=torch.int32, device=ctx.device)
        segs.append((s, p12, keep, n_dev))
        feat, _, _, _ = ctx.features(s, p12, keep_idx=keep[:live].contiguous(), dtype=torch.bfloat16) if live else (None,) * 4
        ref_rows.append(ctx.pool(1, feat) if live else None)
    ctx.pool_fused(1, segs, out=out)
    for s, (ref, live) in enumerate(zip(ref_rows, (37, 0, 120))):
        rows = out[s * 120: (s + 1) * 120]
        if live:
            assert torch.equal(rows[:live], ref), f"segment {s}"
        assert bool((rows[live:] == -7.0).all()), f"segment {s}: rows beyond the device-side count were written"


@pytest.mark.parametrize("th,boxes", [(100.0, False), (10.0, False), (100.0, True), (10.0, True)])
def test_frame_scorer_fused_equals_unfused(ctx, th, boxes):
    """FrameScorer with the fused kernel == the two-kernel sequence, bit for bit: unfiltered, free-space pre-filter,
    detection boxes, both."""
    sc = syn.make_scene(59, "ycbv", n_obj=5, n_pts=1000, n_hypo=400)
    if boxes:
        for ob in sc["objects"]:
            ob["boxes"], ob["box_scores"] = np.asarray([syn.gt_box(sc, ob, 1.0)], dtype=np.float64), np.asarray([0.9])
    ws = [weights.seeded_folded(0), weights.seeded_folded(1)]
    res = []
    for fused in (True, False):
        fs = scoring.FrameScorer(ws, device=0, precision="bf16", k=8, fused=fused, rerank=False, inconst_ratio_th=th)
        res.append(fs.score_frame(sc["img"], sc["depth"], sc["cam_K"], sc["objects"], weight_of=lambda o: o % 2) + (fs.last_scored,))
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1]) and res[0][2] == res[1][2]
    assert (th >= 100 and not boxes) == (res[0][2] == 2000)


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_frame_scorer_with_detection_boxes_equals_oracle_pipeline(ctx, precision):
    """LM-O-style frame (BASELINE.json configs[2]): every object comes with DTOID boxes; the boxes are rasterised on the
    device, the mask-overlap test runs inside the pre-filter pass, and the scored set / top-1 equal the oracle pipeline
    boxes_to_mask -> filterHypoByMask -> features -> scorer -> argmax; also sharded over 3 emulated ranks."""
    import cv2
    sc = syn.make_scene(61, "lmo", n_obj=3, n_pts=256, n_hypo=300)
    ws = [weights.seeded_folded(0), weights.seeded_folded(1)]
    wof = lambda o: o % 2
    meta = glue.K2meta(sc["cam_K"])
    img01 = cv2.GaussianBlur(sc["img"], (5, 5), 0) / 255.
    for o, ob in enumerate(sc["objects"]):
        x1, y1, x2, y2 = syn.gt_box(sc, ob, 1.0)
        ob["boxes"] = np.array([[x1, y1, x2, y2], [5.0, 5.0, 40.0, 40.0]], dtype=np.float64)
        ob["box_scores"] = np.array([0.8, 0.3])
    sc["objects"][2]["boxes"] = np.array([[600.0, 440.0, 630.0, 470.0]])            # a box no hypothesis projects into
    sc["objects"][2]["box_scores"] = np.array([0.9])
    fs = scoring.FrameScorer(ws, device=0, precision=precision, inconst_ratio_th=100.0, k=8)
    S, I = fs.score_frame(sc["img"], sc["depth"], sc["cam_K"], sc["objects"], weight_of=wof)
    n_scored = 0
    for o, ob in enumerate(sc["objects"]):
        mask = zo.boxes_to_mask(sc["depth"], ob["boxes"], ob["box_scores"])
        kept = zo.mask_filter(zo.project_raw(ob["pose_hypos"], ob["model_points"], meta), torch.from_numpy(mask.astype(np.int64)),
                              256, th=0.5).numpy()
        idx = np.nonzero(kept)[0]
        n_scored += len(idx)
        got = [int(i) for i in I[o] if i >= 0]
        assert set(got) <= set(idx.tolist()), f"object {o}: a hypothesis outside the detection mask was scored"
        if len(idx) == 0:
            assert got == []
            continue
        f = zo.features(img01, sc["depth"], ob["pose_hypos"][idx], meta, ob["model_points"], ob["model_colors"], ob["model_normals"])
        ref = zo.scorer(f["point_x"], ws[o % 2])
        assert got[0] == int(idx[int(torch.argmax(ref))]), f"object {o}: top-1 differs from the oracle pipeline"
        assert abs(float(S[o, 0]) - float(ref.max())) <= 1e-4 * float(ref.abs().max()) + 1e-6
    assert fs.last_scored == n_scored and (I[2] < 0).all() and 0 < n_scored < 900
    Sm, Im, _ = _emulate_ranks(fs, sc, 3, wof)
    assert np.array_equal(Im, I) and np.array_equal(Sm, S)


def test_prefilter_of_a_whole_frame_equals_per_object_calls(ctx):
    """zs_prefilter (two launches for all objects) == zs_violations + zs_filter object by object: different cloud sizes,
    one object with a detection mask, one whose hypotheses are all rejected, 40 objects (two projection launches)."""
    from ossid_code_b200.engine import poses_to_rt12
    sc = syn.make_scene(79, "lmo", n_obj=3, n_pts=500, n_hypo=333)
    for key in ("model_points", "model_colors", "model_normals"):
        sc["objects"][1][key] = sc["objects"][1][key][:130].copy()
    far = np.tile(np.eye(4), (333, 1, 1)); far[:, 2, 3] = 0.05
    sc["objects"][2]["pose_hypos"] = far
    ctx.set_frame_u8(sc["img"], sc["depth"], glue.K2meta(sc["cam_K"]))
    x1, y1, x2, y2 = syn.gt_box(sc, sc["objects"][0], 1.2)
    mask = torch.zeros((sc["H"], sc["W"]), dtype=torch.uint8, device=ctx.device)
    mask[y1:y2, x1:x2] = 1
    segs, refs = [], []
    for s in range(40):
        ob = sc["objects"][s % 3]
        ctx.set_object(s, ob["model_points"], ob["model_colors"], ob["model_normals"])
        p12 = poses_to_rt12(np.roll(ob["pose_hypos"], s, axis=0), ctx.device)
        m = mask if s % 3 == 0 else None
        viol = ctx.violations(s, p12, mask=m, mask_th=0.5)
        keep, n_keep = ctx.filter_async(viol, ctx.obj_npts[s], 10.0, info=(info := torch.zeros(2, dtype=torch.int32, device=ctx.device)))
        refs.append((viol, keep, n_keep, info))
        dev = lambda n: torch.full((n,), -3, dtype=torch.int32, device=ctx.device)
        segs.append((s, p12, m, dev(333), dev(333), dev(1), dev(2)))
    ctx.prefilter(segs, 10.0, 0.5)
    for s, (sg, (viol, keep, n_keep, info)) in enumerate(zip(segs, refs)):
        nk = int(n_keep)
        assert torch.equal(sg[3], viol) and int(sg[5]) == nk and torch.equal(sg[4][:nk], keep[:nk]) and torch.equal(sg[6], info), s
    assert int(segs[2][5]) == 1 and int(segs[2][6][0]) == 0, "all rejected: the never-empty rule keeps one"
    assert 0 < int(segs[0][5]) < 333 and bool((segs[0][3] == 0x7fffffff).any())


@pytest.mark.parametrize("th", [100.0, 10.0])
def test_cuda_graph_replay_equals_eager_launches(ctx, th):
    """After one eager run and one capture a single-GPU step is replayed from a CUDA graph; different frames of the same
    shape (new image, depth, poses), a frame of another shape in between, a bigger cloud (context buffers move) and a
    second user of the context's weight slots must all give exactly what the kernel-by-kernel scorer gives."""
    from ossid_code_b200 import zephyr_shim
    ws = [weights.seeded_folded(0), weights.seeded_folded(1)]
    wof = lambda o: o % 2
    fg = scoring.FrameScorer(ws, device=0, precision="bf16", k=6, inconst_ratio_th=th, graph=True)
    fe = scoring.FrameScorer(ws, device=0, precision="bf16", k=6, inconst_ratio_th=th, graph=False)
    frames = [syn.make_scene(seed, "lmo", n_obj=3, n_pts=256, n_hypo=200) for seed in (83, 84, 85, 86, 87)]
    other = syn.make_scene(88, "lmo", n_obj=2, n_pts=256, n_hypo=90)
    bigger = syn.make_scene(89, "lmo", n_obj=3, n_pts=700, n_hypo=200)
    seq = frames[:3] + [other] + frames[3:] + frames + [bigger, frames[0], bigger, bigger, frames[1], bigger, bigger]
    l0 = fg.ctx.launches
    for i, sc in enumerate(seq):
        Sg, Ig = fg.score_frame(sc["img"], sc["depth"], sc["cam_K"], sc["objects"], weight_of=wof)
        Se, Ie = fe.score_frame(sc["img"], sc["depth"], sc["cam_K"], sc["objects"], weight_of=wof)
        assert np.array_equal(Sg, Se) and np.array_equal(Ig, Ie), f"frame {i}"
        if i == 6:                                            # another user takes the context's weight slots
            m = zephyr_shim.PointNet2SSG(8, None, 1)
            m.load_state_dict(weights.seeded_state_dict(9))
            m.to(0).eval()({"point_x": torch.zeros(2, 64, 8, device=ctx.device)})
    assert fg.use_graph and fg.ctx.graph_launches > 0 and fg.ctx.launches > l0, "the step was never replayed from a graph"
    out = fg.score_frames([dict(img=sc["img"], depth=sc["depth"], cam_K=sc["cam_K"], objects=sc["objects"]) for sc in frames], wof)
    for sc, (S, I) in zip(frames, out):
        Se, Ie = fe.score_frame(sc["img"], sc["depth"], sc["cam_K"], sc["objects"], weight_of=wof)
        assert np.array_equal(S, Se) and np.array_equal(I, Ie)


def test_more_objects_than_cloud_slots_raises(ctx):
    sc = syn.make_scene(19, "tiny", n_obj=1, n_pts=32, n_hypo=4)
    fs = scoring.FrameScorer([weights.seeded_folded(0)], device=0, k=2)
    with pytest.raises(ValueError, match="at most 64 objects"):
        fs.upload(sc["img"], sc["depth"], sc["cam_K"], [sc["objects"][0]] * 65)
    # the host-buffer API scores such a frame in groups of 64 instead
    objs = [dict(sc["objects"][0], pose_hypos=np.roll(sc["objects"][0]["pose_hypos"], o, axis=0)) for o in range(70)]
    S, I = fs.score_frame(sc["img"], sc["depth"], sc["cam_K"], objs)
    S0, I0 = fs.score_frame(sc["img"], sc["depth"], sc["cam_K"], objs[:1])
    assert S.shape == (70, 2) and np.array_equal(S[0], S0[0]) and np.array_equal(I[0], I0[0])
    assert all(np.array_equal(S[o], S[0]) and np.array_equal((I[o] - o) % 4, I[0]) for o in range(70))


@pytest.mark.parametrize("chunk", [96, 32768])
def test_frame_scorer_paths_agree_batched_per_scorer_vs_per_object(ctx, chunk):
    """No pre-filter: clouds of one size take the "one MLP launch per chunk of a scorer's objects" path, clouds of
    different sizes the per-object path; both must equal scoring each object on its own (bit for bit), also when
    chunk boundaries fall inside objects."""
    sc = syn.make_scene(29, "lmo", n_obj=4, n_pts=256, n_hypo=150)
    ws = [weights.seeded_folded(0), weights.seeded_folded(1)]
    wof = lambda o: o % 2
    fs = scoring.FrameScorer(ws, device=0, precision="bf16", inconst_ratio_th=100.0, k=5, chunk=chunk)

    def alone(objects):
        out = []
        for o, ob in enumerate(objects):
            one = scoring.FrameScorer(ws, device=0, precision="bf16", inconst_ratio_th=100.0, k=5)
            S, I = one.score_frame(sc["img"], sc["depth"], sc["cam_K"], [dict(ob)], weight_of=lambda _: wof(o))
            out.append((S[0], I[0]))
        return out

    S, I = fs.score_frame(sc["img"], sc["depth"], sc["cam_K"], sc["objects"], weight_of=wof)      # same sizes: batched
    for o, (s1, i1) in enumerate(alone(sc["objects"])):
        assert np.array_equal(S[o], s1) and np.array_equal(I[o], i1), o
    mixed = [dict(ob) for ob in sc["objects"]]
    for key in ("model_points", "model_colors", "model_normals"):
        mixed[1][key] = mixed[1][key][:130].copy()                                                  # a smaller cloud
    for ob in mixed:
        ob.pop("_zs_host", None)
    S2, I2 = fs.score_frame(sc["img"], sc["depth"], sc["cam_K"], mixed, weight_of=wof)              # per-object path
    for o, (s1, i1) in enumerate(alone(mixed)):
        assert np.array_equal(S2[o], s1) and np.array_equal(I2[o], i1), o
    assert np.array_equal(S2[0], S[0]) and not np.array_equal(S2[1], S[1])
