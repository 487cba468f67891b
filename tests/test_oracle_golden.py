"""The oracle against fixtures produced by the reference's own code (oracle/gen_golden.py)."""
import colorsys
import os

import numpy as np
import pytest
import torch

from oracle import zephyr_oracle as zo
from ossid_code_b200 import synthetic as syn, weights, zephyr_utils as glue


def _meta(g):
    return dict(camera_fx=float(g["fx"]), camera_fy=float(g["fy"]), camera_cx=float(g["cx"]),
                camera_cy=float(g["cy"]), camera_scale=1.0)


def _poses44(p):
    return torch.from_numpy(np.asarray(p, np.float64))


def test_projection_matches_projectModelPoint(golden_dir):
    """uv, bounds and front-facing selection bit-exact vs ycbv_sift_dataset.py:303-320 (run in f64)."""
    g = np.load(os.path.join(golden_dir, "projection.npz"))
    H, W = int(g["H"]), int(g["W"])
    f = zo.features(np.zeros((H, W, 3)), np.ones((H, W)), g["poses"], _meta(g), g["model_points"],
                    np.zeros_like(g["model_points"]), g["model_normals"])
    offs = g["ref_offsets"]
    n_checked = 0
    for i in range(len(g["poses"])):
        sel = ((f["mask"][i] & zo.BIT_VALID_PROJ) != 0) & ((f["mask"][i] & zo.BIT_FRONT) != 0)
        idx = torch.nonzero(sel).reshape(-1).numpy()
        ref_idx, ref_uv = g["ref_idx"][offs[i]:offs[i + 1]], g["ref_uv"][offs[i]:offs[i + 1]]
        assert np.array_equal(idx, ref_idx), f"pose {i}: visible point set differs"
        assert np.array_equal(f["uv"][i][sel].numpy(), ref_uv), f"pose {i}: pixel indices differ"
        n_checked += len(idx)
    assert n_checked > 1000
    assert int(f["mask"][-1].bitwise_and(zo.BIT_VALID_PROJ).sum()) == 0   # the off-frame pose


def test_normal_cosine_matches_kptProjGridCos(golden_dir):
    """ncos vs ycbv_object.py:72-74 on every valid projection; fp32 vs fp64, tolerance 1e-5 abs."""
    g = np.load(os.path.join(golden_dir, "projection.npz"))
    H, W = int(g["H"]), int(g["W"])
    f = zo.features(np.zeros((H, W, 3)), np.ones((H, W)), g["poses"], _meta(g), g["model_points"],
                    np.zeros_like(g["model_points"]), g["model_normals"])
    valid = (f["mask"] & zo.BIT_VALID_PROJ) != 0
    ours = f["point_x"][..., 6].numpy()
    ref = g["ref_cos"].T                                    # (poses, points)
    assert valid.sum() > 1000
    np.testing.assert_allclose(ours[valid.numpy()], ref[valid.numpy()], atol=1e-5, rtol=0)


def test_mask_filter_matches_filterHypoByMask(golden_dir):
    g = np.load(os.path.join(golden_dir, "mask_filter.npz"))
    K = g["cam_K"]
    meta = glue.K2meta(K)
    uv = zo.project_raw(g["pose_hypos"], g["model_points"], meta)
    for th, key in ((0.5, "kept_050"), (0.9, "kept_090"), (0.0, "kept_000")):
        kept = zo.mask_filter(uv, torch.from_numpy(g["mask"].astype(np.int64)), g["model_points"].shape[0], th)
        assert np.array_equal(kept.numpy(), g[key])
    assert 0 < g["kept_050"].sum() < len(g["kept_050"])


@pytest.mark.parametrize("tag,th", [("th100", 100.0), ("th10", 10.0)])
def test_glue_mirror_matches_reference_networkInference(golden_dir, tag, th):
    """Our networkInference mirror + oracle objects == the reference's networkInference + same objects."""
    g = np.load(os.path.join(golden_dir, "network_inference.npz"))
    data = dict(img=g["img"], depth=g["depth"], cam_K=g["cam_K"], model_colors=g["model_colors"],
                model_points=g["model_points"], model_normals=g["model_normals"],
                pose_hypos=g["pose_hypos"].copy(), pp_err=np.arange(len(g["pose_hypos"]), dtype=np.float64))
    ds, model = zo.OracleScoreDataset(th), zo.OracleScorer(weights.seeded_folded(int(g["weight_seed"])))
    poses, scores, errs, uv, dt = glue.networkInference(model, ds, data, return_time=True)
    assert np.array_equal(poses, g[f"{tag}_poses"])
    np.testing.assert_allclose(np.asarray(scores).reshape(-1), g[f"{tag}_scores"], rtol=1e-5, atol=1e-6)
    assert np.array_equal(np.asarray(errs), g[f"{tag}_pp_err"])
    assert np.array_equal(uv.numpy(), g[f"{tag}_uv"].astype(np.int64))
    assert dt > 0
    if th < 100:
        assert len(scores) < len(g["pose_hypos"])           # the pre-filter really dropped something


def test_boxes_to_mask_matches_the_reference_statements(golden_dir):
    """oracle.boxes_to_mask vs the mask produced by the reference's own box -> mask statements
    (python/ossid/scripts/online_learning.py:389-405, AST-extracted and executed by oracle/gen_golden.py)."""
    g = np.load(os.path.join(golden_dir, "boxes_mask.npz"))
    for tag in "abc":
        m = zo.boxes_to_mask(g["depth"], g[f"{tag}_boxes"], g[f"{tag}_scores"])
        assert np.array_equal(m.astype(np.uint8), g[f"{tag}_mask"]), tag
    assert g["a_mask"].sum() > 0 and not np.array_equal(g["a_mask"], g["b_mask"])


def test_rgb_to_hsv_matches_colorsys():
    rng = np.random.default_rng(0)
    rgb = np.concatenate([rng.uniform(0, 1, (500, 3)), [[0, 0, 0], [1, 1, 1], [.5, .5, .5], [1, 0, 0], [0, 1, 0],
                                                          [0, 0, 1], [1, 1, 0], [.2, .2, .7], [.7, .2, .2]]]).astype(np.float32)
    ours = zo.rgb_to_hsv(torch.from_numpy(rgb)).numpy()
    ref = np.array([colorsys.rgb_to_hsv(*map(float, c)) for c in rgb])
    np.testing.assert_allclose(ours, ref, atol=2e-6)


def test_features_scalar_recomputation():
    """Independent scalar re-derivation (numpy float32 scalars, one op at a time) of the oracle's features."""
    sc = syn.make_scene(3, "tiny", n_obj=1, n_pts=64, n_hypo=24)
    ob = sc["objects"][0]
    meta = glue.K2meta(sc["cam_K"])
    import cv2
    img = cv2.GaussianBlur(sc["img"], (5, 5), 0) / 255.
    f = zo.features(img, sc["depth"], ob["pose_hypos"], meta, ob["model_points"], ob["model_colors"], ob["model_normals"])
    f32 = np.float32
    fx, fy, cx, cy = (f32(meta[k]) for k in ("camera_fx", "camera_fy", "camera_cx", "camera_cy"))
    ifx, ify = f32(1) / fx, f32(1) / fy
    imgf, dep = img.astype(np.float32), sc["depth"].astype(np.float32)
    H, W = dep.shape
    T = ob["pose_hypos"].astype(np.float32)
    P, Nn, C = (ob[k].astype(np.float32) for k in ("model_points", "model_normals", "model_colors"))

    def hsv(c):
        return zo.rgb_to_hsv(torch.from_numpy(np.asarray(c, np.float32))).numpy()

    n_valid = 0
    with np.errstate(all="ignore"):
        for h in range(T.shape[0]):
            for p in range(0, P.shape[0], 3):
                R, t = T[h, :3, :3], T[h, :3, 3]
                xyz = [f32(f32(f32(R[i, 0] * P[p, 0]) + f32(R[i, 1] * P[p, 1])) + f32(R[i, 2] * P[p, 2])) + t[i] for i in range(3)]
                n3 = [f32(f32(R[i, 0] * Nn[p, 0]) + f32(R[i, 1] * Nn[p, 1])) + f32(R[i, 2] * Nn[p, 2]) for i in range(3)]
                x, y, z = (f32(v) for v in xyz)
                ur = np.rint(f32(f32(x / z) * fx) + cx)
                vr = np.rint(f32(f32(y / z) * fy) + cy)
                valid = bool(0 < z < np.inf and 0 <= ur < W and 0 <= vr < H)
                dot = -f32(f32(f32(x * n3[0]) + f32(y * n3[1])) + f32(z * n3[2]))
                mk = (zo.BIT_FRONT if dot > 0 else 0)
                exp = np.zeros(8, np.float32)
                u = v = 0
                if valid:
                    n_valid += 1
                    u, v = int(ur), int(vr)
                    d = dep[v, u]
                    vd = bool(0 < d < np.inf)
                    dD = f32(d - z) if vd else f32(0)
                    mk |= zo.BIT_VALID_PROJ | (zo.BIT_VALID_DEPTH if vd else 0)
                    mk |= zo.BIT_FREE_SPACE if (vd and dD > f32(0.02)) else 0
                    mk |= zo.BIT_OCCLUDED if (vd and dD < f32(-0.02)) else 0
                    ho, hm = hsv(imgf[v, u]), hsv(C[p])
                    dH = f32(ho[0] - hm[0])
                    dH = f32(dH - f32(1)) if dH > 0.5 else dH
                    dH = f32(dH + f32(1)) if dH < -0.5 else dH
                    nrm = np.sqrt(f32(f32(x * x + y * y) + z * z)) * np.sqrt(f32(f32(n3[0] ** 2 + n3[1] ** 2) + n3[2] ** 2))
                    nc = f32(dot / nrm)
                    exp[:] = [f32(f32(u) - cx) * ifx, f32(f32(v) - cy) * ify, dH, ho[1] - hm[1], ho[2] - hm[2], dD,
                              nc if np.isfinite(nc) else 0, 0]
                assert int(f["mask"][h, p]) == mk
                assert f["uv"][h, p].tolist() == [u, v]
                np.testing.assert_allclose(f["point_x"][h, p].numpy(), exp, rtol=1e-6, atol=1e-7)
    assert n_valid > 50


def test_violation_filter_rules():
    viol = torch.tensor([50, 5, 99, 0, 10, 100], dtype=torch.int32)
    assert zo.violation_filter(viol, 100, 100.0).tolist() == [0, 1, 2, 3, 4, 5]        # th >= 100: off
    assert zo.violation_filter(viol, 100, 10.0).tolist() == [1, 3]                       # strict <
    assert zo.violation_filter(viol, 100, 10.5).tolist() == [1, 3, 4]
    assert zo.violation_filter(torch.tensor([70, 60, 60, 90], dtype=torch.int32), 100, 10.0).tolist() == [1]  # never empty, first min
    assert zo.violation_filter(torch.zeros(0, dtype=torch.int32), 100, 10.0).tolist() == []


def test_topk_is_first_max_on_ties():
    s = torch.tensor([1.0, 3.0, 3.0, -2.0, 3.0, 0.5])
    ts, ti = zo.topk(s, 4)
    assert ti.tolist() == [1, 2, 4, 0] and int(np.argmax(s.numpy())) == ti[0]
    ts, ti = zo.topk(s, 3, index_base=100)
    assert ti.tolist() == [101, 102, 104]


def test_scorer_bf16_emulation_close_to_fp32():
    rng = torch.Generator().manual_seed(1)
    x = torch.randn(6, 50, 8, generator=rng) * 0.3
    w = weights.seeded_folded(0)
    a, b = zo.scorer(x, w), zo.scorer(x, w, bf16=True)
    assert a.shape == (6,) and torch.isfinite(a).all()
    assert float((a - b).abs().max()) <= 2e-2 * float(a.abs().max())
    assert float((a - b).abs().max()) > 0
