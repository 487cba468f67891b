"""Post-scoring refinement (ICP, visibility mask): oracle sanity on CPU, kernel vs oracle on the GPU."""
import numpy as np
import pytest
import torch

from oracle import icp_oracle as io
from ossid_code_b200 import synthetic as syn


def _scene(seed=11, n_pts=600):
    sc = syn.make_scene(seed, "lmo", n_obj=1, n_pts=n_pts, n_hypo=8)
    return sc, sc["objects"][0]


def _perturbed(gt, rng, rot_deg, trans):
    ax = rng.normal(size=3)
    ax /= np.linalg.norm(ax)
    a = np.deg2rad(rot_deg)
    Kx = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    R = np.eye(3) + np.sin(a) * Kx + (1 - np.cos(a)) * Kx @ Kx
    T = gt.copy()
    T[:3, :3] = R @ gt[:3, :3]
    T[:3, 3] = gt[:3, 3] + rng.normal(size=3) * trans
    return T


def _project(pts, T, K, H, W):
    p = pts @ T[:3, :3].T + T[:3, 3]
    u = np.rint(p[:, 0] / p[:, 2] * K[0, 0] + K[0, 2]).astype(np.int64)
    v = np.rint(p[:, 1] / p[:, 2] * K[1, 1] + K[1, 2]).astype(np.int64)
    bad = (u < 0) | (u >= W) | (v < 0) | (v >= H) | (p[:, 2] <= 0)
    u[bad] = 0
    v[bad] = 0
    return np.stack([u, v], 1)


def test_rigid_fit_recovers_a_known_transform():
    rng = np.random.default_rng(0)
    p = rng.normal(size=(50, 3))
    T = _perturbed(np.eye(4), rng, 25.0, 0.3)
    R, t = io.rigid_fit(p, p @ T[:3, :3].T + T[:3, 3])
    assert np.allclose(R, T[:3, :3], atol=1e-10) and np.allclose(t, T[:3, 3], atol=1e-10)
    assert np.isclose(np.linalg.det(R), 1.0)


def test_oracle_icp_converges_on_a_dense_target():
    """Target = the model surface under the ground-truth pose, source = the same model under a perturbed pose."""
    sc, ob = _scene(n_pts=2000)
    rng = np.random.default_rng(1)
    gt, pts = ob["gt_pose"], ob["model_points"]
    tgt = pts @ gt[:3, :3].T + gt[:3, 3]
    T0 = _perturbed(gt, rng, 2.0, 0.003)
    T1, info = io.icp_point_to_point(pts[::2], tgt, T0, max_dist=0.01)
    e0 = np.linalg.norm(pts @ T0[:3, :3].T + T0[:3, 3] - tgt, axis=1).mean()
    e1 = np.linalg.norm(pts @ T1[:3, :3].T + T1[:3, 3] - tgt, axis=1).mean()
    assert 1 <= info["iterations"] <= 30 and info["fitness"] > 0.9
    assert e1 < 0.25 * e0, (e0, e1, info)


def test_oracle_icp_from_depth_runs_and_reports():
    sc, ob = _scene()
    rng = np.random.default_rng(1)
    gt, pts = ob["gt_pose"], ob["model_points"]
    T0 = _perturbed(gt, rng, 2.0, 0.003)
    uv = _project(pts, T0, sc["cam_K"], sc["H"], sc["W"])
    T1, info = io.icp_refinement(sc["depth"], uv, T0, sc["cam_K"], pts)
    assert info["iterations"] >= 1 and info["n_corr"] > 50 and 0 < info["fitness"] <= 1
    assert np.isclose(np.linalg.det(T1[:3, :3]), 1.0, atol=1e-9)
    T2, info2 = io.icp_refinement(sc["depth"], uv, np.eye(4), sc["cam_K"], pts)     # nothing within 1 cm: pose unchanged
    assert info2["n_corr"] == 0 and np.array_equal(T2, np.eye(4))


def test_visibility_rule_matches_the_bop_definition():
    d_test = np.array([[0.0, 1.0, 1.0, 1.0, 0.5]], np.float32)
    d_model = np.array([[1.0, 1.01, 1.02, 0.0, 1.0]], np.float32)
    assert io.estimate_visib_mask(d_test, d_model, 0.015).tolist() == [[True, True, False, False, False]]
    assert io.estimate_visib_mask(d_test, d_model, 0.015, "bop18").tolist() == [[False, True, False, False, False]]


@pytest.mark.gpu
@pytest.mark.parametrize("n_pts,n_pose", [(600, 6), (1000, 40), (97, 3)])
def test_kernel_icp_matches_oracle(n_pts, n_pose):
    from ossid_code_b200 import icp
    sc, ob = _scene(seed=5 + n_pts, n_pts=n_pts)
    rng = np.random.default_rng(n_pts)
    gt, pts = ob["gt_pose"], ob["model_points"].astype(np.float32)
    poses = np.stack([_perturbed(gt, rng, rng.uniform(0.2, 3.0), 0.004) for _ in range(n_pose)])
    poses[-1] = np.eye(4)                                # identity placeholder (online_learning.py:431): nothing matches
    uvs = np.stack([_project(pts, T, sc["cam_K"], sc["H"], sc["W"]) for T in poses])
    T_gpu, st = icp.icp_refine_batch(sc["depth"], uvs, poses, sc["cam_K"], pts)
    worst = 0.0
    for h in range(n_pose):
        T_ref, info = io.icp_refinement(sc["depth"], uvs[h], poses[h].astype(np.float32).astype(np.float64), sc["cam_K"], pts)
        a = pts @ T_gpu[h][:3, :3].T + T_gpu[h][:3, 3]
        b = pts @ T_ref[:3, :3].T + T_ref[:3, 3]
        worst = max(worst, float(np.abs(a - b).max()))
        assert abs(st[h, 0] - info["fitness"]) <= 2.0 / n_pts + 1e-6, (h, st[h], info)
        assert abs(st[h, 1] - info["inlier_rmse"]) <= 2e-5, (h, st[h], info)
    print(f"n_pts={n_pts}: max point displacement between kernel and oracle refinements {worst:.2e} m")
    assert worst <= 2e-6, worst                          # fp32 correspondences vs fp64 kd-tree (measured <= 6e-8 m)
    assert st[-1, 3] == 0 and np.allclose(T_gpu[-1], np.eye(4))
    # the single-pose drop-in signature
    T1, info1 = icp.icpRefinement(sc["depth"], uvs[0], poses[0], sc["cam_K"], pts, inpaint_depth=False, icp_max_dist=0.01)
    assert np.allclose(T1, T_gpu[0]) and info1["iterations"] == int(st[0, 2])


@pytest.mark.gpu
def test_kernel_icp_on_resident_frame_equals_explicit_depth():
    from ossid_code_b200 import icp, zephyr_utils as glue
    from ossid_code_b200.engine import get_context
    sc, ob = _scene(seed=21, n_pts=500)
    ctx = get_context(0)
    ctx.set_frame_u8(sc["img"], sc["depth"], glue.K2meta(sc["cam_K"]))
    rng = np.random.default_rng(3)
    T0 = _perturbed(ob["gt_pose"], rng, 1.0, 0.003)
    uv = _project(ob["model_points"], T0, sc["cam_K"], sc["H"], sc["W"])
    a, sa = icp.icp_refine_batch(sc["depth"], uv, T0[None], sc["cam_K"], ob["model_points"])
    b, sb = icp.icp_refine_batch(None, uv, T0[None], sc["cam_K"], ob["model_points"])
    assert np.array_equal(a, b) and np.array_equal(sa, sb)


@pytest.mark.gpu
def test_kernel_visibility_mask_matches_oracle():
    from ossid_code_b200 import icp
    rng = np.random.default_rng(9)
    d_test = rng.uniform(0.3, 1.5, size=(480, 640)).astype(np.float32)
    d_test[rng.random(d_test.shape) < 0.1] = 0.0
    d_model = (d_test + rng.normal(scale=0.02, size=d_test.shape)).astype(np.float32)
    d_model[rng.random(d_test.shape) < 0.5] = 0.0
    for mode in ("bop19", "bop18"):
        assert np.array_equal(icp.estimate_visib_mask_gt(d_test, d_model, 15 / 1000., mode),
                              io.estimate_visib_mask(d_test, d_model, 15 / 1000., mode))


@pytest.mark.gpu
def test_frame_scorer_refines_its_winners_like_the_oracle():
    """score -> top-k -> ICP of the winners on the resident frame == oracle ICP of the same poses (uv from the oracle)."""
    import cv2
    from oracle import zephyr_oracle as zo
    from ossid_code_b200 import scoring, weights, zephyr_utils as glue
    sc = syn.make_scene(13, "lmo", n_obj=2, n_pts=400, n_hypo=200)
    fs = scoring.FrameScorer([weights.seeded_folded(0)], device=0, precision="fp32", k=3)
    S, I = fs.score_frame(sc["img"], sc["depth"], sc["cam_K"], sc["objects"])
    P, st = fs.refine_winners(sc["objects"], I, n_refine=2)
    img01 = cv2.GaussianBlur(sc["img"], (5, 5), 0) / 255.
    meta = glue.K2meta(sc["cam_K"])
    for o, ob in enumerate(sc["objects"]):
        for j in range(2):
            T0 = ob["pose_hypos"][int(I[o, j])].astype(np.float32).astype(np.float64)
            f = zo.features(img01, sc["depth"], T0[None], meta, ob["model_points"], ob["model_colors"], ob["model_normals"])
            T_ref, info = io.icp_refinement(sc["depth"], f["uv"][0].numpy(), T0, sc["cam_K"], ob["model_points"])
            a = ob["model_points"] @ P[o, j][:3, :3].T + P[o, j][:3, 3]
            b = ob["model_points"] @ T_ref[:3, :3].T + T_ref[:3, 3]
            assert np.abs(a - b).max() <= 2e-4, (o, j, np.abs(a - b).max(), st[o, j], info)
            assert abs(st[o, j, 0] - info["fitness"]) <= 2.0 / 400 + 1e-6
