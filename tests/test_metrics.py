"""ADD / ADI pose errors: oracle vs an independent KD-tree computation (CPU), kernel vs oracle (GPU)."""
import numpy as np
import pytest

from oracle import zephyr_oracle as zo
from ossid_code_b200 import synthetic as syn


def _case(seed=3, n_pts=300, n_hypo=40):
    sc = syn.make_scene(seed, "tiny", n_obj=1, n_pts=n_pts, n_hypo=n_hypo)
    ob = sc["objects"][0]
    return ob["pose_hypos"], ob["gt_pose"], ob["model_points"]


def test_oracle_add_adi_against_kdtree():
    from scipy.spatial import cKDTree
    P, G, pts = _case()
    gt = pts @ G[:3, :3].T + G[:3, 3]
    tree = cKDTree(gt)
    add_ref, adi_ref = [], []
    for T in P:
        est = pts @ T[:3, :3].T + T[:3, 3]
        add_ref.append(np.linalg.norm(est - gt, axis=1).mean())
        adi_ref.append(tree.query(est, k=1)[0].mean())
    np.testing.assert_allclose(zo.pose_errors(P, G, pts, False), add_ref, rtol=1e-12)
    np.testing.assert_allclose(zo.pose_errors(P, G, pts, True), adi_ref, rtol=1e-12)
    assert zo.pose_errors(G[None], G, pts, False)[0] == 0 and zo.pose_errors(G[None], G, pts, True)[0] == 0
    assert (zo.pose_errors(P, G, pts, True) <= zo.pose_errors(P, G, pts, False) + 1e-15).all()   # ADI <= ADD


@pytest.mark.gpu
@pytest.mark.parametrize("n_pts,n_hypo", [(300, 40), (1000, 300), (1, 3), (33, 1)])
def test_kernel_pose_errors_match_oracle(n_pts, n_hypo):
    from ossid_code_b200 import metrics
    P, G, pts = _case(5, n_pts, n_hypo)
    P = P[np.isfinite(P).all(axis=(1, 2))]
    for sym in (False, True):
        got = metrics.pose_errors(P, G, pts, symmetric=sym)
        ref = zo.pose_errors(P, G, pts, sym)
        np.testing.assert_allclose(got, ref, rtol=2e-5, atol=1e-6)      # fp32 kernel vs fp64 oracle
    R, t = P[0][:3, :3], P[0][:3, 3]
    assert abs(metrics.add(R, t, G[:3, :3], G[:3, 3], pts) - zo.pose_errors(P[:1], G, pts, False)[0]) < 1e-5
    assert abs(metrics.adi(R, t, G[:3, :3], G[:3, 3], pts) - zo.pose_errors(P[:1], G, pts, True)[0]) < 1e-5
    assert metrics.pose_errors(np.zeros((0, 4, 4)), G, pts).shape == (0,)
