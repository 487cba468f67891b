"""Seeded synthetic inputs for the Zephyr hypothesis-scoring path.

The reference ships no data, so tests and bench.py draw their
frames, model clouds and pose hypotheses from here.  Everything is a pure
function of an integer seed (numpy ``default_rng``); nothing touches the GPU.

Shapes follow the reference's hand-over at
``python/ossid/scripts/online_learning.py:455-459``:
``img`` uint8 (H,W,3), ``depth`` float32 metres (H,W), ``cam_K`` float64 (3,3),
``model_points/colors/normals`` float64 (N,3), ``pose_hypos`` float64 (M,4,4).
"""
from __future__ import annotations

import numpy as np

# Intrinsics used by BASELINE.json's configs.
INTRINSICS = {
    "lmo": (480, 640, 572.4114, 573.57043, 325.2611, 242.04899),
    "ycbv": (480, 640, 1066.778, 1067.487, 312.9869, 241.3109),
    "hd": (720, 1280, 920.0, 920.0, 640.0, 360.0),
    "tiny": (120, 160, 140.0, 141.0, 80.5, 59.25),     # small fixture frames (tests/golden)
}


def cam_K(name: str) -> np.ndarray:
    _, _, fx, fy, cx, cy = INTRINSICS[name]
    return np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], dtype=np.float64)


def rodrigues(rotvec: np.ndarray) -> np.ndarray:
    """(n,3) rotation vectors -> (n,3,3) rotation matrices."""
    rotvec = np.asarray(rotvec, dtype=np.float64).reshape(-1, 3)
    th = np.linalg.norm(rotvec, axis=1)
    k = rotvec / np.where(th > 0, th, 1.0)[:, None]
    K = np.zeros((len(th), 3, 3))
    K[:, 0, 1], K[:, 0, 2] = -k[:, 2], k[:, 1]
    K[:, 1, 0], K[:, 1, 2] = k[:, 2], -k[:, 0]
    K[:, 2, 0], K[:, 2, 1] = -k[:, 1], k[:, 0]
    s, c = np.sin(th)[:, None, None], np.cos(th)[:, None, None]
    return np.eye(3)[None] + s * K + (1 - c) * (K @ K)


def random_rotations(rng: np.random.Generator, n: int) -> np.ndarray:
    """Uniform rotations from normalised Gaussian quaternions."""
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q.T
    R = np.empty((n, 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - z * w); R[:, 0, 2] = 2 * (x * z + y * w)
    R[:, 1, 0] = 2 * (x * y + z * w); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - x * w)
    R[:, 2, 0] = 2 * (x * z - y * w); R[:, 2, 1] = 2 * (y * z + x * w); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def perturb_pose(rng: np.random.Generator, mat: np.ndarray, n: int) -> np.ndarray:
    """Pose perturbations with the reference's magnitudes.

    Same distribution as ``perturbTrans`` (python/ossid/utils/__init__.py:82-98):
    rotation-vector magnitude N(0, 0.2 rad) about a uniform axis, translation
    N(0, 0.01 m), rotation applied on the left.  Uses a local Generator instead
    of numpy's global state.
    """
    mag = rng.normal(0, 0.2, n)
    axis = rng.normal(0, 1.0, (n, 3))
    axis /= np.linalg.norm(axis, axis=1, keepdims=True)
    Rp = rodrigues(axis * mag[:, None])
    out = np.repeat(mat[None].astype(np.float64), n, axis=0)
    out[:, :3, :3] = Rp @ out[:, :3, :3]
    out[:, :3, 3] += rng.normal(0, 0.01, (n, 3))
    return out


def make_object(seed: int, n_pts: int = 1000):
    """Ellipsoid model cloud: points, outward unit normals, colours in [0,1]."""
    rng = np.random.default_rng(1000 + seed)
    axes = rng.uniform(0.03, 0.12, 3)
    d = rng.normal(size=(n_pts, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    pts = d * axes
    nrm = d / axes
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    # smooth-ish colour field so HSV differences are non-trivial
    base = rng.uniform(0, 1, 3)
    cols = np.clip(base + 0.35 * np.sin(d @ rng.normal(size=(3, 3)) * 3.0), 0, 1)
    return pts.astype(np.float64), cols.astype(np.float64), nrm.astype(np.float64), axes


def _lowpass_noise(rng, H, W, ch):
    """uint8 low-pass filtered uniform noise, (H,W,ch)."""
    small = rng.uniform(0, 255, (H // 8 + 2, W // 8 + 2, ch))
    up = np.repeat(np.repeat(small, 8, axis=0), 8, axis=1)[:H, :W]
    fine = rng.uniform(-20, 20, (H, W, ch))
    return np.clip(up + fine, 0, 255).astype(np.uint8)


def make_frame(seed: int, intr: str = "lmo"):
    """Background RGB-D frame: ground plane at 1 m plus a few boxes, 2 mm noise, 5 % holes."""
    H, W, fx, fy, cx, cy = INTRINSICS[intr]
    rng = np.random.default_rng(2000 + seed)
    depth = np.full((H, W), 1.0, dtype=np.float64)
    for _ in range(int(rng.integers(3, 9))):
        x0, y0 = int(rng.integers(0, W - 40)), int(rng.integers(0, H - 40))
        w, h = int(rng.integers(30, W // 3)), int(rng.integers(30, H // 3))
        depth[y0:y0 + h, x0:x0 + w] = np.minimum(depth[y0:y0 + h, x0:x0 + w], rng.uniform(0.5, 1.5))
    depth += rng.normal(0, 0.002, (H, W))
    img = _lowpass_noise(rng, H, W, 3)
    return img, depth, rng


def place_object(img, depth, K, pts, cols, axes, pose, rng):
    """Render the ellipsoid at ``pose`` into the frame by z-buffered splatting."""
    H, W = depth.shape
    d = rng.normal(size=(120000, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    dense = d * axes
    pc = dense @ pose[:3, :3].T + pose[:3, 3]
    z = pc[:, 2]
    ok = z > 0.05
    u = np.rint(pc[:, 0] / np.where(ok, z, 1) * K[0, 0] + K[0, 2]).astype(np.int64)
    v = np.rint(pc[:, 1] / np.where(ok, z, 1) * K[1, 1] + K[1, 2]).astype(np.int64)
    ok &= (u >= 0) & (u < W) & (v >= 0) & (v < H)
    u, v, z, dd = u[ok], v[ok], z[ok], d[ok]
    order = np.argsort(-z)            # far first, near overwrites
    u, v, z, dd = u[order], v[order], z[order], dd[order]
    # colour of the nearest model point direction (cheap: reuse the analytic field via nearest sample)
    idx = np.argmax(dd @ (pts / np.linalg.norm(pts, axis=1, keepdims=True)).T[:, :256], axis=1)
    front = z < depth[v, u]
    depth[v[front], u[front]] = z[front]
    img[v[front], u[front]] = np.clip(cols[idx[front]] * 255.0, 0, 255).astype(np.uint8)


def make_hypotheses(rng, gt_pose, n, K, H, W):
    """50 % perturbed GT, 40 % uniform in the frustum, 10 % adversarial."""
    n_adv = n // 10
    n_rand = (n * 4) // 10
    n_pert = n - n_adv - n_rand
    out = [perturb_pose(rng, gt_pose, n_pert)]
    rnd = np.repeat(np.eye(4)[None], n_rand, axis=0)
    rnd[:, :3, :3] = random_rotations(rng, n_rand)
    zz = rng.uniform(0.4, 1.6, n_rand)
    uu, vv = rng.uniform(0, W, n_rand), rng.uniform(0, H, n_rand)
    rnd[:, 0, 3] = (uu - K[0, 2]) / K[0, 0] * zz
    rnd[:, 1, 3] = (vv - K[1, 2]) / K[1, 1] * zz
    rnd[:, 2, 3] = zz
    out.append(rnd)
    adv = np.repeat(np.eye(4)[None], n_adv, axis=0)   # identity placeholders, online_learning.py:431
    for i in range(n_adv):
        kind = i % 5
        if kind == 1:                                 # behind the camera
            adv[i] = gt_pose; adv[i, 2, 3] = -abs(gt_pose[2, 3])
        elif kind == 2:                               # straddling z = 0
            adv[i] = gt_pose; adv[i, 2, 3] = 0.01
        elif kind == 3:                               # far off-frame
            adv[i] = gt_pose; adv[i, 0, 3] += 5.0
        elif kind == 4:                               # grazing the image border
            adv[i] = gt_pose; adv[i, 0, 3] = (0 - K[0, 2]) / K[0, 0] * gt_pose[2, 3]
    out.append(adv)
    hyp = np.concatenate(out, axis=0)
    perm = rng.permutation(len(hyp))
    return hyp[perm]


def make_scene(seed: int, intr: str = "lmo", n_obj: int = 1, n_pts: int = 1000,
               n_hypo: int = 1000, holes: float = 0.05):
    """One frame plus ``n_obj`` objects, each with GT pose and ``n_hypo`` hypotheses.

    Returns a dict: img, depth (float32), cam_K, objects=[{model_points, model_colors,
    model_normals, gt_pose, pose_hypos}, ...].
    """
    H, W, *_ = INTRINSICS[intr]
    K = cam_K(intr)
    img, depth, rng = make_frame(seed, intr)
    objects = []
    for o in range(n_obj):
        pts, cols, nrm, axes = make_object(seed * 131 + o, n_pts)
        gt = np.eye(4)
        gt[:3, :3] = random_rotations(rng, 1)[0]
        z = rng.uniform(0.45, 0.9)
        u, v = rng.uniform(0.2 * W, 0.8 * W), rng.uniform(0.2 * H, 0.8 * H)
        gt[:3, 3] = [(u - K[0, 2]) / K[0, 0] * z, (v - K[1, 2]) / K[1, 1] * z, z]
        place_object(img, depth, K, pts, cols, axes, gt, rng)
        objects.append(dict(model_points=pts, model_colors=cols, model_normals=nrm, gt_pose=gt))
    holes_mask = rng.uniform(size=depth.shape) < holes
    depth[holes_mask] = 0.0
    for ob in objects:
        ob["pose_hypos"] = make_hypotheses(rng, ob["gt_pose"], n_hypo, K, H, W)
    return dict(img=img, depth=depth.astype(np.float32), cam_K=K, objects=objects, H=H, W=W)


def gt_box(scene, obj, expand: float = 1.2):
    """DTOID-style box crop: GT projection bbox grown by ``expand``.

    Growth rule follows ``expandBox`` (python/ossid/utils/__init__.py:11-16).
    """
    K, H, W = scene["cam_K"], scene["H"], scene["W"]
    pc = obj["model_points"] @ obj["gt_pose"][:3, :3].T + obj["gt_pose"][:3, 3]
    u = pc[:, 0] / pc[:, 2] * K[0, 0] + K[0, 2]
    v = pc[:, 1] / pc[:, 2] * K[1, 1] + K[1, 2]
    x1, x2, y1, y2 = u.min(), u.max(), v.min(), v.max()
    cx, cy, w, h = (x1 + x2) / 2, (y1 + y2) / 2, x2 - x1, y2 - y1
    x1, x2 = max(0, cx - w / 2 * expand), min(W - 1, cx + w / 2 * expand)
    y1, y2 = max(0, cy - h / 2 * expand), min(H - 1, cy + h / 2 * expand)
    return int(x1), int(y1), int(x2), int(y2)
