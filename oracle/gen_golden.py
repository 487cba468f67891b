"""Generate tests/golden/*.npz by RUNNING the reference's own code in this container.

Run from the repo root (needs /root/reference, which only exists here):

    python -m oracle.gen_golden

What is executed from /root/reference (unmodified):
  * ``ossid.utils.zephyr_utils.networkInference`` and ``filterHypoByMask``
    (python/ossid/utils/zephyr_utils.py:10-71), imported normally after two
    environment shims (a dead ``numpy.lib.type_check`` import at
    python/ossid/utils/__init__.py:4 and a provider for the absent
    ``zephyr.utils.projectPointsUv``);
  * ``projectModelPoint`` (python/ossid/datasets/ycbv_sift_dataset.py:303-334) and
    ``kptProjGridCos`` (python/ossid/datasets/ycbv_object.py:63-77), whose modules
    cannot be imported (faiss, oriented_features, ... are absent): the two function
    definitions are extracted from the source text with ``ast`` and executed with
    numpy only.  No reference source is copied into this repo.

The reference fragments compute in float64, the frozen oracle spec in float32
(SURVEY.md Appendix C #1).  For the projection fixture, model points whose f64
projection falls within AMBIG px of a rounding boundary for any fixture pose are
dropped at generation time so that the comparison is bit-exact; the number
dropped is stored in the fixture.
"""
from __future__ import annotations

import ast
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/python"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
AMBIG = 2e-3


def import_reference(project_fn):
    """Import the reference glue with the two shims; ``project_fn`` backs zephyr.utils.projectPointsUv."""
    m = types.ModuleType("numpy.lib.type_check")
    m.imag = np.imag
    sys.modules["numpy.lib.type_check"] = m
    z, zu = types.ModuleType("zephyr"), types.ModuleType("zephyr.utils")
    zu.projectPointsUv = project_fn
    z.utils = zu
    sys.modules["zephyr"], sys.modules["zephyr.utils"] = z, zu
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import ossid.utils.zephyr_utils as ref_glue
    return ref_glue


def extract_function(path, name):
    """Compile one top-level function out of a reference source file (numpy in scope only)."""
    src = open(path).read()
    tree = ast.parse(src)
    node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)
    ns = {"np": np}
    exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return ns[name]


def extract_box_mask_block(path):
    """The reference's box -> mask statements (python/ossid/scripts/online_learning.py:389-405: ``dtoid_mask =
    np.zeros_like(depth)`` ... the ``for`` over ``final_bbox, final_score``), compiled out of its ``main()``.  Returns a
    function (depth, final_bbox, final_score, expandBox) -> dtoid_mask that executes exactly those statements."""
    tree = ast.parse(open(path).read())
    block = None
    for node in ast.walk(tree):
        if isinstance(node, ast.If) and node.orelse and isinstance(node.orelse[0], ast.Assign):
            tgt = node.orelse[0].targets[0]
            if isinstance(tgt, ast.Name) and tgt.id == "dtoid_mask" and "zeros_like" in ast.unparse(node.orelse[0].value):
                block = node.orelse
                break
    assert block is not None and any(isinstance(n, ast.For) for n in block), "box->mask block not found in the reference"
    code = compile(ast.Module(body=block, type_ignores=[]), path, "exec")

    def run(depth, final_bbox, final_score, expandBox):
        ns = {"np": np, "depth": depth, "final_bbox": final_bbox, "final_score": final_score, "expandBox": expandBox}
        exec(code, ns)
        return ns["dtoid_mask"]
    return run


def f32exact(a):
    """Round to float32 and back so that the oracle's single f32 cast is lossless."""
    return np.asarray(a, dtype=np.float32).astype(np.float64)


def gen_projection():
    from ossid_code_b200 import synthetic as syn
    project_model_point = extract_function(f"{REF}/ossid/datasets/ycbv_sift_dataset.py", "projectModelPoint")
    kpt_proj_grid_cos = extract_function(f"{REF}/ossid/datasets/ycbv_object.py", "kptProjGridCos")
    rng = np.random.default_rng(7)
    H, W, fx, fy, cx, cy = syn.INTRINSICS["tiny"]
    meta = {"camera_fx": float(np.float32(fx)), "camera_fy": float(np.float32(fy)),
            "camera_cx": float(np.float32(cx)), "camera_cy": float(np.float32(cy)), "camera_scale": 1.0}
    pts, _, nrm, _ = syn.make_object(3, 400)
    pts, nrm = f32exact(pts), f32exact(nrm)
    base = np.eye(4)
    base[:3, :3] = syn.random_rotations(rng, 1)[0]
    base[:3, 3] = [0.02, -0.01, 0.6]
    poses = syn.perturb_pose(rng, base, 10)
    edge = base.copy(); edge[0, 3] = (2 - cx) / fx * 0.6          # straddles the left image border
    edge2 = base.copy(); edge2[1, 3] = (H - 3 - cy) / fy * 0.6    # straddles the bottom border
    off = base.copy(); off[0, 3] = 3.0                            # entirely off-frame
    poses = f32exact(np.concatenate([poses, edge[None], edge2[None], off[None]], axis=0))
    # drop points that sit in the f32/f64 rounding-ambiguity band for any pose
    keep = np.ones(len(pts), bool)
    for mat in poses:
        tp = pts @ mat[:3, :3].T + mat[:3, 3]
        pr = tp[:, :2] / tp[:, 2:] * np.array([meta["camera_fx"], meta["camera_fy"]]) + np.array([meta["camera_cx"], meta["camera_cy"]])
        frac = np.abs(pr - np.floor(pr) - 0.5)
        keep &= (frac > AMBIG).all(axis=1) & (tp[:, 2] > 0.05)
        # also keep away from the front-facing sign change
        tn = nrm @ mat[:3, :3].T
        keep &= np.abs((-tp * tn).sum(-1)) > 1e-6
    n_drop = int((~keep).sum())
    pts, nrm = pts[keep], nrm[keep]
    img = np.zeros((H, W, 3), np.uint8)
    uv_cat, idx_cat, offs = [], [], [0]
    for mat in poses:
        uv, idx = project_model_point(pts, nrm, meta, mat, img)
        uv_cat.append(uv); idx_cat.append(idx); offs.append(offs[-1] + len(idx))
    # kptProjGridCos: one "grid view" per pose, every point is a keypoint
    grid_metas = [{"poses": mat[:, :, None]} for mat in poses]
    kpts = [np.zeros((len(pts), 2)) for _ in poses]
    kidx = [np.arange(len(pts)) for _ in poses]
    cos_mat = kpt_proj_grid_cos(pts, nrm, grid_metas, kpts, kidx)
    np.savez_compressed(
        os.path.join(OUT, "projection.npz"),
        model_points=pts, model_normals=nrm, poses=poses, H=H, W=W,
        fx=meta["camera_fx"], fy=meta["camera_fy"], cx=meta["camera_cx"], cy=meta["camera_cy"],
        ref_uv=np.concatenate(uv_cat).astype(np.int32), ref_idx=np.concatenate(idx_cat).astype(np.int32),
        ref_offsets=np.asarray(offs, np.int64), ref_cos=cos_mat, n_dropped_ambiguous=n_drop, ambig_px=AMBIG)
    print(f"projection.npz: {len(pts)} points x {len(poses)} poses, dropped {n_drop} ambiguous points")


def gen_mask_filter(ref_glue):
    from ossid_code_b200 import synthetic as syn
    sc = syn.make_scene(11, "tiny", n_obj=1, n_pts=200, n_hypo=96)
    ob = sc["objects"][0]
    x1, y1, x2, y2 = syn.gt_box(sc, ob, 1.2)
    mask = np.zeros((sc["H"], sc["W"]), np.int64)
    mask[y1:y2, x1:x2] = 1
    meta = dict(camera_fx=sc["cam_K"][0, 0], camera_fy=sc["cam_K"][1, 1],
                camera_cx=sc["cam_K"][0, 2], camera_cy=sc["cam_K"][1, 2], camera_scale=1.0)
    kept = {}
    for th in (0.5, 0.9, 0.0):
        kept[th] = np.asarray(ref_glue.filterHypoByMask(ob["model_points"], meta, ob["pose_hypos"], mask, th=th))
    np.savez_compressed(
        os.path.join(OUT, "mask_filter.npz"), model_points=ob["model_points"], pose_hypos=ob["pose_hypos"],
        mask=mask.astype(np.uint8), cam_K=sc["cam_K"], box=np.asarray([x1, y1, x2, y2]),
        kept_050=kept[0.5], kept_090=kept[0.9], kept_000=kept[0.0])
    print("mask_filter.npz: kept", {k: int(v.sum()) for k, v in kept.items()}, "of", len(ob["pose_hypos"]))


def gen_boxes_mask():
    """DTOID boxes -> mask: the reference's own statements (AST-extracted from online_learning.py's main()) and its own
    expandBox (imported from ossid.utils), on three detection lists: confident + low-score boxes, only low-score boxes,
    boxes that leave the frame / are empty."""
    from ossid.utils import expandBox
    from ossid_code_b200 import synthetic as syn
    run = extract_box_mask_block(f"{REF}/ossid/scripts/online_learning.py")
    sc = syn.make_scene(13, "tiny", n_obj=1, n_pts=64, n_hypo=4)
    depth = sc["depth"]
    depth[:40, :60] = 0.0                                       # a region without depth: a box there leaves "mask * depth>0" empty
    cases = {
        "a": ([[70.2, 30.7, 110.9, 80.1], [10.0, 5.0, 40.0, 30.0], [120.5, 90.0, 150.0, 118.0]], [0.9, 0.3, 0.7]),
        "b": ([[5.0, 4.0, 30.0, 25.0], [80.0, 60.0, 120.0, 100.0], [20.0, 50.0, 60.0, 90.0]], [0.2, 0.4, 0.1]),
        "c": ([[-20.0, -10.0, 30.0, 40.0], [140.0, 100.0, 200.0, 160.0], [50.0, 50.0, 50.0, 70.0], [90.0, 20.0, 100.0, 30.0]],
              [0.6, 0.8, 0.9, 0.45]),
    }
    out = {"depth": depth}
    for tag, (boxes, scores) in cases.items():
        mask = run(depth, [np.asarray(b) for b in boxes], list(scores), expandBox)
        out[f"{tag}_boxes"], out[f"{tag}_scores"], out[f"{tag}_mask"] = np.asarray(boxes), np.asarray(scores), mask.astype(np.uint8)
        print(f"boxes_mask {tag}: {int(mask.sum())} pixels set")
    np.savez_compressed(os.path.join(OUT, "boxes_mask.npz"), **out)


def gen_network_inference(ref_glue):
    from ossid_code_b200 import synthetic as syn, weights
    from oracle import zephyr_oracle as zo
    sc = syn.make_scene(5, "tiny", n_obj=1, n_pts=256, n_hypo=80)
    ob = sc["objects"][0]
    folded = weights.seeded_folded(0)
    out = {}
    for tag, th in (("th100", 100.0), ("th10", 10.0)):
        data = dict(img=sc["img"], depth=sc["depth"], cam_K=sc["cam_K"], model_colors=ob["model_colors"],
                    model_points=ob["model_points"], model_normals=ob["model_normals"],
                    pose_hypos=ob["pose_hypos"].copy(), pp_err=np.arange(len(ob["pose_hypos"]), dtype=np.float64))
        ds, model = zo.OracleScoreDataset(th), zo.OracleScorer(folded)
        poses, scores, errs, uv, dt = ref_glue.networkInference(model, ds, data, return_time=True)
        out[f"{tag}_poses"] = poses
        out[f"{tag}_scores"] = np.asarray(scores, np.float32).reshape(-1)
        out[f"{tag}_pp_err"] = np.asarray(errs)
        out[f"{tag}_uv"] = uv.numpy().astype(np.int16)
        out[f"{tag}_mask"] = ds.last["mask"][ds.last["keep"]].numpy()
        if tag == "th10":                       # features of the filtered run only (keeps the fixture small)
            out[f"{tag}_point_x"] = ds.last["point_x"][ds.last["keep"]].numpy()
        print(f"network_inference {tag}: kept {len(scores)} of {len(ob['pose_hypos'])}, argmax {int(np.argmax(scores))}")
    np.savez_compressed(
        os.path.join(OUT, "network_inference.npz"), img=sc["img"], depth=sc["depth"], cam_K=sc["cam_K"],
        model_points=ob["model_points"], model_colors=ob["model_colors"], model_normals=ob["model_normals"],
        pose_hypos=ob["pose_hypos"], weight_seed=0, **out)


def main():
    from oracle import zephyr_oracle as zo
    os.makedirs(OUT, exist_ok=True)
    ref_glue = import_reference(lambda poses, pts, meta: zo.project_raw(poses, pts, meta).numpy())
    torch.set_num_threads(4)
    gen_projection()
    gen_mask_filter(ref_glue)
    gen_boxes_mask()
    gen_network_inference(ref_glue)


if __name__ == "__main__":
    main()
