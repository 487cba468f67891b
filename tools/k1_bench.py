"""Micro-benchmark of the feature kernel alone (CUDA events, L2-exceeding outputs).  Run on the B200 box."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ossid_code_b200 import synthetic as syn, zephyr_utils as glue  # noqa: E402
from ossid_code_b200.engine import get_context, poses_to_rt12  # noqa: E402

n_hyp = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
n_pts = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
intr = sys.argv[3] if len(sys.argv) > 3 else "ycbv"
ctx = get_context(0)
sc = syn.make_scene(1, intr, n_obj=1, n_pts=n_pts, n_hypo=10000)
ob = sc["objects"][0]
import numpy as np
rng = np.random.default_rng(5)
P = np.concatenate([syn.make_hypotheses(rng, ob["gt_pose"], 10000, sc["cam_K"], sc["H"], sc["W"])
                    for _ in range(-(-n_hyp // 10000))])[:n_hyp]
ctx.set_frame_u8(sc["img"], sc["depth"], glue.K2meta(sc["cam_K"]))
if os.environ.get("ZS_SORT", "1") == "1":
    from ossid_code_b200.scoring import spatial_order
    perm = spatial_order(ob["model_points"])
    ob = {k: (v[perm] if k.startswith("model_") else v) for k, v in ob.items()}
    print("model points in Morton order")
ctx.set_object(0, ob["model_points"], ob["model_colors"], ob["model_normals"])
if os.environ.get("ZS_CROP", "0") != "0":
    # C3-style: every hypothesis inside the DTOID-style box crop
    box = syn.gt_box(sc, ob, 1.2)
    mask = np.zeros((sc["H"], sc["W"]), np.uint8); mask[box[1]:box[3], box[0]:box[2]] = 1
    kept = []
    while sum(len(k) for k in kept) < n_hyp:
        c = syn.make_hypotheses(rng, ob["gt_pose"], 20000, sc["cam_K"], sc["H"], sc["W"])
        kept.append(c[glue.filterHypoByMask(ob["model_points"], glue.K2meta(sc["cam_K"]), c, mask, th=0.5)])
    P = np.concatenate(kept)[:n_hyp]
    print(f"hypotheses inside box {box} ({box[2]-box[0]} x {box[3]-box[1]} px, {(box[2]-box[0])*(box[3]-box[1])*16/1024:.0f} KB)")
    if os.environ.get("ZS_CROP_STAGE", "0") != "0":
        # experiment build (-DZS_CROP_STAGE, run with ZS_LIB=gpurun_ab/libzs_crop.so): stage the centre of the box in
        # shared memory, at most side x side pixels (104 x 104 x 16 B = 169 KB next to the 36 KB cloud)
        side = int(os.environ.get("ZS_CROP_SIDE", "104"))
        w, h = min(side, box[2] - box[0]), min(side, box[3] - box[1])
        x0, y0 = (box[0] + box[2] - w) // 2, (box[1] + box[3] - h) // 2
        os.environ["ZS_CROP_RECT"] = f"{x0},{y0},{w},{h}"
        print(f"crop staged in shared memory: ZS_CROP_RECT={os.environ['ZS_CROP_RECT']} ({w * h * 16 / 1024:.0f} KB)")
if os.environ.get("ZS_SORT_HYP", "0") == "1":
    # hypotheses ordered by the image tile their translation projects to (what co-resident warps then share in L1)
    K = sc["cam_K"]
    t = P[:, :3, 3]
    z = np.where(np.abs(t[:, 2]) > 1e-6, t[:, 2], 1e-6)
    u = np.clip(t[:, 0] / z * K[0, 0] + K[0, 2], 0, sc["W"] - 1).astype(np.int64) // 32
    v = np.clip(t[:, 1] / z * K[1, 1] + K[1, 2], 0, sc["H"] - 1).astype(np.int64) // 32
    P = P[np.argsort(v * 64 + u, kind="stable")]
    print("hypotheses sorted by projected 32x32 image tile")
p12 = poses_to_rt12(P, ctx.device)
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6536.4
for dtype, aux in ((torch.bfloat16, False), (torch.float32, False), (torch.bfloat16, True), (torch.float32, True)):
    feat = torch.empty((n_hyp, n_pts, 8), dtype=dtype, device=ctx.device)
    kw = dict(want_mask=True) if aux else {}
    for _ in range(3):
        ctx.features(0, p12, out=feat, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        ctx.features(0, p12, out=feat, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    byts = n_hyp * (48 + n_pts * (8 * feat.element_size() + (1 if aux else 0)))
    print(f"{str(dtype):16s} aux={aux!s:5s} {n_hyp} x {n_pts}: {ms*1e3:8.1f} us  {byts/ms/1e6:8.1f} GB/s  "
          f"{byts/ms/1e6/peak*100:5.1f}% of {peak} GB/s   {n_hyp/ms/1e3:7.2f} M hyp/s")
