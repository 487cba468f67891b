import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from ossid_code_b200 import synthetic as syn, weights, zephyr_shim, zephyr_utils as glue
sc = syn.make_scene(1, "lmo", n_obj=1, n_pts=1000, n_hypo=1000)
ob = sc["objects"][0]
class A: pass
a = A(); a.inconst_ratio_th = 100
ds = zephyr_shim.ScoreDataset([], "", "lmo", a, mode="test")
m = zephyr_shim.PointNet2SSG(8, a, num_class=1)
m.load_state_dict(weights.seeded_state_dict(0)) if hasattr(weights, "seeded_state_dict") else None
m = m.to(0).eval()
for th in (100, 10):
    ds.inconst_ratio_th = float(th)
    for name, fn in (("mirror", glue.networkInference),):
        ts = []
        for it in range(8):
            data = dict(img=sc["img"], depth=sc["depth"], cam_K=sc["cam_K"], model_points=ob["model_points"],
                        model_colors=ob["model_colors"], model_normals=ob["model_normals"], pose_hypos=ob["pose_hypos"].copy())
            torch.cuda.synchronize(); t0 = time.perf_counter()
            poses, scores, err, uv, dt = fn(m, ds, data, return_time=True)
            torch.cuda.synchronize(); t1 = time.perf_counter()
            _ = glue.to_np(uv)[int(np.argmax(scores))]
            torch.cuda.synchronize(); t2 = time.perf_counter()
            ts.append((t1 - t0, dt, t2 - t1))
        tot, inner, uvt = np.median(np.array(ts), axis=0)
        print(f"th={th} {name}: call {tot*1e3:.2f} ms (timed span {inner*1e3:.2f} ms), uv readback {uvt*1e3:.2f} ms, {len(scores)} scored -> {len(scores)/tot:.0f} hyp/s")

# ---- where the time of one call goes (each stage synchronised; 1,000 hypotheses x 1,000 points) ----
from ossid_code_b200.engine import get_context, poses_to_rt12
ctx = get_context(0)
meta = {k: float(v) for k, v in glue.K2meta(sc["cam_K"]).items()}
img_t, dep_t = torch.from_numpy(sc["img"]), torch.from_numpy(sc["depth"])
tr = torch.from_numpy(ob["pose_hypos"])
pts, cols, nrm = (torch.from_numpy(ob[k]) for k in ("model_points", "model_colors", "model_normals"))
def timed(fn, n=20):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): r = fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n * 1e3, r
t_frame, _ = timed(lambda: ctx.set_frame_u8(img_t, dep_t, meta, blur=True))
t_obj, _ = timed(lambda: ctx.set_object(0, pts, cols, nrm))
t_pose, p12 = timed(lambda: poses_to_rt12(tr, ctx.device))
t_feat, (feat, uv, _, _) = timed(lambda: ctx.features(0, p12, dtype=torch.bfloat16, want_uv=True))
t_feat0, _ = timed(lambda: ctx.features(0, p12, dtype=torch.bfloat16))
m._sync_weights(ctx)
t_score, s = timed(lambda: ctx.score(m._slot, feat))
t_np, _ = timed(lambda: s.cpu().numpy())
print(f"set_frame_u8 {t_frame:.3f}  set_object {t_obj:.3f}  poses {t_pose:.3f}  features+uv {t_feat:.3f} (no uv {t_feat0:.3f})  "
      f"score {t_score:.3f}  scores to host {t_np:.3f}  ms")
