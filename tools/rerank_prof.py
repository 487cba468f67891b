"""Where the fp32 re-rank of the top-k candidates spends its time (CUDA events around each launch group, C2 shape)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ossid_code_b200 import scoring, synthetic as syn, weights
n_obj, k, N = 21, 8, 1000
sc = syn.make_scene(1, "ycbv", n_obj=n_obj, n_pts=N, n_hypo=64)
fs = scoring.FrameScorer([weights.seeded_folded(0), weights.seeded_folded(1)], device=0, k=k, graph=False)
fs.upload(sc["img"], sc["depth"], sc["cam_K"], sc["objects"], lambda o: o % 2)
S, I = fs.run_resident()
ctx, plan = fs.ctx, fs._plan
P = plan.rec[plan.pose_at:].view(torch.float32).view(n_obj, k, 12)
rr = fs._rr
feat = rr["feat"][: n_obj * k * N * 16].view(n_obj * k, 2, N, 8)
def timed(name, fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    print(f"{name:34s} {a.elapsed_time(b) / reps * 1e3:8.1f} us")
rows = {ws: [o for o in plan.order if plan.wslot[o] == ws] for ws in (0, 1)}
timed("gather_poses", lambda: ctx.gather_poses(fs._p12[fs._buf], I, plan.pose_seg, out=plan.rec[plan.pose_at:].view(torch.float32)))
timed("features_multi (21 objects x 8)", lambda: ctx.features_multi([(o, P[o], feat[plan.rr_row[o] * k: plan.rr_row[o] * k + k]) for o in plan.order]))
for ws, mem in rows.items():
    r0, n = plan.rr_row[mem[0]] * k, len(mem) * k
    timed(f"pool split ws{ws} ({n} hyp)", lambda: ctx.pool(ws, feat[r0: r0 + n], out=rr["pooled"][r0: r0 + n]))
    timed(f"head fp32 small ws{ws} ({n} rows)", lambda: ctx.head(ws, rr["pooled"][r0: r0 + n], False, out=rr["scores"][r0: r0 + n]))
nk = n_obj * k
S2, I2 = rr["out"][:nk].view(torch.float32).view(n_obj, k), rr["out"][nk: 2 * nk].view(n_obj, k)
timed("topk_segments (21 x 8)", lambda: ctx.topk_segments(rr["scores"], plan.rr_seg, k, index_map=I.reshape(-1), out=(S2, I2)))
timed("whole re-rank", lambda: fs._rerank(S, I, P))
