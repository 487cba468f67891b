"""Randomised shape sweep of the three tensor-core MLP paths on the GPU (tile boundaries, short hypotheses, hypothesis
counts around the CTA-pair / pair-group counts):
  bf16   zs_k_mlp_tc<false>  vs a torch fp32 evaluation of the same bf16-rounded network            (2^-5 of max)
  split  zs_k_mlp_tc3        vs a torch fp64 evaluation of the fp32 network                          (5e-5 of max)
  fused  zs_k_mlp_tc<true>   vs zs_features(bf16) -> zs_k_mlp_tc<false> on random poses, several segments (bit-exact)
    python tools/k2_fuzz.py [seed] [trials]"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ossid_code_b200 import synthetic as syn, weights, zephyr_utils as glue
from ossid_code_b200.engine import get_context, poses_to_rt12, split_bf16
torch.backends.cuda.matmul.allow_tf32 = False
ctx = get_context(0)
w = weights.seeded_folded(3)
ctx.set_weights(0, w)
dev = ctx.device
bf = lambda t: t.to(torch.bfloat16).to(torch.float32)
W1, W2, W3 = (bf(w[k]).to(dev) for k in ("W1", "W2", "W3"))
D1, D2, D3 = (w[k].double().to(dev) for k in ("W1", "W2", "W3"))
b1, b2, b3 = (w[k].to(dev) for k in ("b1", "b2", "b3"))
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
g = torch.Generator(device="cpu").manual_seed(seed)
trials = int(sys.argv[2]) if len(sys.argv) > 2 else 80
worst = {"bf16": 0.0, "split": 0.0}
special = [(1, 1), (73, 255), (74, 256), (75, 257), (147, 128), (148, 129), (149, 511), (2, 1000), (296, 513), (1, 4096),
           (17, 1000), (18, 1000), (19, 127), (71, 129), (72, 385), (73, 1024)]
for t in range(trials):
    if t < len(special):
        n, N = special[t]
    else:
        n = int(torch.randint(1, 500, (1,), generator=g)); N = int(torch.randint(1, 1600, (1,), generator=g))
    x32 = torch.randn(n, N, 8, generator=g) * 0.5
    x = x32.to(torch.bfloat16).to(dev)
    pooled = ctx.pool(0, x)
    xf = x.to(torch.float32)
    h1 = bf(torch.relu(xf @ W1.T + b1))
    h2 = bf(torch.relu(h1 @ W2.T + b2))
    ref = torch.relu((h2 @ W3.T).amax(dim=1) + b3)
    err = float((pooled - ref).abs().max()) / (float(ref.abs().max()) + 1e-9)
    worst["bf16"] = max(worst["bf16"], err)
    if err > 2 ** -5:
        sys.exit(f"MISMATCH bf16 n={n} N={N}: rel err {err:.3e}")
    pooled3 = ctx.pool(0, split_bf16(x32).to(dev))
    xd = x32.double().to(dev)
    r3 = torch.relu((torch.relu(torch.relu(xd @ D1.T + b1) @ D2.T + b2) @ D3.T).amax(dim=1) + b3)
    err3 = float((pooled3.double() - r3).abs().max()) / (float(r3.abs().max()) + 1e-9)
    worst["split"] = max(worst["split"], err3)
    if err3 > 5e-5:
        sys.exit(f"MISMATCH split n={n} N={N}: rel err {err3:.3e}")
print(f"{trials} shapes ok: worst relative error of the pooled vector bf16 {worst['bf16']:.3e}, split {worst['split']:.3e}")

# fused kernel vs the two-kernel sequence on real frames: random cloud sizes >= 128, 1-40 segments, random counts
rng = np.random.default_rng(seed)
n_f = max(trials // 8, 4)
for t in range(n_f):
    N = int(rng.integers(128, 1300)) if t else 128
    n_seg = int(rng.integers(1, 41))
    sc = syn.make_scene(100 + seed * 31 + t, "lmo", n_obj=1, n_pts=N, n_hypo=64)
    ob = sc["objects"][0]
    ctx.set_frame_u8(sc["img"], sc["depth"], glue.K2meta(sc["cam_K"]))
    segs, refs = [], []
    for s in range(n_seg):
        ctx.set_object(s, ob["model_points"], ob["model_colors"], ob["model_normals"])
        m = int(rng.integers(0, 64)) if s else 64
        p12 = poses_to_rt12(ob["pose_hypos"][rng.permutation(64)[:m]], dev)
        segs.append((s, p12))
        if m:
            refs.append(ctx.pool(0, ctx.features(s, p12, dtype=torch.bfloat16)[0]))
    ref = torch.cat(refs)
    out = torch.empty_like(ref)
    ctx.pool_fused(0, segs, out=out)
    if not torch.equal(out, ref):
        sys.exit(f"MISMATCH fused N={N} segments={n_seg}: {int((out != ref).sum())} values differ")
print(f"{n_f} fused launches bit-identical to zs_features -> zs_pool")
