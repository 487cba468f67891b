"""Randomised shape sweep of the tensor-core MLP kernel against a torch fp32 evaluation of the same bf16-rounded
network on the GPU (tile boundaries, short hypotheses, hypothesis counts around the CTA-pair count)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ossid_code_b200 import weights
from ossid_code_b200.engine import get_context
torch.backends.cuda.matmul.allow_tf32 = False
ctx = get_context(0)
w = weights.seeded_folded(3)
ctx.set_weights(0, w)
dev = ctx.device
bf = lambda t: t.to(torch.bfloat16).to(torch.float32)
W1, W2, W3 = (bf(w[k]).to(dev) for k in ("W1", "W2", "W3"))
b1, b2, b3 = (w[k].to(dev) for k in ("b1", "b2", "b3"))
g = torch.Generator(device="cpu").manual_seed(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
trials = int(sys.argv[2]) if len(sys.argv) > 2 else 80
worst = 0.0
special = [(1, 1), (73, 255), (74, 256), (75, 257), (147, 128), (148, 129), (149, 511), (2, 1000), (296, 513), (1, 4096)]
for t in range(trials):
    if t < len(special):
        n, N = special[t]
    else:
        n = int(torch.randint(1, 500, (1,), generator=g)); N = int(torch.randint(1, 1600, (1,), generator=g))
    x = (torch.randn(n, N, 8, generator=g) * 0.5).to(torch.bfloat16).to(dev)
    pooled = ctx.pool(0, x)
    xf = x.to(torch.float32)
    h1 = bf(torch.relu(xf @ W1.T + b1))
    h2 = bf(torch.relu(h1 @ W2.T + b2))
    ref = torch.relu((h2 @ W3.T).amax(dim=1) + b3)
    err = float((pooled - ref).abs().max()) / (float(ref.abs().max()) + 1e-9)
    worst = max(worst, err)
    if err > 2 ** -5:
        print(f"MISMATCH n={n} N={N}: rel err {err:.3e}")
        sys.exit(1)
print(f"{trials} shapes ok, worst relative error of the pooled vector {worst:.3e}")
