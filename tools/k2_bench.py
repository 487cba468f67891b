"""Throughput of the two tensor-core MLP kernels alone (CUDA events): python tools/k2_bench.py [n] [N] [reps].
bf16: zs_k_mlp_tc on (n,N,8) bf16 features; split: zs_k_mlp_tc3 (fp32-accurate, 3-term) on (n,2,N,8) planes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ossid_code_b200 import weights
from ossid_code_b200.engine import get_context, split_bf16

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
ctx = get_context(0)
ctx.set_weights(0, weights.seeded_folded(0))
x = torch.randn(n, N, 8, generator=torch.Generator().manual_seed(0)) * 0.5
flops = 2.0 * (8 * 64 + 64 * 128 + 128 * 1024) * n * N
for name, feat in (("bf16", x.to(torch.bfloat16).to(ctx.device)), ("split", split_bf16(x).to(ctx.device))):
    out = torch.empty((n, 1024), dtype=torch.float32, device=ctx.device)
    for _ in range(2):
        ctx.pool(0, feat, out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        ctx.pool(0, feat, out=out)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    print(f"{name}: n={n} N={N}: {ms:.3f} ms/launch, {flops / ms / 1e9:.1f} TFLOP/s algorithmic, {n / ms / 1e3:.3f} M hyp/s")
