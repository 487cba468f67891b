"""Micro-benchmark of the tensor-core MLP kernel alone (zs_pool on random bf16 features)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ossid_code_b200 import weights
from ossid_code_b200.engine import get_context
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
ctx = get_context(0)
ctx.set_weights(0, weights.seeded_folded(0))
x = (torch.randn(n, N, 8, device=ctx.device) * 0.5).to(torch.bfloat16)
pooled = torch.empty(n, 1024, device=ctx.device)
for _ in range(3): ctx.pool(0, x, out=pooled)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(iters): ctx.pool(0, x, out=pooled)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / iters
fl = 2.0 * n * N * (8 * 64 + 64 * 128 + 128 * 1024)
tiles = n / 74 * -(-N // 128)
print(f"iters={iters} exp={os.environ.get('ZS_TC_EXPERIMENT','0')} n={n} N={N}: {ms:.3f} ms  {fl/ms/1e9:.1f} TFLOP/s (algorithmic)  "
      f"{ms*1e-3*1.965e9/tiles:.0f} cycles/tile @1965MHz  {n/ms/1e3:.2f} M hyp/s")
