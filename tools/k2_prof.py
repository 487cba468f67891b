"""Per-role barrier-wait breakdown of the tensor-core MLP kernel.

Needs a library built from the instrumented source:
    nvcc <flags of ossid_code_b200/build.py> -DZS_TC_PROF -o gpurun_ab/libzs_prof.so ossid_code_b200/csrc/*.cu -lcuda
    ZS_LIB=gpurun_ab/libzs_prof.so python tools/k2_prof.py [n] [n_pts]
Counters are SM cycles (clock64) of one thread per role, averaged over CTAs, per 128-point tile.
"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ossid_code_b200 import weights
from ossid_code_b200.engine import get_context
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
ctx = get_context(0)
ctx.set_weights(0, weights.seeded_folded(0))
x = (torch.randn(n, N, 8, device=ctx.device) * 0.5).to(torch.bfloat16)
for _ in range(2):
    pooled, h1, h2 = ctx.pool_debug(0, x)
torch.cuda.synchronize()
n_cta = 148
p = h2.reshape(-1)[: n_cta * 24 * 2].view(torch.int64).reshape(n_cta, 24).cpu().double()
tiles = n / (n_cta // 2) * -(-N // 256)   # pair-tiles (256 points) per CTA
rows = [("issuer total", 0), ("issuer wait a3_full (H2 ready)", 1), ("issuer wait d3_empty", 2), ("issuer wait a2_full (H1 ready)", 3),
        ("issuer wait x_full", 4), ("front total", 8), ("front wait a3_empty", 9), ("front wait d1_full", 10),
        ("front wait d2_full", 11), ("maxpool total", 16), ("maxpool wait d3_full", 17), ("maxpool tmem load + wait::ld", 18),
        ("maxpool arrive d3_empty", 20)]
print(f"n={n} N={N} tiles per CTA {tiles:.0f}")
for nm, i in rows:
    col = p[:, i]
    col = col[col > 0] if i < 8 else col          # issuer counters exist in leader CTAs only
    print(f"{nm:34s} {col.mean().item() / tiles:9.0f} cycles/pair-tile")
