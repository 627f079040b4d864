"""One weight gradient through the C ABI, timed with CUDA events (for ncu captures of the tcgen05 weight-gradient kernel).

    python scripts/run_wgrad.py [B Cg Cx D]      # default 4 32 64 128: the merge conv of the top level (dy 32 ch, x 64 ch)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coma_unet_b200 import ops

B, Cg, Cx, D = (int(v) for v in sys.argv[1:5]) if len(sys.argv) >= 5 else (4, 32, 64, 128)
gen = torch.Generator(device="cuda").manual_seed(0)
g = torch.randn(B, D, D, D, Cg, device="cuda", generator=gen).bfloat16()
x = torch.randn(B, D, D, D, Cx, device="cuda", generator=gen).bfloat16()
for _ in range(3):
    dw = ops.wgrad_raw(g, x, ksize=3, stride=1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for _ in range(n):
    dw = ops.wgrad_raw(g, x, ksize=3, stride=1)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
flops = 2.0 * 27 * Cg * Cx * B * D ** 3
print(f"wgrad B={B} Cg={Cg} Cx={Cx} {D}^3: {ms:.3f} ms per call (incl. the zero-fill of dw), {flops / ms / 1e9:.0f} TFLOP/s, "
      f"algorithmic bytes {(Cg + Cx) * 2 * B * D ** 3 / 1e6:.0f} MB")
# spot check against fp32 torch on one tap (centre): dw[13][cg][cx] = sum_v g[v][cg] x[v][cx]
ref = torch.einsum("bdhwg,bdhwx->gx", g[:1, :8].float(), x[:1, :8].float())
got = ops.wgrad_raw(g[:1, :8].contiguous(), x[:1, :8].contiguous(), ksize=3, stride=1)[13]
print("centre-tap check (1 x 8 x D x D slab):", float((got - ref).abs().max() / ref.abs().max()))
