"""Run one conv configuration a few times (for ncu / timing):  python scripts/run_conv.py B Cin Cout D k stride T [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coma_unet_b200 import _lib as L
from coma_unet_b200 import ops

B, Cin, Cout, D, k, s, T = (int(v) for v in sys.argv[1:8])
reps = int(sys.argv[8]) if len(sys.argv) > 8 else 5
stats = int(os.environ.get("STATS", "0"))
x = torch.randn(B, D, D, D, Cin, device="cuda").bfloat16()
w = torch.randn(Cin, Cout, k, k, k, device="cuda") if T else torch.randn(Cout, Cin, k, k, k, device="cuda")
wp = ops.pack_weight(w, bool(T), Cin, Cout, torch.bfloat16)
scale, shift = torch.ones(B, Cout, device="cuda"), torch.zeros(B, Cout, device="cuda")
for _ in range(2):
    y, _ = ops.conv_raw(x, wp, None, ksize=k, stride=s, transposed=bool(T), scale=scale, shift=shift, act=L.ACT_RELU, want_stats=bool(stats))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    y, _ = ops.conv_raw(x, wp, None, ksize=k, stride=s, transposed=bool(T), scale=scale, shift=shift, act=L.ACT_RELU, want_stats=bool(stats))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
vox = (D * s if T else D // s) ** 3 if not T else D ** 3
flops = 2.0 * B * (D ** 3 if T else (D // s) ** 3) * k ** 3 * Cin * Cout
print(f"conv B{B} {Cin}->{Cout} D{D} k{k} s{s} T{T}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s")
