"""Launch a list of conv configurations once each (for one ncu capture):
python scripts/run_layers.py "B,Cin,Cout,D,k,stride,T;B,Cin,..." [launches_per_config]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coma_unet_b200 import _lib as L
from coma_unet_b200 import ops

reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
for spec in sys.argv[1].split(";"):
    B, Cin, Cout, D, k, s, T = (int(v) for v in spec.split(","))
    x = torch.randn(B, D, D, D, Cin, device="cuda").bfloat16()
    w = torch.randn(Cin, Cout, k, k, k, device="cuda") if T else torch.randn(Cout, Cin, k, k, k, device="cuda")
    wp = ops.pack_weight(w, bool(T), Cin, Cout, torch.bfloat16)
    scale, shift = torch.ones(B, Cout, device="cuda"), torch.zeros(B, Cout, device="cuda")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        y, _ = ops.conv_raw(x, wp, None, ksize=k, stride=s, transposed=bool(T), scale=scale, shift=shift, act=L.ACT_RELU)
    e1.record()
    torch.cuda.synchronize()
    flops = 2.0 * B * (D ** 3 if T else (D // s) ** 3) * k ** 3 * Cin * Cout
    ms = e0.elapsed_time(e1) / reps
    print(f"conv B{B} {Cin}->{Cout} D{D} k{k} s{s} T{T}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s", flush=True)
    del x, y
