"""Repeatability stress: the same convolution launched many times must give bit-identical outputs and statistics (a missed
barrier / phase bug in the warp-specialised kernels shows up as run-to-run differences)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coma_unet_b200 import _lib as L
from coma_unet_b200 import ops

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
CASES = [(8, 128, 64, 64, 3, 1, 0), (8, 64, 32, 128, 3, 1, 0), (8, 64, 64, 64, 3, 1, 0), (8, 16, 16, 128, 3, 1, 0), (8, 32, 32, 128, 3, 1, 0),
         (8, 32, 64, 128, 3, 2, 0), (8, 64, 32, 64, 3, 2, 1), (8, 128, 128, 32, 3, 1, 0), (3, 128, 32, 40, 3, 1, 0), (8, 1, 32, 64, 3, 1, 0)]
bad = 0
for B, Cin, Cout, D, k, s, T in CASES:
    g = torch.Generator(device="cuda").manual_seed(Cin * 1000 + Cout)
    x = torch.randn(B, D, D, D, Cin, device="cuda", generator=g).bfloat16()
    w = torch.randn(Cin, Cout, k, k, k, device="cuda", generator=g) if T else torch.randn(Cout, Cin, k, k, k, device="cuda", generator=g)
    wp = ops.pack_weight(w * (Cin * 27) ** -0.5, bool(T), Cin, Cout, torch.bfloat16)
    ref_y, ref_s = ops.conv_raw(x, wp, None, ksize=k, stride=s, transposed=bool(T), want_stats=True)
    ref_s = ref_s.sum(dim=1).clone()
    ref_y = ref_y.clone()
    diff = 0
    for i in range(reps):
        y, st = ops.conv_raw(x, wp, None, ksize=k, stride=s, transposed=bool(T), want_stats=True)
        if not torch.equal(y, ref_y) or not torch.equal(st.sum(dim=1), ref_s):
            diff += 1
    bad += diff
    print(f"B{B} {Cin}->{Cout} D{D} s{s} T{T}: {reps} launches, {diff} differing from the first", flush=True)
print("REPEATABLE" if bad == 0 else f"NOT REPEATABLE: {bad}")
sys.exit(0 if bad == 0 else 1)
