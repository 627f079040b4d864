"""One conv launch for source-level ncu captures: python scripts/run_one_conv.py B Cin Cout D k stride T"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coma_unet_b200 import ops
B, Cin, Cout, D, k, s, T = (int(v) for v in sys.argv[1:8])
x = torch.randn(B, D, D, D, Cin, device="cuda").bfloat16()
w = torch.randn(Cin, Cout, k, k, k, device="cuda") if T else torch.randn(Cout, Cin, k, k, k, device="cuda")
wp = ops.pack_weight(w, bool(T), Cin, Cout, torch.bfloat16)
for _ in range(3):
    y, st = ops.conv_raw(x, wp, None, ksize=k, stride=s, transposed=bool(T), want_stats=True)
torch.cuda.synchronize()
print("done")
