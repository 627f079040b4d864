"""Time coma_gate_fwd with dense and concat-buffer (channel-strided) operands."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coma_unet_b200 import ops


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


C, D = 32, 128
F = C // 2
cat = torch.randn(8, D, D, D, 2 * C, device="cuda").bfloat16()
x = torch.randn(8, D, D, D, C, device="cuda").bfloat16()
gd = torch.randn(8, D, D, D, C, device="cuda").bfloat16()
od = torch.empty_like(x)
wg, wx = torch.randn(F, C, device="cuda") * 0.1, torch.randn(F, C, device="cuda") * 0.1
bsum, wpsi, bpsi = torch.randn(F, device="cuda"), torch.randn(F, device="cuda"), torch.zeros(1, device="cuda")
gb = 3 * x.numel() * 2 / 1e9
for name, g, o in (("dense g, dense out", gd, od), ("strided g (concat right half), dense out", cat[..., C:], od),
                   ("dense g, strided out (concat left half)", gd, cat[..., :C]), ("strided g, strided out (model)", cat[..., C:], cat[..., :C])):
    ms = timeit(lambda: ops.gate_fused(g, x, wg, wx, bsum, wpsi, bpsi, out=o))
    print(f"{name}: {ms:.4f} ms  {gb / ms * 1e3:.0f} GB/s", flush=True)
ms = timeit(lambda: od.copy_(x))
print(f"torch copy (2 units): {2 * x.numel() * 2 / 1e9 / ms * 1e3:.0f} GB/s")
ms = timeit(lambda: torch.add(x, gd, out=od))
print(f"torch add (3 units): {gb / ms * 1e3:.0f} GB/s")
