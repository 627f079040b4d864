"""Time coma_norm_film_act_fwd on hot-path shapes (env COMA_AFFINE_CHUNKS / COMA_AFFINE_UNROLL select variants)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from coma_unet_b200 import _lib as L
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


for C, D in ((16, 128), (32, 128), (64, 64)):
    x = torch.randn(8, D, D, D, C, device="cuda").bfloat16()
    y = torch.empty_like(x)
    A, S = torch.rand(8, C, device="cuda") + 0.5, torch.randn(8, C, device="cuda")
    slope = torch.full((1,), 0.01, device="cuda")
    gb = 2 * x.numel() * 2 / 1e9
    ms = timeit(lambda: ops.affine_act(x, A, S, slope, L.ACT_LEAKY, out=y))
    ms_copy = timeit(lambda: y.copy_(x))
    ms_torch = timeit(lambda: torch.add(x, 1.0, out=y))
    print(f"C={C} D={D}: affine_act {ms:.4f} ms {gb / ms * 1e3:.0f} GB/s | torch copy {gb / ms_copy * 1e3:.0f} GB/s | torch add {gb / ms_torch * 1e3:.0f} GB/s", flush=True)
