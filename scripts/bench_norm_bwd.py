"""Time coma_norm_film_act_bwd (reduce sweep + finalize + apply sweep) on the hot-path shapes of a batch-4 training step.
COMA_DISABLE_NORM_BULK=1 selects the register-staged sweeps instead of the bulk-copy streaming ones (A/B in two runs)."""
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


print("bulk sweeps:", "off" if os.environ.get("COMA_DISABLE_NORM_BULK") else "on")
for C, D, mode in ((32, 128, L.NORM_BATCH), (16, 128, L.NORM_INSTANCE), (64, 64, L.NORM_BATCH), (128, 32, L.NORM_BATCH), (256, 16, L.NORM_BATCH)):
    B = 4
    x = torch.randn(B, D, D, D, C, device="cuda").bfloat16().requires_grad_(True)
    g = (torch.rand(B, C, device="cuda") + 0.5).requires_grad_(True)
    h = torch.randn(B, C, device="cuda").requires_grad_(True)
    dy = torch.randn(B, D, D, D, C, device="cuda").bfloat16()
    y = ops.norm_act(x, g, h, None, ops.NormCfg(mode=mode, act=L.ACT_RELU))
    gb = 5 * x.numel() * 2 / 1e9
    ms = timeit(lambda: torch.autograd.grad(y, (x, g, h), dy, retain_graph=True))
    print(f"C={C} D={D} mode={mode}: norm bwd {ms:.4f} ms, {gb / ms * 1e3:.0f} GB/s algorithmic (5 tensor passes)", flush=True)
