"""Which GPU kernels make up a training step?  torch.profiler (CUPTI) over 2 steps: top kernels by device time, so that the
torch-side kernels between the C-ABI launches (grad accumulation, casts, copies, optimizer) are visible next to ours.

    python scripts/kernel_profile_train.py [batch]
"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
model = bench.build_model(dev)
mri, tau, roi, covars, dicts = bench.make_batch(B, 1234, device=dev)
model.train(True)
crit = bench.build_criterion()
opt = torch.optim.AdamW(model.parameters(), 1e-3, fused=True)


def step():
    opt.zero_grad(set_to_none=True)
    pred, proj, final = model(mri, covars, roi_pred_dicts=dicts, sample_roi_mask=roi)
    z = torch.zeros(final.size(), device=dev)
    loss, gen, _, _ = crit(pred, tau, roi, (final, z, z), (proj[-1], covars[:, -1].float().to(dev)))
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
N = 2
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(N):
        step()
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / N / 1e3, e.count // N) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
rows.sort(key=lambda r: -r[1])
total = sum(r[1] for r in rows)
print(f"B={B}: {total:.2f} ms of GPU kernels per step")
for name, ms, n in rows[:45]:
    print(f"{ms:8.3f} ms {n:5d}x  {name[:150]}")
