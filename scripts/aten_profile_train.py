"""Which ATen ops (with input shapes) account for the GPU time of a training step outside the C-ABI kernels?
torch.profiler over 2 eager steps, grouped by (op, input shapes), top-level ops only."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

dev = torch.device("cuda", 0)
shape = (128, 128, 128)
model = bench.build_model(dev, shape)
mri, tau, roi, covars, dicts = bench.make_batch(4, 1234, shape, device=dev)
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
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
    for _ in range(N):
        step()
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages(group_by_input_shape=True):
    t = getattr(e, "self_device_time_total", 0) or 0
    if t > 0 and e.key.startswith("aten::"):
        rows.append((t / N / 1e3, e.count // N, e.key, str(e.input_shapes)[:110]))
rows.sort(reverse=True)
print(f"ATen self device time per step: {sum(r[0] for r in rows):.3f} ms")
for ms, n, name, shapes in rows[:45]:
    print(f"{ms:7.3f} ms {n:4d}x {name:28s} {shapes}")
