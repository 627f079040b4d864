"""Where does the host time of a training step go?  (cProfile over 3 steps, GPU idle time = step - kernel sum)"""
import cProfile
import os
import pstats
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
model = bench.build_model(dev)
mri, tau, roi, covars, dicts = bench.make_batch(B, 1234, device=dev)
model.train(True)
crit = bench.build_criterion()
opt = torch.optim.AdamW(model.parameters(), 1e-3)


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
import time
t0 = time.perf_counter()
for _ in range(3):
    step()
t_launch = time.perf_counter() - t0
torch.cuda.synchronize()
t_total = time.perf_counter() - t0
st = torch.cuda.memory_stats()
print(f"max allocated {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB, reserved {torch.cuda.memory_reserved() / 2**30:.1f} GiB, "
      f"alloc retries {st.get('num_alloc_retries')}, cudaMalloc calls {st.get('segment.all.allocated')}")
print(f"B={B}: host launch time {t_launch / 3 * 1e3:.1f} ms/step, wall {t_total / 3 * 1e3:.1f} ms/step")
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
