"""One step of the benchmarked workload between cudaProfilerStart/Stop, for ncu launch lists and --set full captures:

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv \
        python scripts/one_step.py --mode train            # every kernel of one batch-4 training step (eager launches)
    python scripts/one_step.py --mode infer [--batch 8] [--steps 1]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="train", choices=["train", "infer"])
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--warmup", type=int, default=3)
args = ap.parse_args()
dev = torch.device("cuda", 0)
shape = (128, 128, 128)
train = args.mode == "train"
B = args.batch or (4 if train else 8)
model = bench.build_model(dev, shape)
mri, tau, roi, covars, dicts = bench.make_batch(B, 1234, shape, device=dev)
if train:
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
else:
    model.eval()
    model.set_training(False)

    def step():
        with torch.no_grad():
            model(mri, covars, roi_pred_dicts=dicts, sample_roi_mask=roi)

for _ in range(args.warmup):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(args.steps):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
