"""Which lines of a training step make the host wait for the GPU?  One step under torch.cuda.set_sync_debug_mode("warn"),
once with the covariates on the device (bench.py's device-timed loop) and once with host covariates (the e2e loop)."""
import os
import sys
import warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from coma_unet_b200.parallel import DataParallelEngine

dev = torch.device("cuda", 0)
model = bench.build_model(dev)
mri, tau, roi, covars, dicts = bench.make_batch(2, 1234, device=dev)
model.train(True)
crit = bench.build_criterion()
engine = DataParallelEngine(model, world_size=1)
opt = torch.optim.AdamW(model.parameters(), 1e-3)


def step(c):
    opt.zero_grad(set_to_none=True)
    pred, proj, final = model(mri, c, roi_pred_dicts=dicts, sample_roi_mask=roi)
    feats, labels = engine.gather_rnc(proj[-1], c[:, -1].float().to(dev, non_blocking=True))
    z = torch.zeros(final.size(), device=dev)
    loss, gen, _, _ = crit(pred, tau, roi, (final, z, z), (feats, labels))
    loss.backward()
    engine.finish()
    opt.step()


for c in (covars, covars.cpu().pin_memory()):
    for _ in range(3):
        step(c)
    torch.cuda.synchronize()
    print("==== covariates on", c.device, flush=True)
    torch.cuda.set_sync_debug_mode("warn")
    with warnings.catch_warnings():
        warnings.simplefilter("always")
        step(c)
    torch.cuda.set_sync_debug_mode("default")
    torch.cuda.synchronize()
