"""Poison every torch.empty / empty_like made by the product package with NaN and check that a training step still gives the same
loss and gradients: any kernel that READS memory it was supposed to fully overwrite (or that accumulates into a buffer nobody
zeroed) shows up as NaN / a changed result.  Such reads are invisible in eager runs (the caching allocator hands back blocks that
hold the previous step's values) and break CUDA-graph replays, whose private pool lays memory out differently."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import coma_unet_b200 as cu
from coma_unet_b200 import ops, model as cmodel, blocks, cond_conv, criterions
from tests.golden import common
import torch.nn as nn
DEV = "cuda"
real_empty, real_empty_like, real_new_empty = torch.empty, torch.empty_like, torch.Tensor.new_empty
POISON = {"on": False, "log": []}
def poison(t):
    if POISON["on"] and t.is_cuda and t.is_floating_point() and t.numel():
        t.fill_(float("nan"))
    return t
class TorchProxy:
    def __getattr__(self, n): return getattr(torch, n)
    def empty(self, *a, **k): return poison(real_empty(*a, **k))
    def empty_like(self, *a, **k): return poison(real_empty_like(*a, **k))
for mod in (ops, cmodel, blocks, cond_conv, criterions):
    mod.torch = TorchProxy()
def new_empty(self, *a, **k): return poison(real_new_empty(self, *a, **k))
torch.Tensor.new_empty = new_empty

def run(dtype, shape, channels, poison_on, train=True):
    m = cu.ContrastiveAttentionUNET_DP(3, 1, 1, channels, [2] * 5, latent_spaces=[2048] * 5, conditional=True, prompt_shape=shape, compute_dtype=dtype)
    m.set_save_attn(None)
    common.fill_deterministic(m, 9).to(DEV)
    gen = cu.RoiMSE(real_empty(36).fill_(225.0), common.ROI_INDICES, voxel_wise=False)
    crit = cu.GenerativeContrastiveLoss(cu.RnCLoss(), gen, nn.TripletMarginLoss(1), 0., 1.)
    crit.gen_loss.batch_reduction = None
    mri, tau, roi, covars, dicts = common.synthetic_batch(2, shape, 92)
    covars[0, 0, 0], covars[1, 0, 0] = 1.0, 0.0
    mri, tau, roi = mri.to(DEV), tau.to(DEV), roi.to(DEV)
    m.train(train); m.set_training(train)
    POISON["on"] = poison_on
    if train:
        pred, proj, final = m(mri, covars, roi_pred_dicts=dicts, sample_roi_mask=roi)
        z = torch.zeros(final.size(), device=DEV)
        loss, g, _, _ = crit(pred, tau, roi, (final, z, z), (proj[-1], covars[:, -1].float().to(DEV)))
        loss.backward()
        POISON["on"] = False
        return float(loss), {k: p.grad.detach().float().clone() for k, p in m.named_parameters() if p.grad is not None}
    with torch.no_grad():
        pred = m(mri, covars, roi_pred_dicts=dicts, sample_roi_mask=roi)
    POISON["on"] = False
    return float(pred.float().abs().sum()), {"pred": pred.float().clone()}

for dtype, shape, ch in ((torch.float32, (32, 32, 32), [8, 16, 32, 64, 128]), (torch.bfloat16, (64, 64, 64), [32, 64, 128, 256, 512])):
    for train in (True, False):
        l0, g0 = run(dtype, shape, ch, False, train)
        l1, g1 = run(dtype, shape, ch, True, train)
        bad = [k for k in g0 if not torch.isfinite(g1[k]).all() or float((g1[k] - g0[k]).abs().max()) > 1e-2 * float(g0[k].abs().max() + 1e-20)]
        print(dtype, "train" if train else "eval", "loss/sum clean", l0, "poisoned", l1, "tensors that changed:", len(bad), bad[:10], flush=True)
