"""Diagnostic (GPU): error of the product model and of the oracle-on-GPU against the CPU fixtures / oracle."""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import coma_unet_b200 as cu
from oracle import criterions as ocrit
from oracle import model as omodel
from tests.golden import check, common

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
DATA, META = check.load()
DEV = "cuda"


def crit(mod):
    gen = mod.RoiMSE(torch.tensor([225.0] * 36), common.ROI_INDICES, voxel_wise=False)
    c = mod.GenerativeContrastiveLoss(mod.RnCLoss(), gen, nn.TripletMarginLoss(1), 0., 1.)
    c.gen_loss.batch_reduction = None
    return c


def run(model, mod, case):
    mri, tau, roi, covars, dicts = common.synthetic_batch(case["batch"], case["shape"], case["seed"])
    mri, tau, roi = mri.to(DEV), tau.to(DEV), roi.to(DEV)
    model.train(True)
    pred, proj, final = model(mri, covars, roi_pred_dicts=dicts, sample_roi_mask=roi)
    z = torch.zeros(final.size(), device=DEV)
    loss, gen, _, _ = crit(mod)(pred, tau, roi, (final, z, z), (proj[-1], covars[:, -1].float().to(DEV)))
    loss.backward()
    return pred.detach(), proj, float(loss.detach()), {k: p.grad for k, p in model.named_parameters() if p.grad is not None}


for name in ["train32", "train32_b1"]:
    case = META[name]
    kw = dict(latent_spaces=[2048] * 5, conditional=True, prompt_shape=tuple(case["shape"]))
    models = {
        "oracle_gpu": (common.fill_deterministic(omodel.ContrastiveAttentionUNET_DP(3, 1, 1, case["channels"], [2] * 5, **kw), case["seed"]).to(DEV), ocrit),
        "fp32": (common.fill_deterministic(cu.ContrastiveAttentionUNET_DP(3, 1, 1, case["channels"], [2] * 5, compute_dtype=torch.float32, **kw), case["seed"]).to(DEV), cu),
        "bf16": (common.fill_deterministic(cu.ContrastiveAttentionUNET_DP(3, 1, 1, case["channels"], [2] * 5, compute_dtype=torch.bfloat16, **kw), case["seed"]).to(DEV), cu),
    }
    res = {k: run(m, mod, case) for k, (m, mod) in models.items()}
    print(f"== {name} (fixture loss {DATA[name + '/loss'][0]:.6f})")
    for k, (pred, proj, loss, grads) in res.items():
        got, want = check.sampled(DATA, f"{name}/pred", pred)
        print(f"  {k:10s} loss {loss:.6f} pred rel(floor1e-3) {check.rel_err(got, want):.2e} scaled {check.scaled_err(got, want):.2e}")
        for gk in sorted({g.split('/grad/')[1].rsplit('/', 1)[0] for g in DATA.files if g.startswith(f'{name}/grad/')}):
            a, b = check.sampled(DATA, f"{name}/grad/{gk}", grads[gk])
            cos = float((a * b).sum() / (np.sqrt((a * a).sum() * (b * b).sum()) + 1e-30))
            print(f"      grad {gk:70s} scaled {check.scaled_err(a, b):.2e} cos {cos:.5f}")

# ---- calibration: what does stock torch bf16 autocast lose on the same model? and where do we lose it? ----
name = "train32"
case = META[name]
kw = dict(latent_spaces=[2048] * 5, conditional=True, prompt_shape=tuple(case["shape"]))
mri, tau, roi, covars, dicts = common.synthetic_batch(case["batch"], case["shape"], case["seed"])
mri, tau, roi = mri.to(DEV), tau.to(DEV), roi.to(DEV)
o = common.fill_deterministic(omodel.ContrastiveAttentionUNET_DP(3, 1, 1, case["channels"], [2] * 5, **kw), case["seed"]).to(DEV).eval()
o.set_training(False)
with torch.no_grad():
    ref = o(mri, covars, roi_pred_dicts=dicts, sample_roi_mask=roi)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ac = o(mri, covars, roi_pred_dicts=dicts, sample_roi_mask=roi).float()
    rb, re, rd = omodel.ObservableAttentionUnet.forward(o, mri, covars)
print(f"eval: torch autocast bf16 vs fp32 oracle: scaled {check.scaled_err(ac.cpu().numpy(), ref.cpu().numpy()):.2e}")
for dt in (torch.float32, torch.bfloat16):
    m = common.fill_deterministic(cu.ContrastiveAttentionUNET_DP(3, 1, 1, case["channels"], [2] * 5, compute_dtype=dt, **kw), case["seed"]).to(DEV).eval()
    m.set_training(False)
    with torch.no_grad():
        out = m(mri, covars, roi_pred_dicts=dicts, sample_roi_mask=roi)
        pb, pe, pd = cu.ObservableAttentionUnet.forward(m, mri, covars)
    print(f"eval: product {dt} vs fp32 oracle: final scaled {check.scaled_err(out.cpu().numpy(), ref.cpu().numpy()):.2e}"
          f"  backbone-out {check.scaled_err(pb.cpu().numpy(), rb.cpu().numpy()):.2e}")
    for i, (a, b) in enumerate(zip(pe, re)):
        print(f"      enc{i} scaled {check.scaled_err(a.cpu().numpy(), b.cpu().numpy()):.2e}")
    for i, (a, b) in enumerate(zip(pd, rd)):
        print(f"      dec{i} scaled {check.scaled_err(a.cpu().numpy(), b.cpu().numpy()):.2e}")

# ---- calibration of bf16 GRADIENTS: torch autocast on the oracle vs fp32 oracle (train step) ----
o32 = common.fill_deterministic(omodel.ContrastiveAttentionUNET_DP(3, 1, 1, case["channels"], [2] * 5, **kw), case["seed"]).to(DEV)
oac = common.fill_deterministic(omodel.ContrastiveAttentionUNET_DP(3, 1, 1, case["channels"], [2] * 5, **kw), case["seed"]).to(DEV)
g32 = run(o32, ocrit, case)[3]


def run_autocast(model):
    mri, tau, roi, covars, dicts = common.synthetic_batch(case["batch"], case["shape"], case["seed"])
    mri, tau, roi = mri.to(DEV), tau.to(DEV), roi.to(DEV)
    model.train(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        pred, proj, final = model(mri, covars, roi_pred_dicts=dicts, sample_roi_mask=roi)
    z = torch.zeros(final.size(), device=DEV)
    loss, gen, _, _ = crit(ocrit)(pred.float(), tau, roi, (final.float(), z, z), (proj[-1].float(), covars[:, -1].float().to(DEV)))
    loss.backward()
    return {k: p.grad for k, p in model.named_parameters() if p.grad is not None}


gac = run_autocast(oac)
mb = common.fill_deterministic(cu.ContrastiveAttentionUNET_DP(3, 1, 1, case["channels"], [2] * 5, compute_dtype=torch.bfloat16, **kw), case["seed"]).to(DEV)
gb = run(mb, cu, case)[3]
print("train-step gradient cosine vs fp32 oracle:   torch-autocast-bf16   this-repo-bf16")
for k in ["model.0.conv.0.conv.weight", "model.0.conv.0.film.2.weight", "model.1.submodule.0.conv.0.conv.weight",
          "model.1.submodule.0.conv.1.conv.weight", "model.1.merge.conv.weight", "model.1.upconv.up.conv.weight",
          "model.1.attention.W_g.0.conv.weight", "final_pred_head.conv.weight", "general_dynamic_prompt"]:
    def cosine(a, b):
        a, b = a.float().flatten().double(), b.float().flatten().double()
        return float((a @ b) / (a.norm() * b.norm() + 1e-30))
    print(f"  {k:60s} {cosine(gac[k], g32[k]):.4f}   {cosine(gb[k], g32[k]):.4f}")
