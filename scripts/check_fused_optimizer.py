"""Two training steps with the single-tensor, foreach and fused AdamW implementations: losses and weights must agree (they did not
while the packed-weight cache was keyed by the version counter alone: fused optimizers do not bump it)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import coma_unet_b200 as cu
from tests.golden import common
import torch.nn as nn
DEV = "cuda"
case = {"channels": [8, 16, 32, 64, 128], "shape": [32, 32, 32], "batch": 2, "seed": 9}
mri, tau, roi, covars, dicts = common.synthetic_batch(2, case["shape"], 92)
covars[0, 0, 0], covars[1, 0, 0] = 1.0, 0.0
mri, tau, roi = mri.to(DEV), tau.to(DEV), roi.to(DEV)

def run(kind, steps=2):
    m = cu.ContrastiveAttentionUNET_DP(3, 1, 1, case["channels"], [2] * 5, latent_spaces=[2048] * 5, conditional=True, prompt_shape=tuple(case["shape"]), compute_dtype=torch.float32)
    m.set_save_attn(None)
    common.fill_deterministic(m, 9).to(DEV)
    m.train(True)
    gen = cu.RoiMSE(torch.tensor([225.0] * 36), common.ROI_INDICES, voxel_wise=False)
    crit = cu.GenerativeContrastiveLoss(cu.RnCLoss(), gen, nn.TripletMarginLoss(1), 0., 1.)
    crit.gen_loss.batch_reduction = None
    kw = {"single": dict(foreach=False, fused=False), "foreach": dict(foreach=True), "fused": dict(fused=True),
          "fused_capturable": dict(fused=True, capturable=True)}[kind]
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3, **kw)
    losses = []
    for it in range(steps):
        opt.zero_grad(set_to_none=True)
        pred, proj, final = m(mri, covars, roi_pred_dicts=dicts, sample_roi_mask=roi)
        z = torch.zeros(final.size(), device=DEV)
        loss, g, _, _ = crit(pred, tau, roi, (final, z, z), (proj[-1], covars[:, -1].float().to(DEV)))
        loss.backward()
        if it == 0:
            info = {k: (tuple(p.grad.shape), tuple(p.grad.stride()), p.grad.is_contiguous(), p.grad.data_ptr() % 16) for k, p in m.named_parameters() if p.grad is not None}
        opt.step()
        losses.append(float(loss))
    torch.cuda.synchronize()
    return losses, {k: v.detach().clone() for k, v in m.named_parameters()}, info

ref_l, ref_w, info = run("single")
print("single", ref_l)
for kind in ("foreach", "fused", "fused_capturable"):
    l, w, _ = run(kind)
    bad = sorted(((float((w[k] - ref_w[k]).abs().max()), k) for k in w), reverse=True)
    print(kind, l, "params differing > 1e-5:", sum(1 for d, k in bad if d > 1e-5), [(k, f"{d:.2e}", info.get(k)) for d, k in bad[:12] if d > 1e-5], flush=True)
