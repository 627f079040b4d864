import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import coma_unet_b200 as cu
from coma_unet_b200.graph import GraphedTrainStep
from coma_unet_b200.parallel import DataParallelEngine
from tests.golden import common
import torch.nn as nn
DEV = "cuda"
case = {"channels": [8, 16, 32, 64, 128], "shape": [32, 32, 32], "batch": 2, "seed": 9}

def build():
    m = cu.ContrastiveAttentionUNET_DP(3, 1, 1, case["channels"], [2] * 5, latent_spaces=[2048] * 5, conditional=True, prompt_shape=tuple(case["shape"]), compute_dtype=torch.float32)
    m.set_save_attn(None)
    return common.fill_deterministic(m, 9).to(DEV)

def criterion():
    gen = cu.RoiMSE(torch.tensor([225.0] * 36), common.ROI_INDICES, voxel_wise=False)
    c = cu.GenerativeContrastiveLoss(cu.RnCLoss(), gen, nn.TripletMarginLoss(1), 0., 1.)
    c.gen_loss.batch_reduction = None
    return c

batches = []
for i in range(5):
    mri, tau, roi, covars, dicts = common.synthetic_batch(2, case["shape"], 90 + i)
    covars[0, 0, 0], covars[1, 0, 0] = 1.0, 0.0
    batches.append((mri.to(DEV), tau.to(DEV), roi.to(DEV), covars, dicts))

names = ["model.0.conv.0.conv.weight", "model.1.merge.conv.weight", "general_dynamic_prompt", "final_pred_head.conv.weight", "model.0.conv.0.adn.N.running_mean"]
def snap(m, tag, loss):
    sd = m.state_dict()
    gr = dict(m.named_parameters())
    print(tag, f"loss {float(loss):.5f}", " ".join(f"{n[-22:]}:{float(sd[n].double().norm()):.6f}/g{(float(gr[n].grad.double().norm()) if n in gr and gr[n].grad is not None else -1):.4e}" for n in names), flush=True)

for mode in ("eager", "graph", "runner-eager"):
    m = build(); m.train(True)
    crit = criterion(); eng = DataParallelEngine(m, world_size=1)
    opt = GraphedTrainStep.make_optimizer(m, 1e-3)
    runner = GraphedTrainStep(m, crit, opt, eng, warmup=2 if mode == "graph" else 100)
    for k, (mri, tau, roi, covars, dicts) in enumerate(batches):
        if mode == "eager":
            opt.zero_grad(set_to_none=True)
            pred, proj, final = m(mri, covars, roi_pred_dicts=dicts, sample_roi_mask=roi)
            z = torch.zeros(final.size(), device=DEV)
            loss, _, _, _ = crit(pred, tau, roi, (final, z, z), (proj[-1], covars[:, -1].float().to(DEV)))
            loss.backward(); opt.step()
        else:
            loss = runner(mri, tau, roi, covars, dicts)
        torch.cuda.synchronize()
        snap(m, f"{mode} step {k}", loss)
