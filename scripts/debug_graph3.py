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
m = cu.ContrastiveAttentionUNET_DP(3, 1, 1, case["channels"], [2] * 5, latent_spaces=[2048] * 5, conditional=True, prompt_shape=tuple(case["shape"]), compute_dtype=torch.float32)
m.set_save_attn(None)
common.fill_deterministic(m, 9).to(DEV)
m.train(True)
gen = cu.RoiMSE(torch.tensor([225.0] * 36), common.ROI_INDICES, voxel_wise=False)
crit = cu.GenerativeContrastiveLoss(cu.RnCLoss(), gen, nn.TripletMarginLoss(1), 0., 1.)
crit.gen_loss.batch_reduction = None
eng = DataParallelEngine(m, world_size=1)
opt = GraphedTrainStep.make_optimizer(m, 1e-3)
runner = GraphedTrainStep(m, crit, opt, eng, warmup=2)
orig = runner._capture
def snap():
    torch.cuda.synchronize()
    return {k: v.detach().clone() for k, v in m.state_dict().items()}
def cap(key):
    before = snap()
    st0 = {id(p): {k: (v.clone() if torch.is_tensor(v) else v) for k, v in opt.state[p].items()} for p in m.parameters() if p in opt.state}
    orig(key)
    after = snap()
    changed = [k for k in before if not torch.equal(before[k], after[k])]
    print("changed by the CAPTURE itself:", len(changed), changed[:8])
    nst = 0
    for p in m.parameters():
        if p in opt.state:
            for k, v in opt.state[p].items():
                if torch.is_tensor(v) and not torch.equal(v, st0[id(p)][k]):
                    nst += 1
    print("optimizer state tensors changed by the capture:", nst)
runner._capture = cap
for i in range(4):
    mri, tau, roi, covars, dicts = common.synthetic_batch(2, case["shape"], 90 + i)
    covars[0, 0, 0], covars[1, 0, 0] = 1.0, 0.0
    loss = runner(mri.to(DEV), tau.to(DEV), roi.to(DEV), covars, dicts)
    torch.cuda.synchronize()
    print(i, float(loss), "launches/graph", runner.launches)
