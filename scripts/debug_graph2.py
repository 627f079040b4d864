import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import coma_unet_b200 as cu
from coma_unet_b200 import ops
from coma_unet_b200.graph import _StaticInputs
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
batches = []
for i in range(3):
    mri, tau, roi, covars, dicts = common.synthetic_batch(2, case["shape"], 90 + i)
    covars[0, 0, 0], covars[1, 0, 0] = 1.0, 0.0
    batches.append((mri.to(DEV), tau.to(DEV), roi.to(DEV), covars, dicts))

def fwd(mri, tau, roi, covars, lut):
    m._prompt_use_override = (True, True)
    pred, proj, final = m(mri, covars, roi_pred_dicts=lut, sample_roi_mask=roi)
    m._prompt_use_override = None
    z = torch.zeros(final.size(), device=DEV)
    loss, g, _, _ = crit(pred, tau, roi, (final, z, z), (proj[-1], covars[:, -1]))
    return pred, proj[-1], final, loss, g

inp = None
s = torch.cuda.Stream()
graph = None
for k, (mri, tau, roi, covars, dicts) in enumerate(batches * 2):
    with torch.no_grad():
        pe, fe, fine, le, ge = m(mri, covars, roi_pred_dicts=dicts, sample_roi_mask=roi) + (None, None)
        z = torch.zeros(fine.size(), device=DEV)
        le, ge, _, _ = crit(pe, tau, roi, (fine, z, z), (fe[-1], covars[:, -1].float().to(DEV)))
    if inp is None:
        inp = _StaticInputs(m, mri, roi, tau, covars)
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s), torch.no_grad():
        inp.fill(m, mri, roi, tau, covars, dicts)
        if k < 2:
            out = fwd(inp.mri, inp.tau, inp.roi, inp.covars, inp.lut)
        else:
            if graph is None:
                ops.invalidate_weight_caches(m)
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=s, capture_error_mode="thread_local"):
                    gout = fwd(inp.mri, inp.tau, inp.roi, inp.covars, inp.lut)
            graph.replay()
            out = gout
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    print(k, "graph" if k >= 2 else "eager-static", "loss", float(le), float(out[3]), "pred diff", float((out[0] - pe).abs().max()), "feat diff", float((out[1] - fe[-1]).abs().max()),
          "gen", ge.flatten().tolist(), out[4].flatten().tolist(),
          "inputs ok", bool(torch.equal(inp.mri, mri)), bool(torch.equal(inp.tau, tau)), bool(torch.equal(inp.roi, roi)), float((inp.covars.cpu().double() - covars).abs().max()), flush=True)
