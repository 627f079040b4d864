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
store = {}
order = []
def mk(name):
    def hook(mod, inp, out):
        t = out
        if isinstance(t, ops.Deferred): t = t.raw
        if isinstance(t, (tuple, list)): t = t[0]
        if torch.is_tensor(t):
            store[name] = t.detach().clone()
            if name not in order: order.append(name)
    return hook
for n, mod in m.named_modules():
    if n: mod.register_forward_hook(mk(n))

from coma_unet_b200.graph import GraphedTrainStep
opt = GraphedTrainStep.make_optimizer(m, 1e-3)
def fwd(inp, backward):
    m._prompt_use_override = (True, True)
    pred, proj, final = m(inp.mri, inp.covars, roi_pred_dicts=inp.lut, sample_roi_mask=inp.roi)
    m._prompt_use_override = None
    z = torch.zeros(final.size(), device=DEV)
    loss, g, _, _ = crit(pred, inp.tau, inp.roi, (final, z, z), (proj[-1], inp.covars[:, -1]))
    store["<pred>"] = pred.detach().clone(); store["<loss>"] = loss.detach().clone()
    if backward:
        if backward == 2: opt.zero_grad(set_to_none=True)
        else:
            for p in m.parameters(): p.grad = None
        loss.backward()
        if backward == 2: opt.step()
    return loss

mri, tau, roi, covars, dicts = common.synthetic_batch(2, case["shape"], 92)
covars[0, 0, 0], covars[1, 0, 0] = 1.0, 0.0
mri, tau, roi = mri.to(DEV), tau.to(DEV), roi.to(DEV)
inp = _StaticInputs(m, mri, roi, tau, covars)
s = torch.cuda.Stream()
for backward in (False, True, 2):
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        inp.fill(m, mri, roi, tau, covars, dicts)
        for _ in range(2):
            fwd(inp, backward)
        torch.cuda.synchronize()
        saved = {k: v.detach().clone() for k, v in m.state_dict().items()}
        osaved = [{k: (v.clone() if torch.is_tensor(v) else v) for k, v in opt.state[p].items()} for p in m.parameters() if p in opt.state]
        fwd(inp, backward)
        torch.cuda.synchronize()
        eager = {k: v.clone() for k, v in store.items()}
        if backward == 2:
            with torch.enable_grad():
                m._prompt_use_override = (True, True)
                m(inp.mri, inp.covars, roi_pred_dicts=inp.lut, sample_roi_mask=inp.roi)
                m._prompt_use_override = None
            torch.cuda.synchronize()
            eager2 = {k: v.clone() for k, v in store.items()}
        with torch.no_grad():
            for k, v in m.state_dict().items(): v.copy_(saved[k])
            i = 0
            for p in m.parameters():
                if p in opt.state:
                    for k, v in opt.state[p].items():
                        if torch.is_tensor(v): v.copy_(osaved[i][k])
                    i += 1
        ops.invalidate_weight_caches(m)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s, capture_error_mode="thread_local"):
            fwd(inp, backward)
        graph_store = dict(store)
        g.replay()
        torch.cuda.synchronize()
    print("backward in graph:", backward, "loss eager", float(eager["<loss>"]), "graph", float(graph_store["<loss>"]))
    if backward == 2:
        n = "model.0.conv.0"
        print("   vs eager forward with the POST-step weights:", float((eager2[n].float() - graph_store[n].float()).abs().max()))
    bad = 0
    for n in order + ["<pred>"]:
        d = float((eager[n].float() - graph_store[n].float()).abs().max())
        if d > 0:
            print("   first mismatches:", n, d, tuple(eager[n].shape))
            bad += 1
            if bad > 6: break
