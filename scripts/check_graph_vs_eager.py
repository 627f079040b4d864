"""One optimizer step replayed from a CUDA graph against the same step run eagerly, for several optimizers (none, SGD, AdamW
foreach / fused): loss, first conv output and every weight after the step must agree.  This is how the stale packed-weight cache
behind fused optimizers was found (DESIGN.md section 6, round-2 findings)."""
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
mri, tau, roi, covars, dicts = common.synthetic_batch(2, case["shape"], 92)
covars[0, 0, 0], covars[1, 0, 0] = 1.0, 0.0
mri, tau, roi = mri.to(DEV), tau.to(DEV), roi.to(DEV)

def experiment(kind, order="in"):
    m = cu.ContrastiveAttentionUNET_DP(3, 1, 1, case["channels"], [2] * 5, latent_spaces=[2048] * 5, conditional=True, prompt_shape=tuple(case["shape"]), compute_dtype=torch.float32)
    m.set_save_attn(None)
    common.fill_deterministic(m, 9).to(DEV)
    m.train(True)
    gen = cu.RoiMSE(torch.tensor([225.0] * 36), common.ROI_INDICES, voxel_wise=False)
    crit = cu.GenerativeContrastiveLoss(cu.RnCLoss(), gen, nn.TripletMarginLoss(1), 0., 1.)
    crit.gen_loss.batch_reduction = None
    lr = torch.tensor(1e-3, device=DEV)
    if kind == "sgd": opt = torch.optim.SGD(m.parameters(), lr=1e-3)
    elif kind == "adamw_foreach": opt = torch.optim.AdamW(m.parameters(), lr=lr, foreach=True, capturable=True)
    elif kind == "adamw_fused": opt = torch.optim.AdamW(m.parameters(), lr=lr, fused=True, capturable=True)
    elif kind == "adamw_fused_floatlr": opt = torch.optim.AdamW(m.parameters(), lr=1e-3, fused=True, capturable=True)
    else: opt = None
    first = {}
    m.model[0].conv[0].register_forward_hook(lambda mod, i, o: first.__setitem__("v", o.detach().clone()))
    inp = _StaticInputs(m, mri, roi, tau, covars)
    def step():
        if order == "in" and opt is not None: opt.zero_grad(set_to_none=True)
        else:
            for p in m.parameters(): p.grad = None
        m._prompt_use_override = (True, True)
        pred, proj, final = m(inp.mri, inp.covars, roi_pred_dicts=inp.lut, sample_roi_mask=inp.roi)
        m._prompt_use_override = None
        z = torch.zeros(final.size(), device=DEV)
        loss, g, _, _ = crit(pred, inp.tau, inp.roi, (final, z, z), (proj[-1], inp.covars[:, -1]))
        loss.backward()
        if opt is not None: opt.step()
        return loss.detach()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        inp.fill(m, mri, roi, tau, covars, dicts)
        for _ in range(2): step()
        torch.cuda.synchronize()
        saved = {k: v.detach().clone() for k, v in m.state_dict().items()}
        osaved = None if opt is None else [{k: (v.clone() if torch.is_tensor(v) else v) for k, v in opt.state[p].items()} for p in m.parameters() if p in opt.state]
        le = float(step()); fe = first["v"].clone()
        torch.cuda.synchronize()
        we = {k: v.detach().clone() for k, v in m.state_dict().items()}
        with torch.no_grad():
            for k, v in m.state_dict().items(): v.copy_(saved[k])
            if opt is not None:
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
            lg = step()
        fg = first["v"]
        g.replay()
        torch.cuda.synchronize()
        wg = {k: v.detach().clone() for k, v in m.state_dict().items()}
    worst = max(((float((wg[k].double() - we[k].double()).abs().max()), k) for k in we if we[k].dtype.is_floating_point))
    print(f"{kind:22s} loss eager {le:.5f} graph {float(lg):.5f} | first conv diff {float((fe - fg).abs().max()):.3e} | worst weight diff after the step {worst[0]:.3e} ({worst[1]})", flush=True)

for kind in ("none", "sgd", "adamw_foreach", "adamw_fused", "adamw_fused_floatlr"):
    experiment(kind)
