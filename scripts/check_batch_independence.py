"""Eval-mode forward of a batch of two against the same samples run alone, fp32 and bf16, 64^3 .. 256^3, with the first layers whose
outputs differ: bf16 results depend on the batch size at the level of bf16 rounding noise."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import coma_unet_b200 as cu
from tests.golden import common
DEV = "cuda"
def rms(a, b): return float((a.float() - b.float()).pow(2).mean().sqrt() / b.float().pow(2).mean().sqrt())
for shape, ch, dtype in (((64,)*3, [32,64,128,256,512], torch.float32), ((64,)*3, [32,64,128,256,512], torch.bfloat16), ((128,)*3, [32,64,128,256,512], torch.bfloat16), ((256,)*3, [32,64,128,256,512], torch.bfloat16)):
    m = cu.ContrastiveAttentionUNET_DP(3, 1, 1, ch, [2]*5, latent_spaces=[2048]*5, conditional=True, prompt_shape=shape, compute_dtype=dtype)
    m.set_save_attn(None)
    common.fill_deterministic(m, 71).to(DEV).eval(); m.set_training(False)
    mri, tau, roi, cov, dicts = common.synthetic_batch(2, shape, 71)
    mri, roi = mri.to(DEV), roi.to(DEV)
    acts = {}
    def mk(n):
        def hook(mod, i, o):
            t = o
            if hasattr(t, "materialize"): t = t.materialize()
            if isinstance(t, (tuple, list)): t = t[0]
            if torch.is_tensor(t) and t.dim() == 5: acts.setdefault(n, t.detach().float().clone())
        return hook
    hs = [mod.register_forward_hook(mk(n)) for n, mod in m.named_modules() if n.count(".") <= 3 and n]
    with torch.no_grad():
        pb = m(mri, cov, roi_pred_dicts=dicts, sample_roi_mask=roi); ab = dict(acts); acts.clear()
        p0 = m(mri[:1], cov[:1], roi_pred_dicts=dicts[:1], sample_roi_mask=roi[:1]); a0 = dict(acts); acts.clear()
        pb2 = m(mri, cov, roi_pred_dicts=dicts, sample_roi_mask=roi)
    for h in hs: h.remove()
    print(shape, dtype, "pred batch vs single rms", rms(pb[:1], p0), "max", float((pb[:1]-p0).abs().max()/p0.abs().max()), "| batch rerun identical:", bool(torch.equal(pb, pb2)))
    n_shown = 0
    for n in ab:
        if n in a0 and ab[n][:1].shape == a0[n].shape:
            r = rms(ab[n][:1], a0[n])
            if r > 1e-4 and n_shown < 6:
                print("    first layers that differ:", n, f"{r:.2e}"); n_shown += 1
    del m
    torch.cuda.empty_cache()
