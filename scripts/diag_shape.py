import sys, os, torch
sys.path.insert(0, "/root/repo")
import coma_unet_b200 as cu
from oracle import model as omodel
from tests.golden import check, common
import tests.test_gpu_model as T
shape = [int(v) for v in sys.argv[1].split(",")]
case = {"channels": [32, 64, 128, 256, 512], "shape": shape, "batch": 1, "seed": 61}
o = T.build(case, None, cls=lambda *a, compute_dtype=None, **k: omodel.ContrastiveAttentionUNET_DP(*a, **k)).cpu().eval()
o.set_training(False)
mri, tau, roi, covars, dicts = T.batch(case)
with torch.no_grad():
    want = o(mri.cpu(), covars, roi_pred_dicts=dicts, sample_roi_mask=roi.cpu()).numpy()
for dt in (torch.bfloat16,) + ((torch.float32,) if len(sys.argv) > 2 else ()):
    m = T.build(case, dt).eval(); m.set_training(False)
    with torch.no_grad():
        got = m(mri, covars, roi_pred_dicts=dicts, sample_roi_mask=roi).cpu().numpy()
    import numpy as np
    e = np.abs(got - want); i = np.unravel_index(e.argmax(), e.shape)
    print(shape, dt, "scaled err", check.scaled_err(got, want), "mean abs err", e.mean() / np.abs(want).max(), "argmax", i, got[i], want[i], "PAIR off" if os.environ.get("COMA_DISABLE_PAIR") else "")
