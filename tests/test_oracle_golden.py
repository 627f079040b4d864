"""CPU: the oracle restatement against fixtures produced by the reference's own code
(tests/golden/make_golden.py).  fp32 vs fp32, so the tolerance is the TF32/fp32 bound 1e-4."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import criterions as ocrit
from oracle import model as omodel
from tests.golden import check, common

DATA, META = check.load()
DATA2, META2 = check.load2()
TOL = 1e-4


def build(case):
    m = omodel.ContrastiveAttentionUNET_DP(3, 1, 1, case["channels"], [2] * 5, latent_spaces=[2048] * 5,
                                           conditional=True, decoder_ds=False, prompt_shape=tuple(case["shape"]))
    m.set_save_attn(None)
    return common.fill_deterministic(m, case["seed"])


def criterion():
    gen = ocrit.RoiMSE(torch.tensor([225.0] * 36), common.ROI_INDICES, voxel_wise=False)
    crit = ocrit.GenerativeContrastiveLoss(ocrit.RnCLoss(), gen, nn.TripletMarginLoss(1), 0., 1.)
    crit.gen_loss.batch_reduction = None
    return crit


@pytest.mark.parametrize("name", ["train32", "train32_b1"])
def test_train_step_matches_reference(name):
    case = META[name]
    m = build(case)
    assert list(m.state_dict().keys()) == case["state_keys"]
    mri, tau, roi, covars, dicts = common.synthetic_batch(case["batch"], case["shape"], case["seed"])
    m.train(True)
    pred, projected, final_repr = m(mri, covars, roi_pred_dicts=dicts, sample_roi_mask=roi)
    zeros = torch.zeros(final_repr.size())
    loss, gen, ps, ds = criterion()(pred, tau, roi, (final_repr, zeros, zeros), (projected[-1], covars[:, -1].float()))
    loss.backward()
    assert check.rel_err(*check.sampled(DATA, f"{name}/pred", pred)) < TOL
    for i, p in enumerate(projected):
        assert check.rel_err(*check.sampled(DATA, f"{name}/proj{i}", p)) < TOL
    assert check.rel_err(*check.sampled(DATA, f"{name}/final_repr", final_repr)) < TOL
    assert check.rel_err([float(loss.detach()), float(ps), float(ds)], DATA[f"{name}/loss"]) < TOL
    assert check.rel_err(gen.detach().numpy(), DATA[f"{name}/gen"]) < TOL
    params = dict(m.named_parameters())
    assert sorted(k for k, p in params.items() if p.grad is None) == case["no_grad_params"]
    for k in make_probe_list(name):
        assert check.scaled_err(*check.sampled(DATA, f"{name}/grad/{k}", params[k].grad)) < 5e-4, k
    sd = m.state_dict()
    for key in [k for k in DATA.files if k.startswith(f"{name}/buf/")]:
        np.testing.assert_allclose(sd[key.split("/buf/")[1]].numpy(), DATA[key], rtol=1e-4, atol=1e-6)
    m.eval()
    m.set_training(False)
    with torch.no_grad():
        pred_eval = m(mri, covars, roi_pred_dicts=dicts, sample_roi_mask=roi)
    assert check.rel_err(*check.sampled(DATA, f"{name}/pred_eval", pred_eval)) < TOL


def make_probe_list(name, data=DATA):
    return sorted({k.split("/grad/")[1].rsplit("/", 1)[0] for k in data.files if k.startswith(f"{name}/grad/")})


def test_train64_matches_reference():
    """Round-2 fixture: 64^3, batch 2 -- well-conditioned BatchNorm at the bottom level, gradients held to 1e-4."""
    name, case = "train64", META2["train64"]
    m = build(case)
    mri, tau, roi, covars, dicts = common.synthetic_batch(case["batch"], case["shape"], case["seed"])
    m.train(True)
    pred, projected, final_repr = m(mri, covars, roi_pred_dicts=dicts, sample_roi_mask=roi)
    zeros = torch.zeros(final_repr.size())
    loss, gen, ps, ds = criterion()(pred, tau, roi, (final_repr, zeros, zeros), (projected[-1], covars[:, -1].float()))
    loss.backward()
    assert check.rel_err(*check.sampled(DATA2, f"{name}/pred", pred)) < TOL
    assert check.rel_err([float(loss.detach()), float(ps), float(ds)], DATA2[f"{name}/loss"]) < TOL
    params = dict(m.named_parameters())
    assert sorted(k for k, p in params.items() if p.grad is None) == case["no_grad_params"]
    for k in make_probe_list(name, DATA2):
        assert check.scaled_err(*check.sampled(DATA2, f"{name}/grad/{k}", params[k].grad)) < 1e-4, k


def test_eval128_full_width_matches_reference():
    case = META2["eval128_full"]
    m = build(case).eval()
    m.set_training(False)
    mri, tau, roi, covars, dicts = common.synthetic_batch(case["batch"], case["shape"], case["seed"])
    with torch.no_grad():
        pred = m(mri, covars, roi_pred_dicts=dicts, sample_roi_mask=roi)
    assert check.rel_err(*check.sampled(DATA2, "eval128_full/pred_eval", pred)) < TOL


def test_eval128_matches_reference():
    case = META["eval128"]
    m = build(case).eval()
    m.set_training(False)
    mri, tau, roi, covars, dicts = common.synthetic_batch(case["batch"], case["shape"], case["seed"])
    with torch.no_grad():
        pred = m(mri, covars, roi_pred_dicts=dicts, sample_roi_mask=roi)
    assert check.rel_err(*check.sampled(DATA, "eval128/pred_eval", pred)) < TOL


def test_criterions_match_reference():
    case = META["crit"]
    B, shape, seed = case["batch"], tuple(case["shape"]), case["seed"]
    mri, tau, roi, covars, _ = common.synthetic_batch(B, shape, seed)
    pred = (tau + 0.2 * common._randn(tau.shape, seed, "pred")).requires_grad_(True)
    feats = common._randn((B, 512), seed, "feats").abs().requires_grad_(True)
    final = common._randn((B, 1, 1, 1, 2048), seed, "final").requires_grad_(True)
    zeros = torch.zeros(final.size())
    loss, gen, ps, ds = criterion()(pred, tau, roi, (final, zeros, zeros), (feats, covars[:, -1].float()))
    loss.backward()
    assert check.rel_err([float(loss.detach()), float(ps), float(ds)], DATA["crit/loss"]) < 1e-5
    assert check.rel_err(gen.detach().numpy(), DATA["crit/gen"]) < 1e-5
    assert check.scaled_err(pred.grad.numpy(), DATA["crit/dpred"]) < 1e-5
    assert check.scaled_err(feats.grad.numpy(), DATA["crit/dfeats"]) < 1e-5
    f2 = common._randn((6, 64), seed, "f2").requires_grad_(True)
    y2 = common._rand((6, 6), seed, "y2")
    l2 = ocrit.RnCLoss()(f2, y2)
    l2.backward()
    assert check.rel_err([float(l2)], DATA["crit/rnc6"]) < 1e-5
    assert check.scaled_err(f2.grad.numpy(), DATA["crit/rnc6_grad"]) < 1e-5
    rm = ocrit.RoiMSE(torch.tensor([225.0] * 36), common.ROI_INDICES, voxel_wise=False)
    assert check.rel_err([float(rm(pred.detach(), tau, roi))], DATA["crit/roimse_mean"]) < 1e-5
    assert ocrit.RnCLoss()(feats[:1].detach(), covars[:1, -1].float()) == 0.0 and float(DATA["crit/rnc1"][0]) == 0.0


def test_metrics_oracle_matches_reference_golden():
    """oracle/metrics.py against the values the reference's own calc_roi_metrics / metric lines produced
    (tests/golden/make_metrics_golden.py)."""
    import numpy as np
    from oracle import metrics as ometrics
    from tests.golden import make_metrics_golden
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "metrics_golden.npz"))
    for name in ("m32", "m48"):
        batch, d, h, w, seed = (int(v) for v in gold[f"{name}/cfg"])
        pred, tau, roi = make_metrics_golden.case(batch, (d, h, w), seed)
        vol = [float(v) for v in ometrics.volume_metrics(pred, tau)]
        np.testing.assert_allclose(vol, gold[f"{name}/volume"], rtol=1e-5)
        got = ometrics.calc_roi_metrics(common.ROI_INDICES, tau, roi, pred)
        for key, t in zip(("roi_maes", "roi_mapes", "roi_rses", "roi_wrrmses", "roi_nonnan"), got):
            np.testing.assert_allclose(t.double().numpy(), gold[f"{name}/{key}"], rtol=1e-5, equal_nan=True)


def test_pad_volume_oracle_matches_reference_golden():
    """oracle/prepare.py::pad_volume against outputs of the reference's own data_util.pad_volume
    (tests/golden/make_prepare_golden.py), bit for bit; the host geometry of the CUDA path predicts the same shapes."""
    import numpy as np
    from coma_unet_b200 import prepare_geometry
    from oracle import prepare as oprep
    from tests.golden import make_prepare_golden as mk
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "prepare_golden.npz"))
    for name, (shape, target) in mk.CASES.items():
        got = oprep.pad_volume(target)(mk.case_input(name)).numpy()
        assert got.shape == gold[name].shape and np.array_equal(got, gold[name]), name
        if shape[1] != target[-3]:          # apply_transforms pads only then (VolumeDataset.py:261-264)
            _, out, _, _ = prepare_geometry(shape[1:], (2.0, 2.0, 2.0), True, (2.0, 2.0, 2.0), target)
            assert tuple(out) == gold[name].shape[1:], name


def test_resample_oracle_known_answers():
    """Known answers of the nearest-neighbour resample (oracle/prepare.py::resize_volume, unpinned restatement of SimpleITK):
    1 mm -> 2 mm keeps every other voxel, equal spacing is the identity, coarse -> fine repeats voxels, and output voxels whose
    centre falls outside the image take the default value."""
    import numpy as np
    from oracle import prepare as oprep
    a = np.arange(6 * 8 * 10, dtype=np.float32).reshape(6, 8, 10)
    assert np.array_equal(oprep.resize_volume(a, (1.0, 1.0, 1.0)), a[::2, ::2, ::2])
    assert np.array_equal(oprep.resize_volume(a, (2.0, 2.0, 2.0)), a)
    up = oprep.resize_volume(a, (4.0, 4.0, 4.0), default_value=-1.0)                  # 4 mm -> 2 mm: c = r / 2, floor(c + .5)
    assert up.shape == (12, 16, 20)
    assert np.array_equal(up[:-1, :-1, :-1][0::2, 0::2, 0::2], a[:, :, :][:6, :8, :10])
    assert np.all(up[-1] == -1.0) and np.all(up[:, -1] == -1.0) and np.all(up[:, :, -1] == -1.0)   # c = size - 0.5: outside
    odd = oprep.resize_volume(np.arange(5, dtype=np.float32).reshape(1, 1, 5), (1.0, 2.0, 2.0))
    assert odd.shape == (1, 1, 2) and odd.ravel().tolist() == [0.0, 2.0]              # round(2.5) = 2 (numpy half-even)
