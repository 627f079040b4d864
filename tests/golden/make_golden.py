"""Generate the golden fixtures by running the REFERENCE's own Python on CPU.  Run in the build
container only (needs /root/reference); the fixtures it writes are committed.

    python -m tests.golden.make_golden            # from the repo root

How the reference is made importable without touching it:

* ``attn_unet_data_parallel`` imports ``monai`` (not installed, unpinned) and ``CondConv`` (missing
  from the reference).  ``oracle.monai_blocks`` / ``oracle.cond_conv`` are injected under those
  names, so what the fixtures pin is the reference's own model code (module tree, recursion and
  tuple unrolling, prompt / ROI-painting / modulator logic, return conventions) running on top of
  the restated blocks.  ``np.int`` (:136) is restored for numpy >= 1.24.
* ``criterions`` imports ``data_util`` / ``VolumeDataset`` (heavy I/O deps) -> empty stubs.  Its
  ``torch.zeros(..., device=roi.get_device())`` idiom (:182) raises on CPU tensors (device -1), so
  the module's ``torch`` global is wrapped by a proxy that maps device -1 to "cpu".  Nothing else
  is altered.
* The reference hard-codes 128^3 prompt parameters (:544-545,610).  For the 32^3 cases the prompt
  ``nn.Parameter`` attributes of the constructed *instance* are replaced; the 128^3 case runs it as is.
"""
from __future__ import annotations

import json
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import cond_conv, monai_blocks  # noqa: E402
from tests.golden import common  # noqa: E402

REF = "/root/reference"


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def import_reference():
    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(np, "float"):
        np.float = float
    _module("monai")
    _module("monai.networks")
    _module("monai.networks.blocks")
    _module("monai.networks.blocks.convolutions", Convolution=monai_blocks.Convolution)
    _module("monai.networks.layers")
    _module("monai.networks.layers.factories", Norm=monai_blocks.Norm)
    _module("monai.networks.nets", attentionunet=monai_blocks)
    _module("monai.metrics")
    _module("monai.metrics.regression", SSIMMetric=object)
    _module("CondConv", CondConvolution=cond_conv.CondConvolution, CondConvBlock=cond_conv.CondConvBlock)
    _module("create_roi_suvr_csv")
    _module("visualization_util", loss_graph=None, metric_graph=None, plot_mae_progression_chart=None,
            boxplot_roi_value_progression=None)
    _module("data_util", save_attention_coeffs=lambda *a, **k: None)
    _module("VolumeDataset")
    sys.path.insert(0, REF)
    import attn_unet_data_parallel as ref_model
    import criterions as ref_crit

    class TorchProxy:
        def __getattr__(self, name):
            return getattr(torch, name)

        @staticmethod
        def _fix(kw):
            if kw.get("device", None) == -1:
                kw["device"] = "cpu"
            return kw

        def zeros(self, *a, **kw):
            return torch.zeros(*a, **self._fix(kw))

        def ones(self, *a, **kw):
            return torch.ones(*a, **self._fix(kw))

    ref_crit.torch = TorchProxy()
    return ref_model, ref_crit


def sample(t: torch.Tensor, n=4096):
    flat = t.detach().reshape(-1).double()
    idx = torch.linspace(0, flat.numel() - 1, min(n, flat.numel())).long()
    return {"shape": list(t.shape), "mean": float(flat.mean()), "std": float(flat.std()) if flat.numel() > 1 else 0.0,
            "absmax": float(flat.abs().max()), "idx": idx.numpy(), "val": flat[idx].float().numpy()}


def pack(prefix, d, out):
    for k, v in d.items():
        out[f"{prefix}/{k}"] = np.asarray(v)


def build_ref_model(ref_model, channels, prompt_shape):
    with open(os.devnull, "w") as devnull:   # the reference constructor prints its ROI dict (:607)
        stdout, sys.stdout = sys.stdout, devnull
        try:
            m = ref_model.ContrastiveAttentionUNET_DP(3, 1, 1, channels, [2] * 5, latent_spaces=[2048] * 5,
                                                      conditional=True, decoder_ds=False)
        finally:
            sys.stdout = stdout
    if tuple(prompt_shape) != (128, 128, 128):
        ps = tuple(prompt_shape)
        for name in ("pos_dynamic_prompt", "neg_dynamic_prompt", "general_dynamic_prompt"):
            setattr(m, name, nn.Parameter(torch.randn(1, 1, *ps)))
        m.reweigh = nn.Parameter(torch.ones(ps))
        m.pos_reweigh = nn.Parameter(torch.ones((1, *ps)))
        m.neg_reweigh = nn.Parameter(torch.ones((1, *ps)))
    m.set_save_attn(None)   # validation.py:156
    return m


def build_ref_criterion(ref_crit):
    gen = ref_crit.RoiMSE(torch.tensor([225.0] * 36), common.ROI_INDICES, voxel_wise=False)   # validation.py:130,146
    crit = ref_crit.GenerativeContrastiveLoss(ref_crit.RnCLoss(), gen, nn.TripletMarginLoss(1), 0., 1.)  # :154
    crit.gen_loss.batch_reduction = None   # attn_unet_data_parallel.py:717
    return crit


PROBE_PARAMS = [
    "model.0.conv.0.conv.weight", "model.0.conv.1.adn.N.weight", "model.0.conv.0.film.2.weight",
    "model.1.submodule.0.conv.0.conv.weight", "model.1.attention.W_g.0.conv.weight", "model.1.attention.psi.1.weight",
    "model.1.upconv.up.conv.weight", "model.1.upconv.up.film.0.weight", "model.1.merge.conv.weight",
    "model.1.merge.adn.A.weight", "model.1.submodule.1.submodule.1.submodule.1.submodule.conv.1.conv.weight",
    "model.2.conv.weight", "model.2.routing.weight", "projection_heads.4.conv.conv.0.conv.weight",
    "fusion_layer.blocks.0.conv.weight", "deep_modulator_3c.blocks.2.conv.weight", "final_pred_head.conv.weight",
    "final_pred_head.adn.A.weight", "pos_dynamic_prompt", "neg_dynamic_prompt", "general_dynamic_prompt",
]


def train_case(ref_model, ref_crit, name, channels, shape, batch, seed, out, meta):
    torch.manual_seed(0)
    m = build_ref_model(ref_model, channels, shape)
    common.fill_deterministic(m, seed)
    mri, tau, roi, covars, dicts = common.synthetic_batch(batch, shape, seed)
    crit = build_ref_criterion(ref_crit)
    m.train(True)
    pred, projected, final_repr = m(mri, covars, roi_pred_dicts=dicts, sample_roi_mask=roi)
    features, labels = projected[-1], covars[:, -1].float()                  # attn_unet_data_parallel.py:842-845
    zeros = torch.zeros(final_repr.size())
    loss, gen, ps, ds = crit(pred, tau, roi, (final_repr, zeros, zeros), (features, labels))
    loss.backward()
    pack(f"{name}/pred", sample(pred, 1 << 20), out)
    for i, p in enumerate(projected):
        pack(f"{name}/proj{i}", sample(p), out)
    pack(f"{name}/final_repr", sample(final_repr), out)
    out[f"{name}/loss"] = np.array([float(loss.detach()), float(ps), float(ds)], dtype=np.float64)
    out[f"{name}/gen"] = gen.detach().numpy()
    params = dict(m.named_parameters())
    nograd = sorted(k for k, p in params.items() if p.grad is None)
    for k in PROBE_PARAMS:
        if params[k].grad is not None:
            pack(f"{name}/grad/{k}", sample(params[k].grad), out)
    sd = m.state_dict()
    for k in ("model.0.conv.0.adn.N.running_mean", "model.0.conv.0.adn.N.running_var",
              "model.1.attention.W_g.1.running_var", "model.0.conv.0.adn.N.num_batches_tracked"):
        out[f"{name}/buf/{k}"] = sd[k].numpy().copy()
    meta[name] = {"kind": "train", "channels": channels, "shape": list(shape), "batch": batch, "seed": seed,
                  "no_grad_params": nograd, "state_keys": list(sd.keys())}
    # eval-mode forward on the same (now BN-updated) weights
    m.eval()
    m.set_training(False)
    with torch.no_grad():
        pred_eval = m(mri, covars, roi_pred_dicts=dicts, sample_roi_mask=roi)
    pack(f"{name}/pred_eval", sample(pred_eval, 1 << 20), out)


def eval_case(ref_model, name, channels, shape, batch, seed, out, meta):
    torch.manual_seed(0)
    m = build_ref_model(ref_model, channels, shape)
    common.fill_deterministic(m, seed)
    mri, tau, roi, covars, dicts = common.synthetic_batch(batch, shape, seed)
    m.eval()
    m.set_training(False)
    with torch.no_grad():
        pred = m(mri, covars, roi_pred_dicts=dicts, sample_roi_mask=roi)
    pack(f"{name}/pred_eval", sample(pred, 1 << 15), out)
    meta[name] = {"kind": "eval", "channels": channels, "shape": list(shape), "batch": batch, "seed": seed}


def criterion_case(ref_crit, out, meta):
    name, B, shape, seed = "crit", 3, (16, 16, 16), 7
    mri, tau, roi, covars, _ = common.synthetic_batch(B, shape, seed)
    pred = (tau + 0.2 * common._randn(tau.shape, seed, "pred")).requires_grad_(True)
    feats = common._randn((B, 512), seed, "feats").abs().requires_grad_(True)
    final = common._randn((B, 1, 1, 1, 2048), seed, "final").requires_grad_(True)
    crit = build_ref_criterion(ref_crit)
    zeros = torch.zeros(final.size())
    loss, gen, ps, ds = crit(pred, tau, roi, (final, zeros, zeros), (feats, covars[:, -1].float()))
    loss.backward()
    out[f"{name}/loss"] = np.array([float(loss.detach()), float(ps), float(ds)], dtype=np.float64)
    out[f"{name}/gen"] = gen.detach().numpy()
    out[f"{name}/dpred"] = pred.grad.numpy()
    out[f"{name}/dfeats"] = feats.grad.numpy()
    # RnC alone, a larger batch, and the mean-reduced RoiMSE
    f2 = common._randn((6, 64), seed, "f2").requires_grad_(True)
    y2 = common._rand((6, 6), seed, "y2")
    l2 = ref_crit.RnCLoss()(f2, y2)
    l2.backward()
    out[f"{name}/rnc6"] = np.array([float(l2)])
    out[f"{name}/rnc6_grad"] = f2.grad.numpy()
    rm = ref_crit.RoiMSE(torch.tensor([225.0] * 36), common.ROI_INDICES, voxel_wise=False)
    out[f"{name}/roimse_mean"] = np.array([float(rm(pred.detach(), tau, roi))])
    out[f"{name}/rnc1"] = np.array([float(ref_crit.RnCLoss()(feats[:1].detach(), covars[:1, -1].float()))])
    meta[name] = {"kind": "criterion", "batch": B, "shape": list(shape), "seed": seed}


def main():
    torch.set_num_threads(os.cpu_count())
    ref_model, ref_crit = import_reference()
    out, meta = {}, {}
    criterion_case(ref_crit, out, meta)
    train_case(ref_model, ref_crit, "train32", [16, 32, 64, 128, 256], (32, 32, 32), 2, 12, out, meta)
    train_case(ref_model, ref_crit, "train32_b1", [8, 16, 32, 64, 128], (32, 32, 32), 1, 13, out, meta)
    eval_case(ref_model, "eval128", [8, 16, 32, 64, 128], (128, 128, 128), 1, 17, out, meta)
    np.savez_compressed(os.path.join(HERE, "golden.npz"), **out)
    with open(os.path.join(HERE, "golden_meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    for k, v in meta.items():
        print(k, "no-grad params:", v.get("no_grad_params"))
    print("wrote", len(out), "arrays;", os.path.getsize(os.path.join(HERE, "golden.npz")) / 1e6, "MB")


if __name__ == "__main__":
    main()
